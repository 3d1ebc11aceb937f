"""Host-side construction of house / cluster / power-grid populations (reset-time work).

Two builders:

* :func:`reference_order_population` consumes python's global ``random`` stream in exactly
  the order ``MADemandResponseEnv.build_environment`` does (env/MA_DemandResponse.py:98-133,
  utils.py:573-709), so that ``random.seed(s)`` followed by constructing
  this environment yields the same houses, start date, phase, outdoor temperature, signal
  ratio and perlin seed as the reference would.
* :func:`synthetic_population` draws E x N houses with numpy for the throughput workloads
  (SURVEY section 8d "Concrete synthetic inputs").

Both return a dict of numpy arrays: per-house ``[E, N]`` keys ``ua cm ca hm cap target deadband
t_air t_mass lockout_dur sso on lockout``; per-env ``[E]`` keys ``t_epoch phase od_temp
artificial_ratio max_power base_power time_since_interp signal cluster_power solar_gain
perlin_seed``.
"""
import datetime as _dt
import math
import random as _random

import numpy as np

from .config_flatten import FlatConfig, epoch_seconds, from_epoch

HOUSE_F = ("ua", "cm", "ca", "hm", "cap", "target", "deadband", "t_air", "t_mass")
HOUSE_I = ("lockout_dur", "sso", "on", "lockout")
ENV_F = ("phase", "od_temp", "artificial_ratio", "max_power", "base_power", "signal", "cluster_power", "solar_gain",
         "perlin_seed")
ENV_I = ("t_epoch", "time_since_interp")


def od_temperature(flat: FlatConfig, date_time: _dt.datetime, phase: float, noise: float) -> float:
    """ClusterHouses.compute_OD_temp (:1057-1081) for the reset-time value."""
    amplitude = (flat.day_temp - flat.night_temp) / 2
    bias = (flat.day_temp + flat.night_temp) / 2
    delay = -6 + phase
    time_day = date_time.hour + date_time.minute / 60.0
    return float(amplitude * np.sin(2 * np.pi * (time_day + delay) / 24) + bias + noise)


def reference_order_population(flat: FlatConfig, rng=_random):
    """One env; returns (population dict with E = 1, explicit comm table or None)."""
    n = flat.n_houses
    hd, vd = flat.house_def, flat.hvac_def
    nh = flat.noise_house["noise_parameters"][flat.noise_house["noise_mode"]]
    nv = flat.noise_hvac["noise_parameters"][flat.noise_hvac["noise_mode"]]
    pop = {k: np.zeros((1, n), np.float64) for k in HOUSE_F}
    pop.update({k: np.zeros((1, n), np.int64) for k in HOUSE_I})
    for i in range(n):
        # apply_house_noise, utils.py:623-666
        pop["t_air"][0, i] = hd["init_air_temp"] + abs(rng.gauss(0, nh["std_start_temp"]))
        pop["t_mass"][0, i] = hd["init_mass_temp"] + abs(rng.gauss(0, nh["std_start_temp"]))
        pop["target"][0, i] = hd["target_temp"] + abs(rng.gauss(0, nh["std_target_temp"]))
        lo, hi = nh["factor_thermo_low"], nh["factor_thermo_high"]
        pop["ua"][0, i] = hd["Ua"] * rng.triangular(lo, hi, 1)
        pop["cm"][0, i] = hd["Cm"] * rng.triangular(lo, hi, 1)
        pop["ca"][0, i] = hd["Ca"] * rng.triangular(lo, hi, 1)
        pop["hm"][0, i] = hd["Hm"] * rng.triangular(lo, hi, 1)
        pop["deadband"][0, i] = hd["deadband"]
        # apply_hvac_noise, utils.py:669-676
        pop["cap"][0, i] = rng.choices(nv["cooling_capacity_list"][vd["cooling_capacity"]])[0]
    # get_random_date_time, utils.py:701-709
    start = flat.start_datetime
    if flat.start_datetime_mode == "random":
        days = rng.randrange(364)
        seconds = rng.randrange(60 * 60 * 24)
        start = start + _dt.timedelta(days=days, seconds=seconds)
    # HVAC.__init__, :430-434 (randint is drawn even when the noise is 0)
    for i in range(n):
        dur = vd["lockout_duration"] + rng.randint(-vd["lockout_noise"], vd["lockout_noise"])
        if dur < 0:
            raise ValueError("HVAC id: {} - Lockout duration must be positive. Current value: {}.".format(i, dur))
        if pop["cap"][0, i] < 0:
            raise ValueError("HVAC id: {} - Cooling capacity must be positive.".format(i))
        pop["lockout_dur"][0, i] = dur
        pop["sso"][0, i] = dur
    # ClusterHouses.__init__, :789-793
    phase = rng.random() * 24 if flat.random_phase_offset else 0
    od = od_temperature(flat, start, phase, rng.gauss(0, flat.temp_std))
    table = flat.explicit_comm_table(sampler=lambda possible, k: rng.sample(possible, k=k))
    # PowerGrid.__init__, :1116 and :1182-1184
    ratio = flat.artificial_ratio * flat.artificial_signal_ratio_range ** (rng.random() * 2 - 1)
    perlin_seed = rng.random() if "perlin" in flat.signal_mode_name else 0.0
    env = {
        "t_epoch": np.array([epoch_seconds(start)], np.int64),
        "phase": np.array([phase], np.float64),
        "od_temp": np.array([od], np.float64),
        "artificial_ratio": np.array([ratio], np.float64),
        "max_power": np.array([float(np.sum(pop["cap"][0] / flat.hvac_cop))], np.float64),
        "base_power": np.zeros(1), "signal": np.zeros(1), "cluster_power": np.zeros(1), "solar_gain": np.zeros(1),
        "time_since_interp": np.array([flat.interp_update_period + 1], np.int64),
        "perlin_seed": np.array([perlin_seed], np.float64),
    }
    # max_power is a sequential sum in id order (:796-802)
    mp = 0
    for i in range(n):
        mp += pop["cap"][0, i] / flat.hvac_cop
    env["max_power"][0] = mp
    pop.update(env)
    return pop, table


def synthetic_population(flat: FlatConfig, n_envs: int, seed: int = 1234, heterogeneous: bool = True,
                         start_temp_std: float = 5.0, target_std: float = 1.0, lockout_noise: int = 8):
    """E x N synthetic houses: T = 20 + |N(0, 5)|, target = 20 (+ |N(0,1)| if
    heterogeneous), thermal parameters x Triangular(0.8, 1.2, 1), capacity from the big-noise list,
    lockout 40 +- 8 s, random start second within the year, all HVACs off."""
    rng = np.random.default_rng(seed)
    e, n = int(n_envs), flat.n_houses
    hd, vd = flat.house_def, flat.hvac_def
    shape = (e, n)
    pop = {}
    pop["t_air"] = hd["init_air_temp"] + np.abs(rng.normal(0, start_temp_std, shape))
    pop["t_mass"] = hd["init_mass_temp"] + np.abs(rng.normal(0, start_temp_std, shape))
    if heterogeneous:
        pop["target"] = hd["target_temp"] + np.abs(rng.normal(0, target_std, shape))
        for key, name in (("ua", "Ua"), ("cm", "Cm"), ("ca", "Ca"), ("hm", "Hm")):
            pop[key] = hd[name] * rng.triangular(0.8, 1.0, 1.2, shape)
        pop["cap"] = rng.choice(np.array([10000.0, 12500.0, 15000.0, 17500.0, 20000.0]), shape)
        pop["lockout_dur"] = (vd["lockout_duration"] + rng.integers(-lockout_noise, lockout_noise + 1, shape)).astype(np.int64)
    else:
        pop["target"] = np.full(shape, float(hd["target_temp"]))
        for key, name in (("ua", "Ua"), ("cm", "Cm"), ("ca", "Ca"), ("hm", "Hm")):
            pop[key] = np.full(shape, float(hd[name]))
        pop["cap"] = np.full(shape, float(vd["cooling_capacity"]))
        pop["lockout_dur"] = np.full(shape, int(vd["lockout_duration"]), np.int64)
    pop["deadband"] = np.full(shape, float(hd["deadband"]))
    pop["sso"] = pop["lockout_dur"].copy()
    pop["on"] = np.zeros(shape, np.int64)
    pop["lockout"] = np.zeros(shape, np.int64)
    start = epoch_seconds(flat.start_datetime)
    pop["t_epoch"] = (start + rng.integers(0, 364 * 86400, e)).astype(np.int64)
    pop["phase"] = rng.random(e) * 24 if flat.random_phase_offset else np.zeros(e)
    sod = pop["t_epoch"] % 86400
    time_day = sod // 3600 + ((sod % 3600) // 60) / 60.0
    amplitude, bias = (flat.day_temp - flat.night_temp) / 2, (flat.day_temp + flat.night_temp) / 2
    pop["od_temp"] = amplitude * np.sin(2 * np.pi * (time_day - 6 + pop["phase"]) / 24) + bias + rng.normal(0, flat.temp_std, e)
    rr = flat.artificial_signal_ratio_range
    pop["artificial_ratio"] = flat.artificial_ratio * rr ** (rng.random(e) * 2 - 1)
    pop["max_power"] = (pop["cap"] / flat.hvac_cop).sum(axis=1)
    for k in ("base_power", "signal", "cluster_power", "solar_gain"):
        pop[k] = np.zeros(e)
    pop["time_since_interp"] = np.full(e, flat.interp_update_period + 1, np.int64)
    pop["perlin_seed"] = rng.random(e)
    return pop


def concat_populations(pops):
    """Stacks single-env populations along the env axis."""
    out = {}
    for k in pops[0]:
        out[k] = np.concatenate([np.atleast_1d(p[k]) if np.ndim(p[k]) <= 1 else p[k] for p in pops], axis=0)
    return out


def shard_population(pop, rank: int, world: int):
    """Contiguous env-axis shard of rank `rank`: clusters never span GPUs."""
    e = len(pop["t_epoch"])
    lo, hi = (e * rank) // world, (e * (rank + 1)) // world
    return {k: np.asarray(v)[lo:hi] for k, v in pop.items()}


def synthetic_interp_table(seed: int = 0, size: int = 4199040):
    """Stand-in for monteCarlo/mergedGridSearchResultFinal.npy (a missing large blob of the
    reference): uniform [0, 6000) W over the full grid shape, flat C order."""
    return np.random.default_rng(seed).uniform(0, 6000, size)


def population_spec(flat: FlatConfig):
    """MdrPopulationSpec for the device-side population draw (mdr_populate): the same config entries
    reference_order_population() reads, flattened for the kernel."""
    from . import _lib
    hd, vd = flat.house_def, flat.hvac_def
    nh = flat.noise_house["noise_parameters"][flat.noise_house["noise_mode"]]
    nv = flat.noise_hvac["noise_parameters"][flat.noise_hvac["noise_mode"]]
    sp = _lib.MdrPopulationSpec()
    sp.init_air_temp, sp.init_mass_temp = float(hd["init_air_temp"]), float(hd["init_mass_temp"])
    sp.target_temp, sp.deadband = float(hd["target_temp"]), float(hd["deadband"])
    sp.ua, sp.cm, sp.ca, sp.hm = float(hd["Ua"]), float(hd["Cm"]), float(hd["Ca"]), float(hd["Hm"])
    sp.std_start_temp, sp.std_target_temp = float(nh["std_start_temp"]), float(nh["std_target_temp"])
    sp.factor_thermo_low, sp.factor_thermo_high = float(nh["factor_thermo_low"]), float(nh["factor_thermo_high"])
    caps = list(nv["cooling_capacity_list"][vd["cooling_capacity"]])
    if not 1 <= len(caps) <= 8:
        raise ValueError("cooling_capacity_list must have 1..8 entries for the device-side draw, got %d" % len(caps))
    for i, c in enumerate(caps):
        if c < 0:
            raise ValueError("HVAC - Cooling capacity must be positive.")
        sp.cap_list[i] = float(c)
    sp.n_cap = len(caps)
    sp.lockout_duration, sp.lockout_noise = int(vd["lockout_duration"]), int(vd["lockout_noise"])
    if sp.lockout_duration - sp.lockout_noise < 0:
        raise ValueError("HVAC - Lockout duration must be positive.")
    sp.random_start = int(flat.start_datetime_mode == "random")
    sp.start_epoch = epoch_seconds(flat.start_datetime)
    sp.random_phase = int(bool(flat.random_phase_offset))
    sp.interp_update_period = int(flat.interp_update_period)
    sp.artificial_ratio, sp.artificial_ratio_range = float(flat.artificial_ratio), float(flat.artificial_signal_ratio_range)
    return sp
