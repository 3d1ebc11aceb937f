"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference step path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``marl-demandresponse-original_b200``) never does; it fails loudly when its CUDA library
is missing instead of falling back to this code.

What it restates (all citations relative to the reference repository root):

* ``HVAC.step``                     env/MA_DemandResponse.py:463-492
* ``HVAC.get_Q/power_consumption``  env/MA_DemandResponse.py:494-523
* ``SingleHouse.update_temperature``env/MA_DemandResponse.py:664-738
* ``house_solar_gain``              utils.py:1277-1350
* ``ClusterHouses.compute_OD_temp`` env/MA_DemandResponse.py:1057-1081
* ``ClusterHouses.step``            env/MA_DemandResponse.py:1005-1055
* ``build_agent_comm_links``        env/MA_DemandResponse.py:806-902
* ``make_cluster_obs_dict/message`` env/MA_DemandResponse.py:904-1003, 624-662
* ``compute_rewards`` + penalties   env/MA_DemandResponse.py:234-373, utils.py:1266-1274
* ``PowerGrid.step/interpolatePower`` env/MA_DemandResponse.py:1195-1316
* ``clipInterpolationPoint``        utils.py:1214-1221
* ``interpolateGridFast``           monteCarlo/interpolation.py:113-142 (+ scipy interpn
                                    linear path, scipy/interpolate/_rgi.py:520-549)
* ``Perlin.calculate_noise``        utils.py:1231-1253 (octave mixing only; the lattice
                                    noise itself is third party and *replayed*)
* ``normStateDict``                 utils.py:740-880
* ``MADemandResponseEnv.step``      env/MA_DemandResponse.py:174-210 (sequencing)

Pinning: checked in ``tests/test_oracle.py`` against (1) the known-answer vectors that
SURVEY.md section 8c lists (produced by the reference), and (2) the golden traces under
``tests/golden/*.npz`` produced by running the imported, unmodified reference in the
build container with ``oracle/make_golden.py``.

Random draws are never re-implemented here: the outdoor-temperature Gaussian draw, the
perlin value, the interpolation sample ids and the message-drop mask are *inputs* of
``step`` (host-replayed), exactly as the GPU environment receives them in parity mode.

State layout ("snapshot"): a dict of numpy arrays; per-house arrays are ``[E, N]``, per-env
arrays ``[E]``.  See ``snapshot_keys``.
"""
from __future__ import annotations

import datetime as _dt
import itertools
import math

import numpy as np

EPOCH = _dt.datetime(1970, 1, 1)

HOUSE_KEYS_F = ("ua", "cm", "ca", "hm", "cap", "cop", "latent", "target", "deadband", "t_air", "t_mass")
HOUSE_KEYS_I = ("lockout_dur", "sso", "on", "lockout")
ENV_KEYS = (
    "t_epoch",           # int64 naive seconds since 1970-01-01 (current datetime)
    "phase",             # OD-temperature phase offset (hours)
    "od_temp",           # current outdoor temperature (drawn at the previous step)
    "artificial_ratio",  # PowerGrid.artificial_ratio
    "max_power",         # ClusterHouses.max_power
    "base_power",        # PowerGrid.base_power
    "time_since_interp", # PowerGrid.time_since_last_interp (int seconds)
    "signal",            # PowerGrid.current_signal
    "cluster_power",     # ClusterHouses.cluster_hvac_power
    "solar_gain",        # solar gain used in the last thermal update
)
snapshot_keys = HOUSE_KEYS_F + HOUSE_KEYS_I + ENV_KEYS


def to_datetime(t_epoch: int) -> _dt.datetime:
    return EPOCH + _dt.timedelta(seconds=int(t_epoch))


def from_datetime(d: _dt.datetime) -> int:
    return int((d - EPOCH).total_seconds())


# --------------------------------------------------------------------------------------
# scalar helpers
# --------------------------------------------------------------------------------------
def deadband_l2(target, deadband, value):
    """utils.py:1266-1274 (vectorised, strict inequalities, same expressions)."""
    target = np.asarray(target, dtype=np.float64)
    value = np.asarray(value, dtype=np.float64)
    hi = target + deadband / 2
    lo = target - deadband / 2
    out = np.zeros(np.broadcast(target, value).shape, dtype=np.float64)
    above = hi < value
    below = (~above) & (lo > value)
    out = np.where(above, (value - hi) ** 2, out)
    out = np.where(below, (lo - value) ** 2, out)
    return out


_SOLAR_COEFF = (
    4.36579418e01, 1.58055357e02, 8.76635241e01, -4.55944821e01, 3.24275366e00,
    -4.56096472e-01, -1.47795612e01, 4.68950855e00, -3.73313090e01, 5.78827663e00,
    1.04354810e00, 2.12969604e-02, 2.58881400e-03, -5.11397219e-04, 1.56398008e-02,
    -1.18302764e-01, -2.71446436e-01, -3.97855577e-02,
)


def house_solar_gain(date_time: _dt.datetime, window_area: float, shading_coeff: float) -> float:
    """utils.py:1277-1350: CIBSE polynomial in (hour-7.5, month+day/30-1), same term order."""
    x = date_time.hour + date_time.minute / 60 - 7.5
    if x < 0 or x > 10:
        scl = 0
    else:
        y = date_time.month + date_time.day / 30 - 1
        c = _SOLAR_COEFF
        scl = (
            c[0] + x * c[1] + y * c[2] + x**2 * c[3] + x**2 * y * c[4] + x**2 * y**2 * c[5]
            + y**2 * c[6] + x * y**2 * c[7] + x * y * c[8] + x**3 * c[9] + y**3 * c[10]
            + x**3 * y * c[11] + x**3 * y**2 * c[12] + x**3 * y**3 * c[13] + x**2 * y**3 * c[14]
            + x * y**3 * c[15] + x**4 * c[16] + y**4 * c[17]
        )
    return window_area * shading_coeff * scl


def od_temp_model(date_time: _dt.datetime, day_temp, night_temp, phase, noise) -> float:
    """env/MA_DemandResponse.py:1070-1081; `noise` is the replayed random.gauss(0, std) draw."""
    amplitude = (day_temp - night_temp) / 2
    bias = (day_temp + night_temp) / 2
    delay = -6 + phase
    time_day = date_time.hour + date_time.minute / 60.0
    temperature = amplitude * np.sin(2 * np.pi * (time_day + delay) / 24) + bias
    temperature += noise
    return float(temperature)


def hvac_step(on, lockout, sso, lockout_dur, command, dt):
    """env/MA_DemandResponse.py:463-492, vectorised; integer/bool exact."""
    on = on.astype(bool)
    command = np.asarray(command).astype(bool)
    sso = sso + np.where(~on, dt, 0).astype(sso.dtype)
    lock = ~(on | (sso >= lockout_dur))
    new_on = np.where(lock, False, command)
    sso = np.where(~lock & new_on, 0, sso)
    lock = np.where(~lock & ~new_on & (sso + dt < lockout_dur), True, lock)
    return new_on, lock, sso


def etp_update(t_air, t_mass, od_temp, q_a, ua, ca, hm, cm, dt):
    """env/MA_DemandResponse.py:681-738 in the written operation order (Kelvin offset 273)."""
    od_k = od_temp + 273
    ta_k = t_air + 273
    tm_k = t_mass + 273
    q_m = 0
    a = cm * ca / hm
    b = cm * (ua + hm) / hm + ca
    c = ua
    d = q_m + q_a + ua * od_k
    g = q_m / hm
    r1 = (-b + np.sqrt(b**2 - 4 * a * c)) / (2 * a)
    r2 = (-b - np.sqrt(b**2 - 4 * a * c)) / (2 * a)
    dta0 = hm * tm_k / ca - (ua + hm) * ta_k / ca + ua * od_k / ca + q_a / ca
    a1 = (r2 * ta_k - dta0 - r2 * d / c) / (r2 - r1)
    a2 = ta_k - d / c - a1
    a3 = r1 * ca / hm + (ua + hm) / hm
    a4 = r2 * ca / hm + (ua + hm) / hm
    new_ta = a1 * np.exp(r1 * dt) + a2 * np.exp(r2 * dt) + d / c
    new_tm = a1 * a3 * np.exp(r1 * dt) + a2 * a4 * np.exp(r2 * dt) + g + d / c
    return new_ta - 273, new_tm - 273


# --------------------------------------------------------------------------------------
# neighbour tables (env/MA_DemandResponse.py:806-902)
# --------------------------------------------------------------------------------------
def comm_links(mode: str, n: int, nb_agents_comm: int, row_size=5, distance_comm=2, sampler=None):
    """Returns int32 [N, C].  `sampler(possible_ids, k)` replays random.sample for random_fixed."""
    nb_comm = int(min(nb_agents_comm, n - 1))
    if mode == "neighbours":
        out = []
        for i in range(n):
            before = [(i - nb_comm // 2 + k) % n for k in range(nb_comm // 2)]
            after = [(i + 1 + k) % n for k in range(int(math.ceil(nb_comm / 2)))]
            out.append(before + after)
    elif mode == "closed_groups":
        out = []
        for i in range(n):
            base = i - (i % (nb_comm + 1))
            if base + nb_comm <= n:
                ids = [base + k for k in range(nb_agents_comm + 1)]
            else:
                ids = [n - nb_comm - 1 + k for k in range(nb_comm + 1)]
            ids.remove(i)
            out.append(ids)
    elif mode == "random_fixed":
        out = []
        for i in range(n):
            possible = list(range(n))
            possible.remove(i)
            out.append(list(sampler(possible, nb_comm)))
    elif mode == "neighbours_2D":
        if n % row_size != 0:
            raise ValueError("Neighbours 2D row_size must be a divisor of nb_agents")
        max_y = n // row_size
        if distance_comm >= (row_size + 1) // 2 or distance_comm >= (max_y + 1) // 2:
            raise ValueError("Neighbours 2D distance_comm too large")
        pattern = [
            (dx, dy)
            for dx in range(-distance_comm, distance_comm + 1)
            for dy in range(-distance_comm, distance_comm + 1)
            if abs(dx) + abs(dy) <= distance_comm and (dx != 0 or dy != 0)
        ]
        out = []
        for i in range(n):
            x, y = i % row_size, i // row_size
            ids = []
            for dx, dy in pattern:
                xn, yn = x + dx, y + dy
                if xn < 0:
                    xn += row_size
                if xn >= row_size:
                    xn -= row_size
                if yn < 0:
                    yn += max_y
                if yn >= max_y:
                    yn -= max_y
                ids.append(yn * row_size + xn)
            out.append(ids)
    elif mode in ("no_message", "random_sample"):
        out = [[] for _ in range(n)]
    else:
        raise ValueError("Cluster property: unknown agents_comm_mode '{}'.".format(mode))
    width = len(out[0]) if out else 0
    return np.asarray(out, dtype=np.int32).reshape(n, width)


# --------------------------------------------------------------------------------------
# interpolation (monteCarlo/interpolation.py:113-142 + scipy linear interpn)
# --------------------------------------------------------------------------------------
class PowerInterp:
    """`table` is the flat fp64 array of monteCarlo/mergedGridSearchResultFinal.npy (C order
    over `dict_keys`); `parameters_dict` is interp_parameters_dict.json."""

    def __init__(self, table, parameters_dict, dict_keys):
        self.keys = list(dict_keys)
        self.grid = [np.asarray(parameters_dict[k], dtype=np.float64) for k in self.keys]
        self.values = np.asarray(table, dtype=np.float64).reshape([len(g) for g in self.grid])

    def clip(self, point: dict) -> dict:
        """utils.py:1214-1221."""
        out = {}
        for k, v in point.items():
            g = self.grid[self.keys.index(k)]
            out[k] = min(max(v, g.min()), g.max())
        return out

    def fast(self, point: dict) -> float:
        coords = [point[k] for k in self.keys]
        near = [int(np.argmin(np.abs(self.grid[i] - coords[i]))) for i in range(4)]
        i_hvac = int(np.argmin(np.abs(self.grid[7] - coords[7])))
        sub = self.values[near[0], near[1], near[2], near[3]][:, :, :, i_hvac, :, :]
        dims = [4, 5, 6, 8, 9]
        idx, w = [], []
        for d in dims:
            g, x = self.grid[d], coords[d]
            i = int(np.searchsorted(g, x, side="right")) - 1
            i = min(max(i, 0), len(g) - 2)
            idx.append(i)
            w.append((x - g[i]) / (g[i + 1] - g[i]))
        value = 0.0
        for corner in itertools.product((0, 1), repeat=5):
            weight = 1.0
            for bit, y in zip(corner, w):
                weight = weight * (y if bit else 1 - y)
            value = value + sub[tuple(i + b for i, b in zip(idx, corner))] * weight
        return float(value)


def perlin_mix(octave_values, nb_octaves):
    """utils.py:1247-1253: note the precedence of `2**nb_octaves - 1` (= 31 for 5 octaves)."""
    noise = 0
    for j in range(nb_octaves - 1):
        noise += octave_values[j] / (2**j)
    noise += octave_values[-1] / (2**nb_octaves - 1)
    return noise


# --------------------------------------------------------------------------------------
# the environment
# --------------------------------------------------------------------------------------
class OracleEnv:
    """Batched restatement: `snap` holds [E, N] / [E] arrays (see snapshot_keys), `config`
    is the reference's nested config dict (read with the reference's own keys)."""

    def __init__(self, config: dict, snap: dict, comm_table=None, interp: PowerInterp | None = None):
        self.config = config
        env_prop = config["default_env_prop"]
        self.env_prop = env_prop
        self.house_def = config["default_house_prop"]
        self.hvac_def = config["default_hvac_prop"]
        self.cluster_prop = env_prop["cluster_prop"]
        self.grid_prop = env_prop["power_grid_prop"]
        self.reward_prop = env_prop["reward_prop"]
        self.dt = int(env_prop["time_step"])
        self.s = {k: np.array(v, copy=True) for k, v in snap.items()}
        # the reference never noises COP / latent fraction: default them from the config when absent
        for k, name in (("cop", "COP"), ("latent", "latent_cooling_fraction")):
            if k not in self.s:
                self.s[k] = np.full(np.shape(self.s["cap"]), float(self.hvac_def[name]))
        if "solar_gain" not in self.s:
            self.s["solar_gain"] = np.zeros(np.shape(self.s["t_epoch"]))
        for k in HOUSE_KEYS_F:
            self.s[k] = np.atleast_2d(np.asarray(self.s[k], dtype=np.float64))
        for k in HOUSE_KEYS_I:
            self.s[k] = np.atleast_2d(np.asarray(self.s[k])).astype(np.int64)
        for k in ENV_KEYS:
            dt_ = np.int64 if k in ("t_epoch", "time_since_interp") else np.float64
            self.s[k] = np.atleast_1d(np.asarray(self.s[k], dtype=dt_))
        self.E, self.N = self.s["t_air"].shape
        tm = self.cluster_prop["temp_parameters"][self.cluster_prop["temp_mode"]]
        self.day_temp, self.night_temp = tm["day_temp"], tm["night_temp"]
        self.mode = self.cluster_prop["agents_comm_mode"]
        if comm_table is None:
            p2d = self.cluster_prop["agents_comm_parameters"]["neighbours_2D"]
            comm_table = comm_links(
                self.mode, self.N, self.cluster_prop["nb_agents_comm"], p2d["row_size"], p2d["distance_comm"]
            )
        self.comm = np.asarray(comm_table, dtype=np.int64)  # [N, C] or [E, N, C]
        self.interp = interp
        self.nb_features = None

    # -- PowerGrid.step, env/MA_DemandResponse.py:1236-1316 ---------------------------
    def _interpolate_power(self, e, date_time, ids):
        """env/MA_DemandResponse.py:1195-1234; `ids` replays random.choices when N > interp_nb_agents."""
        s = self.s
        ip = self.grid_prop["base_power_parameters"]["interpolation"]
        if self.house_def["solar_gain_bool"]:
            date = date_time.timetuple().tm_yday
            hour = (date_time - date_time.replace(hour=0, minute=0, second=0, microsecond=0)).total_seconds()
        else:
            date, hour = 0.0, 0.0
        n = self.N
        if n <= ip["interp_nb_agents"]:
            ids = list(range(n))
            factor = 1
        else:
            ids = [int(i) for i in ids]
            factor = float(n) / ip["interp_nb_agents"]
        base = 0
        for i in ids:
            point = {
                "date": date,
                "hour": hour,
                "Ua_ratio": s["ua"][e, i] / self.house_def["Ua"],
                "Cm_ratio": s["cm"][e, i] / self.house_def["Cm"],
                "Ca_ratio": s["ca"][e, i] / self.house_def["Ca"],
                "Hm_ratio": s["hm"][e, i] / self.house_def["Hm"],
                "air_temp": s["t_air"][e, i] - s["target"][e, i],
                "mass_temp": s["t_mass"][e, i] - s["target"][e, i],
                "OD_temp": s["od_temp"][e] - s["target"][e, i],
                "HVAC_power": s["cap"][e, i],
            }
            base += self.interp.fast(self.interp.clip(point))
        return base * factor

    def grid_step(self, e, date_time, sig_noise=0.0, interp_ids=None):
        s, gp = self.s, self.grid_prop
        if gp["base_power_mode"] == "constant":
            s["base_power"][e] = gp["base_power_parameters"]["constant"]["avg_power_per_hvac"] * self.N
        elif gp["base_power_mode"] == "interpolation":
            s["time_since_interp"][e] += self.dt
            if s["time_since_interp"][e] >= gp["base_power_parameters"]["interpolation"]["interp_update_period"]:
                s["base_power"][e] = self._interpolate_power(e, date_time, interp_ids)
                s["time_since_interp"][e] = 0
        else:
            raise ValueError("base_power_mode")
        base = s["base_power"][e]
        mode = gp["signal_mode"]
        params = gp["signal_parameters"][mode]
        if mode == "flat":
            sig = base
        elif mode == "sinusoidals":
            amplitudes = [base * r for r in params["amplitude_ratios"]]
            periods = params["periods"]
            if len(periods) != len(amplitudes):
                raise ValueError("periods and amplitude_ratios lists should have the same length")
            time_sec = date_time.hour * 3600 + date_time.minute * 60 + date_time.second
            sig = base
            for a, p in zip(amplitudes, periods):
                sig += a * np.sin(2 * np.pi * time_sec / p)
        elif mode == "regular_steps":
            amplitude = params["amplitude_per_hvac"] * self.N
            ratio = base / amplitude
            period = params["period"]
            time_sec = date_time.hour * 3600 + date_time.minute * 60 + date_time.second
            sig = amplitude * np.heaviside((time_sec % period) - (1 - ratio) * period, 1)
        elif "perlin" in mode:
            amplitude = params["amplitude_ratios"]
            sig = np.maximum(0, base + (base * amplitude * sig_noise))
        else:
            raise ValueError("Invalid power grid signal mode")
        sig = sig * s["artificial_ratio"][e]
        sig = np.minimum(sig, s["max_power"][e])
        s["signal"][e] = sig
        return float(sig)

    # -- reward, env/MA_DemandResponse.py:234-373 --------------------------------------
    def _rewards(self, e, power, signal_old):
        s, rp = self.s, self.reward_prop
        n = self.N
        if rp["sig_penalty_mode"] != "common_L2":
            raise ValueError("Unknown signal penalty mode")
        sig_pen = ((power - signal_old) / n) ** 2
        norm_temp = float(deadband_l2(self.house_def["target_temp"], 0, self.house_def["target_temp"] + 1))
        norm_sig = float(deadband_l2(rp["norm_reg_sig"], 0, 0.75 * rp["norm_reg_sig"]))
        pen = deadband_l2(s["target"][e], s["deadband"][e], s["t_air"][e])
        mode = rp["temp_penalty_mode"]
        if mode == "individual_L2":
            tp = pen
        elif mode == "common_L2":
            acc = 0
            for v in pen:
                acc += v / n
            tp = np.full(n, acc)
        elif mode == "common_max":
            mx = 0
            for v in pen:
                if v > mx:
                    mx = v
            tp = np.full(n, mx, dtype=np.float64)
        elif mode == "mixture":
            pr = rp["temp_penalty_parameters"]["mixture"]
            c_l2, c_max = 0, 0
            for v in pen:
                c_l2 += v / n
                if v > c_max:
                    c_max = v
            a_i, a_c, a_m = pr["alpha_ind_L2"], pr["alpha_common_L2"], pr["alpha_common_max"]
            tp = (a_i * pen + a_c * c_l2 + a_m * c_max) / (a_i + a_c + a_m)
        else:
            raise ValueError("Unknown temperature penalty mode")
        return -1 * (rp["alpha_temp"] * tp / norm_temp + rp["alpha_sig"] * sig_pen / norm_sig)

    # -- normStateDict, utils.py:740-880 -----------------------------------------------
    def obs(self, msg_keep=None, comm=None):
        """Normalised observation tensor [E, N, F] (float64) of the *current* state.
        `msg_keep` [E, N, C] replays `np.random.rand() > comm_defect_prob` (1 = delivered)."""
        s = self.s
        sp = self.env_prop["state_properties"]
        mp = self.env_prop["message_properties"]
        hd, vd = self.house_def, self.hvac_def
        norm_sig = self.reward_prop["norm_reg_sig"]
        E, N = self.E, self.N
        cols = []
        cols.append((s["t_air"] - 20) / 5)
        cols.append((s["t_mass"] - 20) / 5)
        cols.append((s["target"] - 20) / 5)
        if sp["thermal"]:
            cols.append(np.repeat(((s["od_temp"] - 20) / 5)[:, None], N, 1))
        cols.append(s["deadband"])
        dts = [to_datetime(t) for t in s["t_epoch"]]
        if sp["day"]:
            day = np.array([d.timetuple().tm_yday for d in dts], dtype=np.float64)
            cols.append(np.repeat(np.sin(day * 2 * np.pi / 365)[:, None], N, 1))
            cols.append(np.repeat(np.cos(day * 2 * np.pi / 365)[:, None], N, 1))
        if sp["hour"]:
            hour = np.array([d.hour for d in dts], dtype=np.float64)
            cols.append(np.repeat(np.sin(hour * 2 * np.pi / 24)[:, None], N, 1))
            cols.append(np.repeat(np.cos(hour * 2 * np.pi / 24)[:, None], N, 1))
        if sp["solar_gain"]:
            cols.append(np.repeat((s["solar_gain"] / 1000)[:, None], N, 1))
        cols.append(s["cap"] / vd["cooling_capacity"])
        if sp["thermal"]:
            cols += [s["ua"] / hd["Ua"], s["cm"] / hd["Cm"], s["ca"] / hd["Ca"], s["hm"] / hd["Hm"]]
        if sp["hvac"]:
            cols += [s["cop"] / vd["COP"], s["latent"] / vd["latent_cooling_fraction"]]
        cols.append((s["on"] != 0).astype(np.float64))
        cols.append((s["lockout"] != 0).astype(np.float64))
        cols.append(s["sso"] / s["lockout_dur"])
        cols.append(s["lockout_dur"] / s["lockout_dur"])
        denom = norm_sig * self.cluster_prop["nb_agents"]
        cols.append(np.repeat((s["signal"] / denom)[:, None], N, 1))
        cols.append(np.repeat((s["cluster_power"] / denom)[:, None], N, 1))
        own = np.stack([np.asarray(c, dtype=np.float64) for c in cols], axis=-1)

        comm = self.comm if comm is None else np.asarray(comm, dtype=np.int64)
        if comm.ndim == 2:
            comm = np.broadcast_to(comm, (E,) + comm.shape)
        C = comm.shape[-1]
        if C == 0:
            return own
        ee = np.arange(E)[:, None, None]
        p_cur = np.where(s["on"] != 0, s["cap"] / s["cop"], 0.0)
        p_max = s["cap"] / s["cop"]
        fields = [
            (s["t_air"] - s["target"])[ee, comm] / 5,
            s["sso"][ee, comm] / s["lockout_dur"][:, :, None],
            p_cur[ee, comm] / norm_sig,
            p_max[ee, comm] / norm_sig,
        ]
        if mp["thermal"]:
            fields += [s["ua"][ee, comm] / hd["Ua"], s["cm"][ee, comm] / hd["Cm"],
                       s["ca"][ee, comm] / hd["Ca"], s["hm"][ee, comm] / hd["Hm"]]
        if mp["hvac"]:
            fields += [s["cop"][ee, comm] / vd["COP"],
                       s["latent"][ee, comm] / vd["latent_cooling_fraction"],
                       s["cap"][ee, comm] / vd["cooling_capacity"]]
        msg = np.stack(fields, axis=-1)  # [E, N, C, M]
        if msg_keep is not None:
            msg = msg * (np.asarray(msg_keep).reshape(E, N, C, 1) != 0)
        return np.concatenate([own, msg.reshape(E, N, -1)], axis=-1)

    # -- MADemandResponseEnv.step, env/MA_DemandResponse.py:174-210 ---------------------
    def step(self, actions, od_noise, sig_noise=None, interp_ids=None, msg_keep=None, comm=None):
        """actions [E, N] truthy; od_noise [E] replayed gauss draws; sig_noise [E] replayed
        perlin values; interp_ids [E, k] replayed random.choices; msg_keep [E, N, C];
        comm [E, N, C] replayed per-step neighbour table (random_sample mode).
        Returns obs [E,N,F], reward [E,N], cluster power [E], signal [E]."""
        s = self.s
        E, N, dt = self.E, self.N, self.dt
        actions = np.asarray(actions).reshape(E, N)
        od_noise = np.broadcast_to(np.asarray(od_noise, dtype=np.float64), (E,))
        sig_noise = np.zeros(E) if sig_noise is None else np.broadcast_to(np.asarray(sig_noise, dtype=np.float64), (E,))
        rewards = np.zeros((E, N))
        s["t_epoch"] = s["t_epoch"] + dt
        # HVAC state machine then thermal update with the OLD outdoor temperature
        on, lock, sso = hvac_step(s["on"], s["lockout"], s["sso"], s["lockout_dur"], actions, dt)
        s["on"], s["lockout"], s["sso"] = on.astype(np.int64), lock.astype(np.int64), sso
        for e in range(E):
            date_time = to_datetime(s["t_epoch"][e])
            if self.house_def["solar_gain_bool"]:
                gain = house_solar_gain(date_time, self.house_def["window_area"], self.house_def["shading_coeff"])
            else:
                gain = 0
            s["solar_gain"][e] = gain
            q_hvac = np.where(on[e], -1 * s["cap"][e] / (1 + s["latent"][e]), 0)
            q_a = q_hvac + gain
            s["t_air"][e], s["t_mass"][e] = etp_update(
                s["t_air"][e], s["t_mass"][e], s["od_temp"][e], q_a, s["ua"][e], s["ca"][e], s["hm"][e], s["cm"][e], dt
            )
            s["od_temp"][e] = od_temp_model(date_time, self.day_temp, self.night_temp, s["phase"][e], od_noise[e])
            # sequential fp64 sum in id order (env/MA_DemandResponse.py:1042-1050)
            p_each = np.where(on[e], s["cap"][e] / s["cop"][e], 0.0)
            power = 0
            for v in p_each:
                power += v
            s["cluster_power"][e] = power
            rewards[e] = self._rewards(e, power, s["signal"][e])  # OLD signal
            ids = None if interp_ids is None else np.asarray(interp_ids).reshape(E, -1)[e]
            self.grid_step(e, date_time, sig_noise[e], ids)
        obs = self.obs(msg_keep, comm)
        return obs, rewards, s["cluster_power"].copy(), s["signal"].copy()


def greedy_myopic_actions(t_air, target, power, lockout, signal):
    """TEST INFRASTRUCTURE -- agents/greedy_myopic_controller.py:29-49 restated for [E, N] arrays: houses sorted by
    -(house_temp - house_target_temp) ascending (the reference's pandas quicksort leaves the order of exact ties
    unspecified; here ties go by house id), then one sequential pass per cluster.  Note the reference's operator
    precedence: `A or (B and not lockout)` -- the lockout only guards the second clause.  `power` = capacity / COP,
    `signal` = obs["reg_signal"][0] per cluster.  Returns uint8 [E, N]."""
    t_air, target, power = np.asarray(t_air, np.float64), np.asarray(target, np.float64), np.asarray(power, np.float64)
    lockout, signal = np.asarray(lockout), np.asarray(signal, np.float64)
    e, n = t_air.shape
    out = np.zeros((e, n), np.uint8)
    for i in range(e):
        order = np.argsort(-(t_air[i] - target[i]), kind="stable")
        total = 0.0
        for k in order:
            pc = power[i, k]
            if (pc + total < signal[i]) or (abs(pc + total - signal[i]) < abs(total - signal[i]) and not lockout[i, k]):
                total += pc
                out[i, k] = 1
    return out
