"""TEST / BASELINE INFRASTRUCTURE ONLY -- per-object pure-Python port of the reference step.

The reference environment is single-threaded pure Python that loops over one object per house
(HVAC -> SingleHouse -> ClusterHouses -> PowerGrid) and builds one dict per agent per step plus
one dict per (agent, neighbour) message; `utils.normStateDict` then flattens every agent's dict
into the float vector the learners consume.  The reference itself cannot travel to the GPU box
(`/root/reference` is absent there), so this file restates that *execution structure* -- python
objects, python floats, per-house loops, per-message dicts, per-agent normalisation -- to serve as
the CPU arm of `bench.py` (`--impl reference` and `cpu_baseline`, kind = "port").  It is pinned to
the reference through `tests/test_oracle.py::test_scalar_port_*` (golden traces).

Reference citations: env/MA_DemandResponse.py:174-210 (step), :463-523 (HVAC), :664-738 (ETP),
:904-1003 (obs dict + messages), :1005-1081 (cluster step, OD temperature), :234-373 (reward),
:1236-1316 (power grid); utils.py:740-880 (normStateDict), :1266-1274 (deadbandL2).
Only the constant-base-power / replayed-noise path is ported (what the timed sample uses).
"""
from __future__ import annotations

import datetime as _dt
import math

EPOCH = _dt.datetime(1970, 1, 1)


class Hvac:
    __slots__ = ("cop", "cap", "latent", "lockout_duration", "turned_on", "lockout", "seconds_since_off", "dt",
                 "max_consumption")

    def __init__(self, cap, cop, latent, lockout_duration, dt, on=False, lockout=False, sso=None):
        self.cop, self.cap, self.latent, self.lockout_duration, self.dt = cop, cap, latent, lockout_duration, dt
        self.turned_on, self.lockout = on, lockout
        self.seconds_since_off = lockout_duration if sso is None else sso
        self.max_consumption = cap / cop

    def step(self, command):
        if self.turned_on == False:  # noqa: E712
            self.seconds_since_off += self.dt
        if self.turned_on or self.seconds_since_off >= self.lockout_duration:
            self.lockout = False
        else:
            self.lockout = True
        if self.lockout:
            self.turned_on = False
        else:
            self.turned_on = command
            if self.turned_on:
                self.seconds_since_off = 0
            elif self.seconds_since_off + self.dt < self.lockout_duration:
                self.lockout = True

    def get_q(self):
        return -1 * self.cap / (1 + self.latent) if self.turned_on else 0

    def power_consumption(self):
        return self.max_consumption if self.turned_on else 0


class House:
    __slots__ = ("ua", "cm", "ca", "hm", "target", "deadband", "t_air", "t_mass", "hvac", "solar_gain")

    def __init__(self, ua, cm, ca, hm, target, deadband, t_air, t_mass, hvac):
        self.ua, self.cm, self.ca, self.hm, self.target, self.deadband = ua, cm, ca, hm, target, deadband
        self.t_air, self.t_mass, self.hvac, self.solar_gain = t_air, t_mass, hvac, 0

    def update_temperature(self, od_temp, dt, gain):
        hm, ca, ua, cm = self.hm, self.ca, self.ua, self.cm
        od_k, ta_k, tm_k = od_temp + 273, self.t_air + 273, self.t_mass + 273
        self.solar_gain = gain
        q_a = self.hvac.get_q() + gain
        q_m = 0
        a = cm * ca / hm
        b = cm * (ua + hm) / hm + ca
        c = ua
        d = q_m + q_a + ua * od_k
        g = q_m / hm
        r1 = (-b + math.sqrt(b**2 - 4 * a * c)) / (2 * a)
        r2 = (-b - math.sqrt(b**2 - 4 * a * c)) / (2 * a)
        dta0 = hm * tm_k / ca - (ua + hm) * ta_k / ca + ua * od_k / ca + q_a / ca
        a1 = (r2 * ta_k - dta0 - r2 * d / c) / (r2 - r1)
        a2 = ta_k - d / c - a1
        a3 = r1 * ca / hm + (ua + hm) / hm
        a4 = r2 * ca / hm + (ua + hm) / hm
        new_ta = a1 * math.exp(r1 * dt) + a2 * math.exp(r2 * dt) + d / c
        new_tm = a1 * a3 * math.exp(r1 * dt) + a2 * a4 * math.exp(r2 * dt) + g + d / c
        self.t_air, self.t_mass = new_ta - 273, new_tm - 273

    def message(self, empty=False):
        if not empty:
            return {
                "current_temp_diff_to_target": self.t_air - self.target,
                "hvac_seconds_since_off": self.hvac.seconds_since_off,
                "hvac_curr_consumption": self.hvac.power_consumption(),
                "hvac_max_consumption": self.hvac.max_consumption,
                "hvac_lockout_duration": self.hvac.lockout_duration,
            }
        return {"current_temp_diff_to_target": 0, "hvac_seconds_since_off": 0, "hvac_curr_consumption": 0,
                "hvac_max_consumption": 0, "hvac_lockout_duration": 0}


def deadband_l2(target, deadband, value):
    if target + deadband / 2 < value:
        return (value - (target + deadband / 2)) ** 2
    if target - deadband / 2 > value:
        return ((target - deadband / 2) - value) ** 2
    return 0.0


class ScalarEnv:
    """One cluster.  `snap` is a single-env snapshot (see oracle/mdr_oracle.py), `config` the
    reference config dict.  Supports: `neighbours` messages, individual_L2 penalty, constant or interpolated base
    power (`interp` = oracle.mdr_oracle.PowerInterp; env/MA_DemandResponse.py:1195-1255), flat / sinusoidals /
    perlin (replayed noise) signal, solar gain off."""

    def __init__(self, config, snap, interp=None):
        ep = config["default_env_prop"]
        self.config = config
        self.dt = int(ep["time_step"])
        self.n = int(ep["cluster_prop"]["nb_agents"])
        hv = config["default_hvac_prop"]
        self.houses = {}
        for i in range(self.n):
            hvac = Hvac(float(snap["cap"][i]), hv["COP"], hv["latent_cooling_fraction"], int(snap["lockout_dur"][i]),
                        self.dt, bool(snap["on"][i]), bool(snap["lockout"][i]), int(snap["sso"][i]))
            self.houses[i] = House(float(snap["ua"][i]), float(snap["cm"][i]), float(snap["ca"][i]), float(snap["hm"][i]),
                                   float(snap["target"][i]), float(snap["deadband"][i]), float(snap["t_air"][i]),
                                   float(snap["t_mass"][i]), hvac)
        self.datetime = EPOCH + _dt.timedelta(seconds=int(snap["t_epoch"]))
        self.phase, self.od_temp = float(snap["phase"]), float(snap["od_temp"])
        self.ratio, self.max_power = float(snap["artificial_ratio"]), float(snap["max_power"])
        self.signal, self.cluster_power = float(snap["signal"]), float(snap["cluster_power"])
        tm = ep["cluster_prop"]["temp_parameters"][ep["cluster_prop"]["temp_mode"]]
        self.day_temp, self.night_temp = tm["day_temp"], tm["night_temp"]
        nb_comm = min(ep["cluster_prop"]["nb_agents_comm"], self.n - 1)
        self.comm = {}
        for i in range(self.n):
            before = [(i - nb_comm // 2 + k) % self.n for k in range(nb_comm // 2)]
            after = [(i + 1 + k) % self.n for k in range(int(math.ceil(nb_comm / 2)))]
            self.comm[i] = before + after
        self.gp = ep["power_grid_prop"]
        self.rp = ep["reward_prop"]
        self.house_def = config["default_house_prop"]
        self.interp = interp
        self.base_power = float(snap["base_power"]) if "base_power" in snap else 0.0
        self.time_since_interp = int(snap["time_since_interp"]) if "time_since_interp" in snap else 0

    def _interpolate_power(self, ids):
        """PowerGrid.interpolatePower :1195-1234 (solar gain off: hour = date = 0); `ids` replays random.choices."""
        ip = self.gp["base_power_parameters"]["interpolation"]
        hd = self.house_def
        if self.n <= ip["interp_nb_agents"]:
            ids, factor = list(range(self.n)), 1
        else:
            factor = float(self.n) / ip["interp_nb_agents"]
        base = 0
        for i in ids:
            house = self.houses[int(i)]
            point = {"date": 0.0, "hour": 0.0, "Ua_ratio": house.ua / hd["Ua"], "Cm_ratio": house.cm / hd["Cm"],
                     "Ca_ratio": house.ca / hd["Ca"], "Hm_ratio": house.hm / hd["Hm"],
                     "air_temp": house.t_air - house.target, "mass_temp": house.t_mass - house.target,
                     "OD_temp": self.od_temp - house.target, "HVAC_power": house.hvac.cap}
            base += self.interp.fast(self.interp.clip(point))
        return base * factor

    def _obs_dict(self):
        obs = {}
        for i, house in self.houses.items():
            hvac = house.hvac
            d = {
                "OD_temp": self.od_temp, "datetime": self.datetime, "house_temp": house.t_air,
                "house_mass_temp": house.t_mass, "hvac_turned_on": hvac.turned_on,
                "hvac_seconds_since_off": hvac.seconds_since_off, "hvac_lockout": hvac.lockout,
                "house_target_temp": house.target, "house_deadband": house.deadband, "house_Ua": house.ua,
                "house_Cm": house.cm, "house_Ca": house.ca, "house_Hm": house.hm, "house_solar_gain": house.solar_gain,
                "hvac_COP": hvac.cop, "hvac_cooling_capacity": hvac.cap, "hvac_latent_cooling_fraction": hvac.latent,
                "hvac_lockout_duration": hvac.lockout_duration,
            }
            d["message"] = [self.houses[j].message() for j in self.comm[i]]
            obs[i] = d
        return obs

    def step(self, action_dict, od_noise, sig_noise=0.0, interp_ids=None):
        self.datetime += _dt.timedelta(seconds=self.dt)
        for i, house in self.houses.items():
            house.hvac.step(action_dict[i])
            house.update_temperature(self.od_temp, self.dt, 0)
        amplitude, bias = (self.day_temp - self.night_temp) / 2, (self.day_temp + self.night_temp) / 2
        time_day = self.datetime.hour + self.datetime.minute / 60.0
        self.od_temp = amplitude * math.sin(2 * math.pi * (time_day + (-6 + self.phase)) / 24) + bias + od_noise
        obs = self._obs_dict()
        power = 0
        for house in self.houses.values():
            power += house.hvac.power_consumption()
        self.cluster_power = power
        # rewards with the OLD signal
        sig_pen = ((power - self.signal) / self.n) ** 2
        norm_temp = deadband_l2(self.house_def["target_temp"], 0, self.house_def["target_temp"] + 1)
        norm_sig = deadband_l2(self.rp["norm_reg_sig"], 0, 0.75 * self.rp["norm_reg_sig"])
        rewards = {}
        for i, house in self.houses.items():
            pen = deadband_l2(house.target, house.deadband, house.t_air)
            rewards[i] = -1 * (self.rp["alpha_temp"] * pen / norm_temp + self.rp["alpha_sig"] * sig_pen / norm_sig)
        # power grid
        if self.gp["base_power_mode"] == "interpolation":      # PowerGrid.step :1250-1255
            self.time_since_interp += self.dt
            if self.time_since_interp >= self.gp["base_power_parameters"]["interpolation"]["interp_update_period"]:
                self.base_power = self._interpolate_power(interp_ids)
                self.time_since_interp = 0
            base = self.base_power
        else:
            base = self.gp["base_power_parameters"]["constant"]["avg_power_per_hvac"] * self.n
        mode = self.gp["signal_mode"]
        params = self.gp["signal_parameters"][mode]
        if mode == "flat":
            sig = base
        elif mode == "sinusoidals":
            time_sec = self.datetime.hour * 3600 + self.datetime.minute * 60 + self.datetime.second
            sig = base
            for ratio, period in zip(params["amplitude_ratios"], params["periods"]):
                sig += base * ratio * math.sin(2 * math.pi * time_sec / period)
        else:
            sig = max(0, base + (base * params["amplitude_ratios"] * sig_noise))
        self.signal = min(sig * self.ratio, self.max_power)
        for d in obs.values():
            d["reg_signal"] = self.signal
            d["cluster_hvac_power"] = power
        return obs, rewards, {i: False for i in obs}, {"cluster_hvac_power": power}

    def norm_state(self, s):
        """utils.normStateDict (default flags): dict -> flat list of 11 + 4*C floats."""
        norm = self.rp["norm_reg_sig"]
        lock = s["hvac_lockout_duration"]
        out = [(s["house_temp"] - 20) / 5, (s["house_mass_temp"] - 20) / 5, (s["house_target_temp"] - 20) / 5,
               s["house_deadband"], s["hvac_cooling_capacity"] / self.config["default_hvac_prop"]["cooling_capacity"],
               1 if s["hvac_turned_on"] else 0, 1 if s["hvac_lockout"] else 0, s["hvac_seconds_since_off"] / lock,
               lock / lock, s["reg_signal"] / (norm * self.n), s["cluster_hvac_power"] / (norm * self.n)]
        for m in s["message"]:
            out += [m["current_temp_diff_to_target"] / 5, m["hvac_seconds_since_off"] / lock,
                    m["hvac_curr_consumption"] / norm, m["hvac_max_consumption"] / norm]
        return out


def timed_rollout(config, snap, steps, seed=0, interp=None, normalise=True):
    """Steps one cluster `steps` times with a bang-bang policy (agents/bangbang_controllers.py:41-61)
    and normalises every agent's observation, like a learner's rollout loop (train_ppo.py:62-116).
    `interp` (PowerInterp) enables the interpolated base power with its refresh every interp_update_period.
    Returns (house_steps, seconds)."""
    import random
    import time

    rng = random.Random(seed)
    env = ScalarEnv(config, snap, interp)
    nb = config["default_env_prop"]["power_grid_prop"]["base_power_parameters"]["interpolation"]["interp_nb_agents"]
    obs = env._obs_dict()
    for d in obs.values():
        d["reg_signal"], d["cluster_hvac_power"] = env.signal, env.cluster_power
    t0 = time.perf_counter()
    for _ in range(steps):
        act = {i: obs[i]["house_temp"] > obs[i]["house_target_temp"] for i in obs}
        ids = rng.choices(range(env.n), k=nb) if (interp is not None and env.n > nb) else None
        obs, rew, _, _ = env.step(act, rng.gauss(0, 0.5), rng.uniform(-0.3, 0.3), ids)
        vecs = [env.norm_state(obs[i]) for i in obs] if normalise else [None] * env.n
    dt = time.perf_counter() - t0
    assert len(vecs) == env.n
    return env.n * steps, dt
