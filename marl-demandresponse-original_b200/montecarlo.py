"""Regenerates the Monte-Carlo base-power table on the GPU.

The reference's `PowerInterpolator` (monteCarlo/interpolation.py:18-142) reads
`monteCarlo/mergedGridSearchResultFinal.npy`, a 4 199 040-entry table that is *not shipped*
(`.MISSING_LARGE_BLOBS`).  Its provenance is monteCarlo/monteCarlo.py:133-201: for every combination
of the grid in `interp_parameters_dict.json`, a fresh 1-house environment (no noise, constant outdoor
temperature = target + OD_temp, lockout 1 s, constant base power) is driven by a bang-bang controller
for 75 steps, and the entry is the mean of the last 10 running averages of the HVAC power
(:193-197).  That is "the env step path, 3.1e8 times": here every combination is one single-house
cluster of a `VecDemandResponseEnv` stepped with the on-device bang-bang action source.

`regenerate_table()` returns the flat fp64 array in the C order of `interp_dict_keys.csv`, i.e. exactly
what `np.load(path_datafile)` would have returned.
"""
import copy
import datetime as _dt

import numpy as np

from .config_flatten import FlatConfig, epoch_seconds
from .default_config import INTERP_GRID, INTERP_KEYS, default_config

NB_TIME_STEPS_BY_SIM = 75  # monteCarlo/monteCarlo.py:23
NB_TIME_STEPS_AVG = 10     # monteCarlo/monteCarlo.py:24


def grid_shape():
    return tuple(len(INTERP_GRID[k]) for k in INTERP_KEYS)


def mc_config(config, od_offset):
    """The per-combination config edits of monteCarlo.py:138-171 that are shared by a whole batch."""
    cfg = copy.deepcopy(config)
    cfg["noise_house_prop"]["noise_mode"] = "no_noise"
    cfg["noise_hvac_prop"]["noise_mode"] = "no_noise"
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = 1
    ep["start_datetime_mode"] = "fixed"
    cfg["default_hvac_prop"]["lockout_duration"] = 1
    ep["cluster_prop"]["temp_mode"] = "constant"
    target = cfg["default_house_prop"]["target_temp"]
    ep["cluster_prop"]["temp_parameters"]["constant"]["day_temp"] = target + od_offset
    ep["cluster_prop"]["temp_parameters"]["constant"]["night_temp"] = target + od_offset
    ep["power_grid_prop"]["base_power_mode"] = "constant"
    ep["power_grid_prop"]["signal_mode"] = "flat"  # the bang-bang controller never looks at the signal
    return cfg


def mc_population(config, idx):
    """Population of single-house clusters for the grid multi-indices `idx` [E, 10] (all with the same
    OD_temp index)."""
    hd, vd = config["default_house_prop"], config["default_hvac_prop"]
    g = {k: np.asarray(INTERP_GRID[k], dtype=np.float64) for k in INTERP_KEYS}
    col = {k: idx[:, i] for i, k in enumerate(INTERP_KEYS)}
    e = idx.shape[0]
    target = float(hd["target_temp"])
    od = target + g["OD_temp"][col["OD_temp"]]
    d0 = _dt.datetime(2021, 1, 1)
    hour = g["hour"][col["hour"]]
    # monteCarlo.py:140-142: int(hour // 3600), int(hour % 3600 // 60), int(hour % 60)
    sod = (hour // 3600).astype(np.int64) * 3600 + (hour % 3600 // 60).astype(np.int64) * 60 + (hour % 60).astype(np.int64)
    t_epoch = epoch_seconds(d0) + g["date"][col["date"]].astype(np.int64) * 86400 + sod
    one = lambda v: np.asarray(v, dtype=np.float64).reshape(e, 1)
    pop = {
        "ua": one(hd["Ua"] * g["Ua_ratio"][col["Ua_ratio"]]), "cm": one(hd["Cm"] * g["Cm_ratio"][col["Cm_ratio"]]),
        "ca": one(hd["Ca"] * g["Ca_ratio"][col["Ca_ratio"]]), "hm": one(hd["Hm"] * g["Hm_ratio"][col["Hm_ratio"]]),
        "cap": one(g["HVAC_power"][col["HVAC_power"]]),
        "target": one(np.full(e, target)), "deadband": one(np.full(e, float(hd["deadband"]))),
        "t_air": one(target + g["air_temp"][col["air_temp"]]), "t_mass": one(target + g["mass_temp"][col["mass_temp"]]),
        "lockout_dur": np.ones((e, 1), np.int64), "sso": np.ones((e, 1), np.int64),
        "on": np.zeros((e, 1), np.int64), "lockout": np.zeros((e, 1), np.int64),
        "t_epoch": t_epoch, "phase": np.zeros(e), "od_temp": od, "artificial_ratio": np.ones(e),
        "max_power": g["HVAC_power"][col["HVAC_power"]] / float(vd["COP"]),
        "base_power": np.zeros(e), "signal": np.zeros(e), "cluster_power": np.zeros(e), "solar_gain": np.zeros(e),
        "time_since_interp": np.zeros(e, np.int64), "perlin_seed": np.zeros(e),
    }
    return pop


def regenerate_entries(idx, config=None, device=None, precision="fp64", max_envs=1 << 20):
    """Table entries for the grid multi-indices `idx` [M, 10] -> float64 [M]."""
    from .vec_env import VecDemandResponseEnv
    import torch

    config = default_config() if config is None else config
    idx = np.asarray(idx, dtype=np.int64).reshape(-1, len(INTERP_KEYS))
    out = np.zeros(idx.shape[0], dtype=np.float64)
    od_col = INTERP_KEYS.index("OD_temp")
    for o in np.unique(idx[:, od_col]):
        rows = np.nonzero(idx[:, od_col] == o)[0]
        cfg = mc_config(config, float(INTERP_GRID["OD_temp"][int(o)]))
        for lo in range(0, len(rows), max_envs):
            sel = rows[lo:lo + max_envs]
            pop = mc_population(cfg, idx[sel])
            env = VecDemandResponseEnv(cfg, pop, precision=precision, device=device, action_source="bangbang",
                                       with_obs=False)
            env.reset_tensor()
            total = torch.zeros(len(sel), dtype=torch.float64, device=env.device)
            avg = torch.zeros_like(total)
            for i in range(NB_TIME_STEPS_BY_SIM):  # monteCarlo.py:193-199
                env.run(1)
                total += env.env["cluster_power"]
                if i >= NB_TIME_STEPS_BY_SIM - NB_TIME_STEPS_AVG:
                    avg += total / ((i + 1) * NB_TIME_STEPS_AVG)
            out[sel] = avg.cpu().numpy()
    return out


def regenerate_table(config=None, device=None, precision="fp64"):
    """The whole table (4 199 040 entries, ~3.1e8 house-steps), flat, C order of interp_dict_keys.csv."""
    shape = grid_shape()
    idx = np.stack(np.unravel_index(np.arange(int(np.prod(shape))), shape), axis=1)
    return regenerate_entries(idx, config, device, precision)
