"""Helpers shared by the CPU (oracle) and GPU (parity) tests: load a golden trace."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# monteCarlo/interp_parameters_dict.json + interp_dict_keys.csv of the reference (grid axes)
INTERP_KEYS = ["Ua_ratio", "Cm_ratio", "Ca_ratio", "Hm_ratio", "air_temp", "mass_temp", "OD_temp",
               "HVAC_power", "hour", "date"]
INTERP_GRID = {
    "Ua_ratio": [0.9, 1, 1.1], "Cm_ratio": [0.9, 1, 1.1], "Ca_ratio": [0.9, 1, 1.1], "Hm_ratio": [0.9, 1, 1.1],
    "air_temp": [-4, -2, -1, -0.3, 0, 0.3, 1, 2, 4], "mass_temp": [-4, -2, 0, 2, 4],
    "OD_temp": [1, 3, 5, 7, 9, 11, 13, 15], "HVAC_power": [10000, 15000],
    "hour": [0.0, 10800.0, 21600.0, 25200.0, 27000.0, 39600.0, 46800.0, 57600.0, 61200.0, 63000.0, 75600.0, 86399.0],
    "date": [0, 79, 171, 263, 354, 364],
}
TABLE_SIZE = 4199040
_table = None


def synthetic_table():
    global _table
    if _table is None:
        _table = np.random.default_rng(0).uniform(0, 6000, TABLE_SIZE)
    return _table


def names():
    """trace fixtures (oracle/make_golden.py); mc_* files are not traces: table samples (oracle/make_mc_golden.py), controller decisions (oracle/make_greedy_golden.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith("mc_"))


def fix_config(cfg):
    """JSON turned the int keys of cooling_capacity_list into strings."""
    for key in ("noise_hvac_prop", "noise_hvac_prop_test"):
        for mode in cfg[key]["noise_parameters"].values():
            if "cooling_capacity_list" in mode:
                mode["cooling_capacity_list"] = {int(k): v for k, v in mode["cooling_capacity_list"].items()}
    return cfg


class Golden:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.config = fix_config(json.loads(str(self.z["config_json"])))
        self.snap = {k[5:]: v for k, v in self.z.items() if k.startswith("snap_")}
        self.steps = int(self.z["steps"])
        self.seed = int(self.z["seed"])
        self.n = self.z["actions"].shape[1]
        self.check_steps = [int(i) for i in self.z["check_steps"]]
        self.obs_steps = [int(i) for i in self.z["obs_steps"]]
        self.comm = self.z.get("comm")
        self.comm_t = self.z.get("comm_t")
        self.uses_interp = self.config["default_env_prop"]["power_grid_prop"]["base_power_mode"] == "interpolation"

    def __getattr__(self, k):
        try:
            return self.__dict__["z"][k]
        except KeyError:
            raise AttributeError(k)

    def batched_snap(self):
        """snapshot with a leading env axis of 1"""
        out = {}
        for k, v in self.snap.items():
            out[k] = np.asarray(v)[None] if np.ndim(v) >= 1 else np.asarray(v).reshape(1)
        return out
