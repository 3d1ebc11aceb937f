"""TEST INFRASTRUCTURE ONLY -- golden entries of the Monte-Carlo base-power table, produced by the
UNMODIFIED reference function `eval_parameters_bangbang_average_consumption`
(monteCarlo/monteCarlo.py:133-201).  The module itself cannot be imported (it parses argv and runs the
whole 4.2M-combination sweep at import time), so the function's source is read from the reference
file at run time and executed in a namespace holding the reference's own globals.  Nothing of it is
written into this repository; only the sampled (index, value) pairs are saved.

    TZ=UTC python oracle/make_mc_golden.py
"""
import ast
import copy
import datetime
import os
import sys
from datetime import date, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stubs  # noqa: E402

KEYS = ["Ua_ratio", "Cm_ratio", "Ca_ratio", "Hm_ratio", "air_temp", "mass_temp", "OD_temp", "HVAC_power", "hour", "date"]


def main():
    import json
    Env, norm, cfg, ref_utils = ref_stubs.import_reference()
    from agents.bangbang_controllers import BangBangController  # type: ignore
    path = os.path.join(ref_stubs.REFERENCE_ROOT, "monteCarlo", "monteCarlo.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "eval_parameters_bangbang_average_consumption"][0]
    ns = dict(copy=copy, datetime=datetime, timedelta=timedelta, config_dict=cfg, MADemandResponseEnv=Env,
              BangBangController=BangBangController, get_actions=ref_utils.get_actions, d0=date(2021, 1, 1),
              NB_TIME_STEPS_BY_SIM=75, NB_TIME_STEPS_AVG=10)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    evaluate = ns["eval_parameters_bangbang_average_consumption"]
    grid = json.load(open(os.path.join(ref_stubs.REFERENCE_ROOT, "monteCarlo", "interp_parameters_dict.json")))
    shape = [len(grid[k]) for k in KEYS]
    rng = np.random.default_rng(42)
    idx = np.stack([rng.integers(0, n, 96) for n in shape], axis=1)
    idx[0] = 0
    idx[1] = [n - 1 for n in shape]
    idx[2] = [1, 1, 1, 1, 4, 2, 7, 1, 5, 2]   # a midday, midsummer entry: solar gain on
    vals = np.array([evaluate(*[grid[k][i] for k, i in zip(KEYS, row)]) for row in idx], dtype=np.float64)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "mc_table_samples.npz")
    np.savez_compressed(out, idx=idx.astype(np.int16), values=vals)
    print("saved", out, "min/mean/max", vals.min(), vals.mean(), vals.max(), "nonzero", int((vals > 0).sum()))


if __name__ == "__main__":
    main()
