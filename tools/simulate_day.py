"""BASELINE config 3, literally: 1M houses (10,000 clusters x 100) with heterogeneous thermal parameters and lockout,
ONE simulated day at 4 s steps (21,600 steps) on one GPU, bang-bang control on the device, deploy metrics on the
device (main-deploy.py:102-209).  Prints the wall time of the day, house-steps/s and the reference's summary figures."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mdr_b200

E, N, STEPS = 10000, 100, 21600
interp = len(sys.argv) > 1 and sys.argv[1] == "interp"
cfg = mdr_b200.make_default_config()
ep = cfg["default_env_prop"]
ep["cluster_prop"]["nb_agents"] = N
ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
cfg["default_house_prop"]["solar_gain_bool"] = False
flat = mdr_b200.FlatConfig(cfg)
pop = mdr_b200.synthetic_population(flat, E, seed=2021)
env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=2021, action_source="bangbang", with_obs=False,
                                    interp_table=mdr_b200.synthetic_interp_table() if interp else None)
env.reset_tensor()
env.enable_metrics()
env.run(75)  # warm-up (also a fused launch)
env.enable_metrics(reset=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
chunk = 2700  # 3 simulated hours per launch
for _ in range(STEPS // chunk):
    env.run(chunk)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
dev_s = e0.elapsed_time(e1) * 1e-3
summ = env.metrics_summary()
print("base power: %s; %d clusters x %d houses, %d steps (one day at %d s): device time %.3f s (wall %.3f s) -> %.3g house-steps/s"
      % ("interpolation" if interp else "constant", E, N, STEPS, flat.time_step, dev_s, wall, E * N * STEPS / dev_s))
for k in ("rmse_signal_per_agent", "rmse_temp", "rms_max_error_temp", "mean_temp_error", "mean_signal", "mean_consumption", "mean_od_temp"):
    v = summ[k]
    print("  %-24s mean over clusters %.4g  (min %.4g, max %.4g)" % (k, float(v.mean()), float(v.min()), float(v.max())))
assert torch.isfinite(env.temps).all() and float(summ["steps"].min()) == STEPS
