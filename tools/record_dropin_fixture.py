"""GPU-box half of the drop-in proof: runs the drop-in `MADemandResponseEnv` (dict API) on BASELINE
config 0's shape -- main-deploy.py's default: 50 houses, 4 s steps, bang-bang -- and records what the reference's
*unmodified* controllers will be fed in the build container (tests/test_dropin_reference.py): every `obs_dict`, the
actions a restated bang-bang rule chose on it, the rewards, and the device-normalised observation matrix.
Usage: python tools/record_dropin_fixture.py [out.pkl]   (default gpurun_out/dropin_c0.pkl)"""
import os
import pickle
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mdr_b200

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dropin_c0.pkl"
cfg = mdr_b200.make_default_config()
ep = cfg["default_env_prop"]
ep["cluster_prop"]["nb_agents"] = 50                      # cli.py:627-632 (main-deploy default)
ep["power_grid_prop"]["base_power_mode"] = "constant"     # the interpolation table is a missing blob of the reference
cfg["default_house_prop"]["solar_gain_bool"] = False      # both reference CLIs force it off (utils.py:437)
random.seed(1)
np.random.seed(1)
env = mdr_b200.MADemandResponseEnv(cfg)
obs = env.reset()
steps = 24
rec = dict(config=cfg, obs=[], actions=[], rewards=[], obs_tensor=[], power=[])
for t in range(steps + 1):
    rec["obs"].append(obs)
    rec["obs_tensor"].append(env.obs_tensor().cpu().numpy().copy())
    if t == steps:
        break
    # agents/bangbang_controllers.py:41-61 restated (the real class is applied to these dicts in the build container)
    act = {k: bool(obs[k]["house_temp"] > obs[k]["house_target_temp"]) for k in obs}
    rec["actions"].append(act)
    obs, rew, done, info = env.step(act)
    rec["rewards"].append(rew)
    rec["power"].append(info["cluster_hvac_power"])
    assert not any(done.values())
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
with open(out, "wb") as f:
    pickle.dump(rec, f, protocol=4)
print("recorded %d steps x %d houses -> %s (%d KB)" % (steps, env.nb_agents, out, os.path.getsize(out) // 1024))
