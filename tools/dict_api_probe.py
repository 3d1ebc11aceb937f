"""Speed of the drop-in dict API (MADemandResponseEnv.step with python dicts in and out) -- BASELINE config 0:
bang-bang control of the default 50-house... cluster sizes, constant base power (the table blob is not shipped)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdr_b200

for n in (int(a) for a in (sys.argv[1:] or ["50", "1000"])):
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = n
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "constant"
    cfg["default_house_prop"]["solar_gain_bool"] = False
    random.seed(1)
    env = mdr_b200.MADemandResponseEnv(cfg, precision="fp64")
    obs = env.reset()
    def bangbang(o):  # agents/bangbang_controllers.py:50-61
        return {i: o[i]["house_temp"] > o[i]["house_target_temp"] for i in o}
    for _ in range(20):
        obs, rew, done, info = env.step(bangbang(obs))
    steps = 300 if n <= 100 else 60
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, rew, done, info = env.step(bangbang(obs))
    dt = time.perf_counter() - t0
    print("dict API, N=%d: %.3f ms per step (%.3g house-steps/s incl. the python controller)" % (n, dt / steps * 1e3, n * steps / dt))
