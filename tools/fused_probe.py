"""Where does the fused multi-step kernel spend its time? (signal mode x metrics on/off, c3 shape)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mdr_b200

E, N, K = 10000, 100, 75
for signal, interp in (("perlin", False), ("sinusoidals", False), ("flat", False), ("perlin", True)):
    for metrics in (False, True):
        cfg = mdr_b200.make_default_config()
        ep = cfg["default_env_prop"]
        ep["cluster_prop"]["nb_agents"] = N
        ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
        ep["power_grid_prop"]["signal_mode"] = signal
        cfg["default_house_prop"]["solar_gain_bool"] = False
        flat = mdr_b200.FlatConfig(cfg)
        pop = mdr_b200.synthetic_population(flat, E, seed=1)
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=1, action_source="bangbang", with_obs=False,
                                            interp_table=mdr_b200.synthetic_interp_table() if interp else None)
        env.reset_tensor()
        if metrics:
            env.enable_metrics()
        for _ in range(3):
            env.run(K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            env.run(K)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * K)
        env.set_launch_options(no_fused=True)
        if metrics:
            env.disable_metrics()
        for _ in range(20):
            env.step_tensor(None)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(300):
            env.step_tensor(None)
        e1.record(); torch.cuda.synchronize()
        env.set_launch_options(no_fused=False)
        per_step = e0.elapsed_time(e1) * 1e3 / 300
        print("signal %-12s interp %-5s metrics %-5s: fused %.2f us per env step (%.3g house-steps/s); one launch per step %.2f us"
              % (signal, interp, metrics, us, E * N / us * 1e6, per_step))
