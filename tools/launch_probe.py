"""Where does the time between launches go?  Per-step device time (CUDA events) of a workload issued as (a) one python
call per step, (b) one C call for 50 steps (MDR_FLAG_NO_FUSED), (c) without programmatic dependent launch; plus the
host time of a python call.  Usage: python tools/launch_probe.py c3big [c1 ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mdr_b200

for name in sys.argv[1:] or ["c3big"]:
    w = bench.WORKLOADS[name]
    cfg = bench.workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    E, N = w["envs"], w["houses"]
    pop = mdr_b200.synthetic_population(flat, E, seed=1)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=w.get("precision", "fp32"), seed=1, interp_table=table,
                                        action_source=w["action_source"], with_obs=w["obs"])
    env.reset_tensor()
    env.set_launch_options(no_fused=True)
    act = (torch.rand(E, N, device="cuda") < 0.5).to(torch.uint8) if w["action_source"] == "array" else None
    def run(calls, n_steps):
        for _ in range(3):
            env.step_tensor(act, n_steps=n_steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(calls):
            env.step_tensor(act, n_steps=n_steps)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (calls * n_steps), t_issue * 1e6 / calls
    print(name, env.launch_geometry())
    for label, kw in (("PDL", dict(no_pdl=False)), ("no PDL", dict(no_pdl=True))):
        env.set_launch_options(**kw)
        a, ha = run(300, 1)
        b, hb = run(8, 50)
        print("  %-7s one python call per step: %.2f us/step (host %.1f us per call);  50 steps per C call: %.2f us/step (host %.1f us per call)"
              % (label, a, ha, b, hb))
