"""Host time per step_tensor call vs device time per launch (is the launch loop host-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mdr_b200

for name in sys.argv[1:] or ["c4", "c2", "c3"]:
    w = bench.WORKLOADS[name]
    cfg = bench.workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    E, N = w["envs"], w["houses"]
    pop = mdr_b200.synthetic_population(flat, E, seed=1234)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", device="cuda:0", seed=1234, interp_table=table,
                                        action_source=w["action_source"], with_obs=w["obs"])
    env.reset_tensor()
    act = (torch.rand(E, N, device="cuda:0") < 0.5).to(torch.uint8)
    a = act if w["action_source"] == "array" else None
    for _ in range(50):
        env.step_tensor(a)
    torch.cuda.synchronize()
    K = 600
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        env.step_tensor(a)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    host_us = (t1 - t0) / K * 1e6
    dev_us = e0.elapsed_time(e1) / K * 1e3
    e0.record()
    env.step_tensor(a, n_steps=K)
    e1.record(); torch.cuda.synchronize()
    c_us = e0.elapsed_time(e1) / K * 1e3
    print("%s: python loop host %.1f us/call, device %.1f us/step; one C call with n_steps=%d: %.1f us/step"
          % (name, host_us, dev_us, K, c_us))
