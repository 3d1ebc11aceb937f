"""e2e of mdr_step_host on c4 by slice / thread count of the MdrHostCtx pipeline (host-bound: scales with the expansion threads)."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, mdr_b200
w = bench.WORKLOADS["c4"]; cfg = bench.workload_config(w); flat = mdr_b200.FlatConfig(cfg)
E, N = w["envs"], w["houses"]
pop = mdr_b200.synthetic_population(flat, E, seed=1234)
env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", device="cuda:0", seed=1234, interp_table=mdr_b200.synthetic_interp_table())
env.reset_tensor(); env.stagger_interp_clock(seed=77)
acts = [(np.random.default_rng(i).random((E, N)) < 0.5).astype(np.uint8) for i in range(4)]
for slices in (4, 8, 12, 16):
    for threads in (8, 12, 15):
        env.host_pipeline(True, n_threads=threads, n_slices=slices)
        for i in range(3): env.step_host(acts[i & 3])
        t0 = time.perf_counter()
        for i in range(20): env.step_host(acts[i & 3])
        dt = (time.perf_counter() - t0) / 20
        print("slices %2d threads %2d: %.2f ms/step, %.3g house-steps/s" % (slices, threads, dt * 1e3, E * N / dt), env.host_pipeline_info(), flush=True)
