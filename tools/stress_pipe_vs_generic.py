"""Long-run stress: the pipelined kernel (with PDL, back-to-back launches) against the generic kernel on the same
inputs for thousands of steps at full size; catches rare ordering bugs (mbarrier ring, cp.async stages, deferred
refresh, programmatic dependent launch) that a 30-step parity test would miss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mdr_b200

steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
schedules = sys.argv[3].split(",") if len(sys.argv) > 3 else ["claimed", "strided"]
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["c4", "c2"]):
  for schedule in schedules:
    w = bench.WORKLOADS[name]
    cfg = bench.workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    E, N = w["envs"], w["houses"]
    pop = mdr_b200.synthetic_population(flat, E, seed=77)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    mk = lambda: mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", device="cuda:0", seed=77, interp_table=table,
                                               action_source=w["action_source"], with_obs=w["obs"])
    a, b = mk(), mk()
    a.set_launch_options(no_fused=True, static_tiles=(schedule == "strided"))
    b.set_launch_options(no_fused=True, no_pipeline=True)   # the generic kernel
    a.reset_tensor(); b.reset_tensor()
    if w["interp"]:  # refresh clocks staggered like in bench.py: every launch has due tiles
        a.stagger_interp_clock(seed=3)
        b.time_since_interp.copy_(a.time_since_interp)
    g = torch.Generator(device="cuda").manual_seed(5)
    ring = [(torch.rand(E, N, device="cuda", generator=g) < 0.5).to(torch.uint8) for _ in range(8)]
    use = w["action_source"] == "array"
    worst = 0.0
    for t in range(steps):
        act = ring[t & 7] if use else None
        oa = a.step_tensor(act)
        if t % 500 == 499 or t == steps - 1:
            # catch b up with the generic kernel and compare
            while b.step_index < a.step_index:
                ob = b.step_tensor(ring[b.step_index & 7] if use else None)
            torch.cuda.synchronize()
            assert torch.equal(a.hvac, b.hvac), (name, t, "hvac")
            assert torch.equal(a.t_epoch, b.t_epoch) and torch.equal(a.time_since_interp, b.time_since_interp), (name, t)
            assert torch.equal(oa[2], ob[2]), (name, t, "power")
            d = float((a.temps - b.temps).abs().max())
            ds = float(((oa[3] - ob[3]).abs() / ob[3].abs().clamp(min=1.0)).max())
            worst = max(worst, d)
            assert d < 2e-3 and ds < 1e-4, (name, t, d, ds)
            if oa[0] is not None:
                assert torch.isfinite(oa[0]).all()
                torch.testing.assert_close(oa[0], ob[0], rtol=1e-4, atol=2e-3)
    print("%s [%s tiles]: %d steps, pipelined == generic (integer state bit-exact, max |dT| %.2e)" % (name, schedule, steps, worst), flush=True)
