"""Small workloads that drive every kernel family of libmdr_b200.so once, for `compute-sanitizer`
(tools/sanitize.sh runs it under memcheck / racecheck / synccheck / initcheck with `--kernel-name kns=mdr`, so only this
library's kernels are instrumented).  Sizes are tiny (instrumented kernels run 10-100x slower) but chosen so that the
pipelined kernel's steady-state loop is entered: the persistent grid is capped at 2 CTAs x >= 12 tiles, with staggered
interpolation refreshes, ragged last tile and the mbarrier ring wrapping."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mdr_b200

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["pipe", "generic", "fused", "cluster", "populate"]


def config(n, interp=False, signal="perlin", **kw):
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    ep["power_grid_prop"]["signal_mode"] = signal
    cfg["default_house_prop"]["solar_gain_bool"] = bool(kw.get("solar", False))
    ep["cluster_prop"]["comm_defect_prob"] = kw.get("defect", 0.0)
    ep["reward_prop"]["temp_penalty_mode"] = kw.get("penalty", "individual_L2")
    ep["cluster_prop"]["agents_comm_mode"] = kw.get("comm", "neighbours")
    for k in kw.get("state", ()):
        ep["state_properties"][k] = True
    return cfg, mdr_b200.FlatConfig(cfg)


table = mdr_b200.synthetic_interp_table()
gen = torch.Generator(device="cuda").manual_seed(1)


def run(name, env, steps, n_steps=1, use_actions=True):
    e, n = env.n_envs, env.n_houses
    env.reset_tensor()
    if env.flat.base_power_mode:
        env.stagger_interp_clock(seed=2)
    for t in range(steps):
        act = (torch.rand(e, n, device="cuda", generator=gen) < 0.5).to(torch.uint8) if use_actions else None
        env.step_tensor(act, n_steps=n_steps)
    torch.cuda.synchronize()
    print("sanitize_driver: %-28s %s, %d x %d, %d steps -> ok" % (name, env.launch_geometry()["kernel"].split(" ")[0], e, n, steps * n_steps),
          flush=True)


if "pipe" in which:
    for n_envs, n, interp, nb in ((49, 100, True, 10), (95, 50, False, 10), (85, 30, True, 4)):
        cfg, flat = config(n, interp)
        cfg["default_env_prop"]["cluster_prop"]["nb_agents_comm"] = nb
        pop = mdr_b200.synthetic_population(mdr_b200.FlatConfig(cfg), n_envs, seed=3)
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table if interp else None, seed=3)
        env.set_launch_options(max_ctas=2)
        assert env.launch_geometry()["kernel"].startswith("mdr::step_pipe_kernel")
        run("pipelined N=%d C=%d%s" % (n, nb, " interp" if interp else ""), env, 6)
    cfg, flat = config(100, True)
    pop = mdr_b200.synthetic_population(flat, 60, seed=4)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table, seed=4, action_source="bangbang", with_obs=False)
    env.set_launch_options(max_ctas=2, no_fused=True)
    run("pipelined no-obs bangbang", env, 6, use_actions=False)

if "generic" in which:
    cases = [
        ("fp64", 5, 100, dict(interp=True, solar=True, signal="sinusoidals")),
        ("fp64", 3, 160, dict(interp=True, penalty="common_L2", signal="regular_steps")),
        ("fp32", 9, 37, dict(defect=0.3, penalty="mixture", state=("hour", "day", "solar_gain", "thermal", "hvac"))),
        ("fp64", 2, 1000, dict(interp=True, penalty="common_max")),
        ("fp32", 4, 25, dict(comm="neighbours_2D")),
        ("fp64", 6, 12, dict(comm="closed_groups", signal="flat")),
    ]
    for prec, n_envs, n, kw in cases:
        cfg, flat = config(n, **kw)
        pop = mdr_b200.synthetic_population(flat, n_envs, seed=5)
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=prec, interp_table=table if kw.get("interp") else None, seed=5)
        env.set_launch_options(no_pipeline=True)
        env.enable_metrics()
        run("generic %s N=%d %s" % (prec, n, ",".join("%s" % v for v in kw.values())), env, 4)
    cfg, flat = config(40)
    pop = mdr_b200.synthetic_population(flat, 6, seed=6)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64", seed=6, action_source="greedy")
    run("generic greedy", env, 4, use_actions=False)

if "fused" in which:
    for prec, n_envs, n, interp, metrics in (("fp32", 20, 100, False, True), ("fp64", 7, 50, True, True), ("fp32", 5, 500, True, False)):
        cfg, flat = config(n, interp)
        pop = mdr_b200.synthetic_population(flat, n_envs, seed=7)
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=prec, interp_table=table if interp else None, seed=7,
                                            action_source="bangbang", with_obs=False)
        if metrics:
            env.enable_metrics()
        run("fused %s N=%d%s" % (prec, n, " interp" if interp else ""), env, 2, n_steps=80, use_actions=False)

if "cluster" in which and hasattr(mdr_b200._lib, "HAS_CLUSTER_PATH"):
    for prec, n_envs, n, kw in (("fp64", 2, 1000, dict(interp=True)), ("fp32", 3, 4096, dict()), ("fp32", 1, 20000, dict(interp=True))):
        cfg, flat = config(n, **kw)
        pop = mdr_b200.synthetic_population(flat, n_envs, seed=8)
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=prec, interp_table=table if kw.get("interp") else None, seed=8)
        run("large cluster %s N=%d" % (prec, n), env, 3)

if "populate" in which:
    cfg, flat = config(50)
    pop = mdr_b200.synthetic_population(flat, 16, seed=9)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=9)
    env.reset_envs()
    mask = torch.zeros(16, dtype=torch.uint8, device="cuda")
    mask[3] = mask[9] = 1
    env.reset_envs(mask)
    torch.cuda.synchronize()
    print("sanitize_driver: populate + masked reset -> ok", flush=True)
print("sanitize_driver: done")
