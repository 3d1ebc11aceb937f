"""The STEADY-STATE loop of the persistent pipelined kernel (`mdr::step_pipe_kernel`, the kernel bench.py times)
against the oracle.  The small parity cases elsewhere give every CTA a single tile; here

* the persistent grid is capped (`MdrConfig.max_ctas`) so that every CTA walks >= 20 tiles: double-buffered
  cp.async stages, several wraps of the prologue warp's mbarrier ring, the bulk-store drain before a staging
  tile is reused, ragged last tiles, interpolation refreshes that fall due in late tiles (staggered clocks),
  message counts other than 10;
* at the full BASELINE sizes (c4: 16 384 x 100, c2: 4 096 x 50) 64 sampled clusters -- including the last tile
  and tiles far beyond the first wave of the 444-CTA grid -- are compared with the oracle for 160 steps on every
  output (temperatures, integer state, reward, all observation columns, signal, base power);
* the pipelined kernel is compared with the generic kernel over 300 back-to-back launches at full size
  (programmatic dependent launch, ring reuse across launches);
* fp32 is compared with the fp64 oracle over config 3's whole simulated day (21 600 steps).

Reference: env/MA_DemandResponse.py:174-210 and everything it calls.
"""
import numpy as np
import pytest

import golden_util as gu
from oracle import mdr_oracle as orc

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-4, atol=2e-4)        # fp32: 1e-4 relative (north star); floor for values around 0 (degC-scale)
TOL_OBS = dict(rtol=1e-4, atol=5e-5)    # normalised observation features are O(1): tighter floor
TOL_W = dict(rtol=1e-4, atol=1e-2)      # watts


def _config(n, interp, signal, nb_comm=10):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["cluster_prop"]["nb_agents_comm"] = nb_comm
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    ep["power_grid_prop"]["signal_mode"] = signal
    cfg["default_house_prop"]["solar_gain_bool"] = False
    cfg["default_house_prop"]["deadband"] = 0.5
    return cfg, mdr_b200.FlatConfig(cfg)


def _subset(pop, ids):
    return {k: np.asarray(v)[ids] for k, v in pop.items() if k != "perlin_seed"}


def _compare(env, oracle, ids, out, o_out, t):
    obs, rew, p, s = out
    o_obs, o_rew, o_p, o_s = o_out
    ix = np.asarray(ids)
    hv = env.hvac[ix].cpu().numpy()
    assert np.array_equal(hv & 1, oracle.s["on"]), t
    assert np.array_equal((hv >> 1) & 1, oracle.s["lockout"]), t
    assert np.array_equal(hv >> 2, oracle.s["sso"]), t
    assert np.array_equal(p[ix].cpu().numpy(), o_p), t
    np.testing.assert_allclose(s[ix].cpu().numpy(), o_s, err_msg="signal %d" % t, **TOL_W)
    np.testing.assert_allclose(env.env["base_power"][ix].cpu().numpy(), oracle.s["base_power"], err_msg="base %d" % t, **TOL_W)
    np.testing.assert_allclose(env.temps[ix].cpu().numpy()[..., 0], oracle.s["t_air"], err_msg="t_air %d" % t, **TOL)
    np.testing.assert_allclose(env.temps[ix].cpu().numpy()[..., 1], oracle.s["t_mass"], err_msg="t_mass %d" % t, **TOL)
    np.testing.assert_allclose(rew[ix].cpu().numpy(), o_rew, err_msg="reward %d" % t, **TOL)
    np.testing.assert_allclose(obs[ix].cpu().numpy(), o_obs, err_msg="obs %d" % t, **TOL_OBS)
    assert np.array_equal(env.time_since_interp[ix].cpu().numpy(), oracle.s["time_since_interp"]), t


@pytest.mark.parametrize("n_envs,n,interp,signal,nb_comm,max_ctas", [
    (239, 50, False, "perlin", 10, 3),        # 60 tiles of 4 envs (last one ragged) on 3 CTAs: 20 tiles per CTA
    (89, 100, True, "perlin", 10, 2),         # c4 tile shape, 45 tiles on 2 CTAs, staggered refresh clocks
    (70, 160, True, "sinusoidals", 10, 3),    # one env per tile, sampled interpolation ids, 23+ tiles per CTA
    (301, 30, True, "regular_steps", 4, 2),   # generic message count (kC = 0), 7 envs per tile, 22 tiles per CTA
    (53, 224, False, "flat", 10, 2),          # largest cluster of the pipelined kernel, 26+ tiles per CTA
    (45, 500, True, "perlin", 10, 6),         # split kernel: 3 CTAs per env, 2 clusters walk 22+ envs each, inline refreshes
    (64, 1000, False, "sinusoidals", 10, 5),  # split kernel, c3big's tile shape (5 x 200): ONE cluster walks all 64 envs
])
@pytest.mark.parametrize("tiles", ["claimed", "strided"])
def test_capped_grid_many_tiles_per_cta(n_envs, n, interp, signal, nb_comm, max_ctas, tiles):
    """`tiles`: the pipelined kernel's two schedules -- tiles claimed in address order (all warps produce the per-env
    records behind a grid barrier, due tiles first and refreshed by their own CTA; the default from 14 tiles per CTA,
    i.e. here) and the fixed strided list per CTA (prologue warp + shared-memory ring, shared due-tile queue)."""
    import mdr_b200
    if n > 224 and tiles == "strided":
        pytest.skip("the split kernel has one schedule")
    steps = 160 if interp else 40
    cfg, flat = _config(n, interp, signal, nb_comm)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=500 + n)
    rng = np.random.default_rng(600 + n)
    if interp:  # refreshes fall due at different steps in different tiles (early and late in a CTA's walk)
        pop["time_since_interp"] = (rng.integers(0, flat.interp_update_period // flat.time_step, n_envs) * flat.time_step)
        pop["base_power"] = rng.uniform(2000.0, 5000.0, n_envs) * n
    table = gu.synthetic_table() if interp else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table)
    env.set_launch_options(max_ctas=max_ctas, static_tiles=(tiles == "strided"))
    with_metrics = interp and n <= 224   # the accumulator variant of the same kernel (signal terms of due envs come from the refresh)
    if with_metrics:
        env.enable_metrics()
    geo = env.launch_geometry()
    assert geo["kernel"].startswith("mdr::step_pipe_split_kernel" if n > 224 else "mdr::step_pipe_kernel")
    assert geo["tiles"] >= 20 * max_ctas, geo
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"},
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    # both sides start from the population's signal state (no reset step: the staggered clocks must survive)
    sig0 = rng.uniform(3000.0, 5000.0, n_envs) * n
    oracle.s["signal"][:] = sig0
    env.env["signal"].copy_(__import__("torch").as_tensor(sig0))
    env.precompute()
    ids_all = list(range(n_envs))
    refreshes = 0
    acc = np.zeros((n_envs, 13))
    for t in range(steps):
        act = rng.integers(0, 2, (n_envs, n)).astype(np.uint8)
        odn, sgn = rng.normal(0, 0.5, n_envs), rng.uniform(-0.5, 0.5, n_envs)
        ids = rng.integers(0, n, (n_envs, flat.interp_nb_agents)).astype(np.int32)
        before = oracle.s["base_power"].copy()
        o_out = oracle.step(act, odn, sgn, ids)
        refreshes += int((oracle.s["base_power"] != before).sum())
        out = env.step_tensor(act, od_noise=odn, signal_noise=sgn, interp_ids=ids)
        _compare(env, oracle, ids_all, out, o_out, t)
        # main-deploy.py:124-149 / metrics.py:22-30 on the oracle's outputs
        err = oracle.s["t_air"] - oracle.s["target"]
        d = o_out[3] - o_out[2]
        acc[:, 0] += 1
        acc[:, 1] += o_out[1].mean(axis=1)
        acc[:, 2] += err.mean(axis=1)
        acc[:, 3] += np.abs(err).mean(axis=1)
        acc[:, 4] += (err ** 2).sum(axis=1)
        acc[:, 5] += np.abs(err).max(axis=1) ** 2
        acc[:, 6] = np.maximum(acc[:, 6], np.abs(err).max(axis=1))
        acc[:, 7] += oracle.s["od_temp"]
        acc[:, 8] += o_out[3]
        acc[:, 9] += o_out[2]
        acc[:, 10] += d
        acc[:, 11] += np.abs(d)
        acc[:, 12] += d ** 2
    if interp:
        assert refreshes >= 2 * n_envs
    if with_metrics:
        m = env.metrics.cpu().numpy()
        assert np.array_equal(m[:, 0], acc[:, 0])
        np.testing.assert_allclose(m[:, [7, 8, 9]], acc[:, [7, 8, 9]], rtol=1e-6)
        np.testing.assert_allclose(m[:, [1, 2, 3, 4, 5, 6]], acc[:, [1, 2, 3, 4, 5, 6]], rtol=2e-3, atol=1e-4)   # fp32 temperatures
        np.testing.assert_allclose(m[:, [10, 11, 12]], acc[:, [10, 11, 12]], rtol=1e-4, atol=1.0)


@pytest.mark.parametrize("n_envs,n,interp,tiles,stagger", [
    (16384, 100, True, "claimed", False),   # the default bench kernel; refresh clocks in phase: steps 75 and 150 refresh EVERY tile
    (16384, 100, True, "claimed", True),    # ... staggered clocks (what bench.py times): ~55 due tiles in every launch
    (16384, 100, True, "strided", True),
    (4096, 50, False, "strided", False),    # c2: 4.6 tiles per CTA, strided lists by default
])
def test_full_size_sampled_envs_match_oracle(n_envs, n, interp, tiles, stagger):
    """BASELINE configs 4 and 2 at full size: 64 sampled clusters against the oracle, every output, 160 steps."""
    import torch
    import mdr_b200
    steps = 160
    cfg, flat = _config(n, interp, "perlin")
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=31)
    table = gu.synthetic_table() if interp else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table)
    env.set_launch_options(static_tiles=(tiles == "strided"))
    geo = env.launch_geometry()
    assert geo["kernel"].startswith("mdr::step_pipe_kernel")
    g = geo["envs_per_cta"]
    rng = np.random.default_rng(32)
    # first tile, last tile (both envs), tiles beyond 8 waves of the 444-CTA grid (c2 has 2.3 waves: beyond the first),
    # and a random spread
    far = 444 * 8 * g if 444 * 10 * g < n_envs else 444 * g
    special = [0, 1, n_envs - 1, n_envs - 2, far, far + 1, far + g * 444 + 3]
    ids = sorted(set(special) | set(int(i) for i in rng.integers(0, n_envs, 57)) | set(int(i) for i in rng.integers(far, n_envs, 8)))
    assert len(ids) >= 64 and max(ids) == n_envs - 1 and sum(i >= far for i in ids) >= 8
    oracle = orc.OracleEnv(cfg, _subset(pop, ids), interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    sgn0 = rng.uniform(-0.5, 0.5, n_envs)
    for k, e in enumerate(ids):
        oracle.grid_step(k, orc.to_datetime(oracle.s["t_epoch"][k]), sgn0[e])
    obs0 = env.reset_tensor(signal_noise=sgn0)
    np.testing.assert_allclose(obs0[np.asarray(ids)].cpu().numpy(), oracle.obs(), **TOL_OBS)
    if stagger:  # like a rollout whose clusters were reset at different times (bench.py's default)
        env.stagger_interp_clock(seed=30)
        oracle.s["time_since_interp"][:] = env.time_since_interp[np.asarray(ids)].cpu().numpy()
    gen = torch.Generator(device="cuda").manual_seed(33)
    for t in range(steps):
        act = (torch.rand(n_envs, n, device="cuda", generator=gen) < 0.5).to(torch.uint8)
        odn, sgn = rng.normal(0, 0.5, n_envs), rng.uniform(-0.5, 0.5, n_envs)
        out = env.step_tensor(act, od_noise=odn, signal_noise=sgn)
        o_out = oracle.step(act[np.asarray(ids)].cpu().numpy(), odn[ids], sgn[ids])
        if t % 4 == 0 or t >= steps - 12 or (interp and 70 <= t % 75 <= 76):
            _compare(env, oracle, ids, out, o_out, t)
    _compare(env, oracle, ids, out, o_out, steps)


@pytest.mark.parametrize("workload", ["c4", "c2"])
def test_pipelined_equals_generic_over_300_launches(workload):
    """Back-to-back launches (PDL, ring and stage reuse across launches, deferred refreshes) at full size against the
    generic kernel on the same inputs: integer state bit-exact, temperatures / observations within fp32 rounding."""
    import torch
    import bench
    import mdr_b200
    steps = 300
    w = bench.WORKLOADS[workload]
    cfg = bench.workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    n_envs, n = w["envs"], w["houses"]
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=77)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    mk = lambda: mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=77, interp_table=table,
                                               action_source=w["action_source"], with_obs=w["obs"])
    a, b = mk(), mk()
    b.set_launch_options(no_pipeline=True, no_fused=True)
    a.set_launch_options(no_fused=True)
    assert a.launch_geometry()["kernel"].startswith("mdr::step_pipe_kernel") and b.launch_geometry()["kernel"] == "mdr::step_kernel"
    a.reset_tensor()
    b.reset_tensor()
    if w["interp"]:
        a.stagger_interp_clock(seed=3)
        b.time_since_interp.copy_(a.time_since_interp)
    gen = torch.Generator(device="cuda").manual_seed(5)
    ring = [(torch.rand(n_envs, n, device="cuda", generator=gen) < 0.5).to(torch.uint8) for _ in range(8)]
    use = w["action_source"] == "array"
    for t in range(steps):
        oa = a.step_tensor(ring[t & 7] if use else None)
        if t % 100 == 99:
            while b.step_index < a.step_index:
                ob = b.step_tensor(ring[b.step_index & 7] if use else None)
            torch.cuda.synchronize()
            assert torch.equal(a.hvac, b.hvac), (t, "hvac")
            assert torch.equal(a.t_epoch, b.t_epoch) and torch.equal(a.time_since_interp, b.time_since_interp), t
            assert torch.equal(oa[2], ob[2]), (t, "power")
            torch.testing.assert_close(a.temps, b.temps, rtol=1e-4, atol=2e-3)   # bang-bang may diverge by a flip
            torch.testing.assert_close(oa[3], ob[3], rtol=1e-4, atol=1e-2)
            torch.testing.assert_close(oa[1], ob[1], rtol=1e-3, atol=2e-3)
            if oa[0] is not None:
                torch.testing.assert_close(oa[0], ob[0], rtol=1e-4, atol=2e-3)


def test_fp32_over_one_simulated_day():
    """Config 3's duration: 21 600 steps (24 h at 4 s) in fp32 against the fp64 oracle, actions recorded from the
    oracle's own bang-bang decisions (so the integer state must stay bit-exact).  The affine increment form
    T += (M - I)(T - T_ss) keeps the fp32 error at rounding level; the bound is the north star's 1e-4 relative."""
    import torch
    import mdr_b200
    steps, n_envs, n = 21600, 2, 24
    cfg, flat = _config(n, False, "sinusoidals")
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=91)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32")
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"})
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]))
    env.reset_tensor()
    rng = np.random.default_rng(92)
    odn = rng.normal(0, 0.5, (steps, n_envs))
    worst_rel = worst_abs = 0.0
    act = torch.empty(n_envs, n, dtype=torch.uint8, device="cuda")
    for t in range(steps):
        a = (oracle.s["t_air"] > oracle.s["target"]).astype(np.uint8)   # agents/bangbang_controllers.py:50-61
        o_obs, o_rew, o_p, o_s = oracle.step(a, odn[t])
        act.copy_(torch.from_numpy(a))
        obs, rew, p, s = env.step_tensor(act, od_noise=odn[t])
        if t % 600 == 599 or t == steps - 1:
            hv = env.hvac.cpu().numpy()
            assert np.array_equal(hv >> 2, oracle.s["sso"]) and np.array_equal(hv & 1, oracle.s["on"]), t
            assert np.array_equal(p.cpu().numpy(), o_p), t
            tt = env.temps.cpu().numpy().astype(np.float64)
            for k, col in (("t_air", 0), ("t_mass", 1)):
                d = np.abs(tt[..., col] - oracle.s[k])
                worst_abs = max(worst_abs, float(d.max()))
                worst_rel = max(worst_rel, float((d / np.abs(oracle.s[k])).max()))
            np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=1e-4, atol=2e-4)
            np.testing.assert_allclose(obs.cpu().numpy(), o_obs, rtol=1e-4, atol=5e-5)
    print("fp32 vs fp64 oracle over %d steps: max |dT| %.3g K, max relative %.3g" % (steps, worst_abs, worst_rel))
    assert worst_rel < 1e-4, (worst_rel, worst_abs)
