"""Fused multi-step ("deploy") kernel and the on-device metric accumulators, through the C ABI:
  * K fused steps == K single-step launches (same Philox counters, same arithmetic);
  * fp64 fused bang-bang rollout vs the numpy oracle stepped with the bang-bang rule (no random draws:
    temp_std = 0, sinusoidal / flat / regular-step signals), bit-exact integers, 1e-9 on the reals;
  * the accumulators vs the quantities main-deploy.py:124-209 sums, recomputed with torch from per-step outputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(n, signal="sinusoidals", temp_std=None, solar=False):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["power_grid_prop"]["base_power_mode"] = "constant"
    ep["power_grid_prop"]["signal_mode"] = signal
    if temp_std is not None:
        ep["cluster_prop"]["temp_parameters"][ep["cluster_prop"]["temp_mode"]]["temp_std"] = temp_std
    cfg["default_house_prop"]["solar_gain_bool"] = solar
    return cfg


def _env(cfg, n_envs, precision, action_source, seed=3):
    import mdr_b200
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=seed, action_source=action_source, with_obs=False)
    env.reset_tensor()
    return flat, pop, env


def _per_step(env, k):
    env.set_launch_options(no_fused=True)
    try:
        for _ in range(k):
            out = env.step_tensor(None)
    finally:
        env.set_launch_options(no_fused=False)
    return out


@pytest.mark.parametrize("precision,n_envs,n,source,signal", [
    ("fp64", 7, 50, "bangbang", "perlin"), ("fp64", 5, 37, "random", "sinusoidals"), ("fp32", 300, 100, "bangbang", "perlin"),
    ("fp32", 9, 3, "random", "flat"), ("fp64", 3, 500, "bangbang", "regular_steps")])
def test_fused_equals_single_steps(precision, n_envs, n, source, signal):
    import torch
    cfg = _cfg(n, signal)
    _, _, a = _env(cfg, n_envs, precision, source)
    _, _, b = _env(cfg, n_envs, precision, source)
    k = 77
    _, rew_a, p_a, s_a = a.run(k)
    _, rew_b, p_b, s_b = _per_step(b, k)
    torch.cuda.synchronize()
    assert torch.equal(a.hvac, b.hvac)
    assert torch.equal(a.t_epoch, b.t_epoch) and torch.equal(p_a, p_b)
    tol = dict(rtol=0, atol=1e-9) if precision == "fp64" else dict(rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(a.temps, b.temps, **tol)
    torch.testing.assert_close(rew_a, rew_b, **tol)
    torch.testing.assert_close(a.env["od_temp"], b.env["od_temp"], rtol=0, atol=1e-9)
    torch.testing.assert_close(s_a, s_b, rtol=1e-12, atol=1e-9)
    # a second fused call continues from the first (step counter, clock)
    a.run(5)
    _per_step(b, 5)
    assert torch.equal(a.hvac, b.hvac) and torch.equal(a.t_epoch, b.t_epoch)


@pytest.mark.parametrize("signal,solar", [("sinusoidals", False), ("flat", True), ("regular_steps", False)])
def test_fused_fp64_vs_oracle_bangbang(signal, solar):
    import torch
    from oracle import mdr_oracle as orc
    n_envs, n, k = 4, 23, 150
    cfg = _cfg(n, signal, temp_std=0.0, solar=solar)
    flat, pop, env = _env(cfg, n_envs, "fp64", "bangbang")
    oracle = orc.OracleEnv(cfg, {kk: v for kk, v in pop.items() if kk != "perlin_seed"})
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), 0.0)
    zeros = np.zeros(n_envs)
    for _ in range(k):
        act = (oracle.s["t_air"] > oracle.s["target"]).astype(np.uint8)  # agents/bangbang_controllers.py:50-61
        _, o_rew, o_p, o_s = oracle.step(act, zeros, zeros)
    _, rew, p, s = env.run(k)
    torch.cuda.synchronize()
    assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"])
    assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"])
    assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"])
    assert np.array_equal(p.cpu().numpy(), o_p)
    np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(env.t_mass.cpu().numpy(), oracle.s["t_mass"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=0, atol=1e-9)
    np.testing.assert_allclose(s.cpu().numpy(), o_s, rtol=1e-12, atol=0)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_device_metrics_match_per_step_recomputation(precision):
    import torch
    from mdr_b200 import _lib
    n_envs, n, k = 11, 60, 90
    cfg = _cfg(n, "perlin")
    _, _, a = _env(cfg, n_envs, precision, "bangbang")
    _, _, b = _env(cfg, n_envs, precision, "bangbang")
    a.enable_metrics()
    a.run(40)
    a.run(k - 40)  # accumulates across calls
    ref = torch.zeros(n_envs, _lib.N_METRICS, dtype=torch.float64, device="cuda")
    target = b.coef_b[..., 3].double()
    for _ in range(k):
        _, rew, p, s = _per_step(b, 1)
        err = b.t_air.double() - target
        d = s - p
        ref[:, 0] += 1
        ref[:, 1] += rew.double().sum(1) / n
        ref[:, 2] += err.sum(1) / n
        ref[:, 3] += err.abs().sum(1) / n
        ref[:, 4] += (err * err).sum(1)
        mx = err.abs().max(1).values
        ref[:, 5] += mx * mx
        ref[:, 6] = torch.maximum(ref[:, 6], mx)
        ref[:, 7] += b.env["od_temp"]
        ref[:, 8] += s
        ref[:, 9] += p
        ref[:, 10] += d
        ref[:, 11] += d.abs()
        ref[:, 12] += d * d
    torch.cuda.synchronize()
    rtol = 1e-9 if precision == "fp64" else 2e-4
    torch.testing.assert_close(a.metrics, ref, rtol=rtol, atol=1e-6 if precision == "fp64" else 1e-2)
    summ = a.metrics_summary()
    assert torch.all(summ["steps"] == k) and torch.isfinite(summ["rmse_temp"]).all()


@pytest.mark.parametrize("precision,penalty,interp", [("fp64", "mixture", True), ("fp32", "individual_L2", False), ("fp64", "common_max", False)])
def test_single_step_metrics_with_array_actions(precision, penalty, interp):
    """Configurations that cannot take the fused path (policy actions from an array, interpolated base power, common_*
    penalties) accumulate the same 13 quantities in the generic single-step kernel."""
    import torch
    import mdr_b200
    import golden_util as gu
    from mdr_b200 import _lib
    n_envs, n, k = 6, 45, 160 if interp else 40
    cfg = _cfg(n, "perlin")
    cfg["default_env_prop"]["reward_prop"]["temp_penalty_mode"] = penalty
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=4)
    table = gu.synthetic_table() if interp else None
    a = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=4, interp_table=table)
    b = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=4, interp_table=table)
    a.reset_tensor()
    b.reset_tensor()
    a.enable_metrics()
    ref = torch.zeros(n_envs, _lib.N_METRICS, dtype=torch.float64, device="cuda")
    target = b.coef_b[..., 3].double()
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(k):
        act = (torch.rand(n_envs, n, device="cuda", generator=g) < 0.4).to(torch.uint8)
        a.step_tensor(act)
        _, rew, p, s = b.step_tensor(act)
        err = b.t_air.double() - target
        d = s - p
        mx = err.abs().max(1).values
        upd = [torch.ones(n_envs, device="cuda", dtype=torch.float64), rew.double().sum(1) / n, err.sum(1) / n, err.abs().sum(1) / n,
               (err * err).sum(1), mx * mx, None, b.env["od_temp"], s, p, d, d.abs(), d * d]
        for i, u in enumerate(upd):
            if u is not None:
                ref[:, i] += u
        ref[:, 6] = torch.maximum(ref[:, 6], mx)
    torch.cuda.synchronize()
    # (b runs the pipelined kernel in fp32, a the generic one: compare at the precision's tolerance)
    torch.testing.assert_close(a.metrics, ref, rtol=1e-9 if precision == "fp64" else 2e-4, atol=1e-6 if precision == "fp64" else 1e-2)
    assert torch.equal(a.hvac, b.hvac)


def test_fused_full_size_day_slice_is_deterministic():
    """c3 shape (10,000 clusters x 100 houses): 600 fused steps twice from the same state give identical bits."""
    import torch
    cfg = _cfg(100, "perlin")
    _, _, a = _env(cfg, 10000, "fp32", "bangbang")
    _, _, b = _env(cfg, 10000, "fp32", "bangbang")
    a.run(600)
    b.run(300)
    b.run(300)
    torch.cuda.synchronize()
    assert torch.equal(a.temps, b.temps) and torch.equal(a.hvac, b.hvac) and torch.equal(a.env["signal"], b.env["signal"])
    assert torch.isfinite(a.temps).all()


@pytest.mark.parametrize("n_envs,n,signal", [(6, 40, "sinusoidals"), (3, 300, "flat"), (2, 1000, "regular_steps")])
def test_greedy_myopic_on_device_vs_oracle(n_envs, n, signal):
    """MDR_ACT_GREEDY (agents/greedy_myopic_controller.py:29-49 on the device) in fp64 against the oracle stepped with
    oracle.greedy_myopic_actions, which is pinned to the reference controller (tests/golden/mc_greedy_myopic.npz)."""
    import torch
    import mdr_b200
    from oracle import mdr_oracle as orc
    cfg = _cfg(n, signal, temp_std=0.0)
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=9)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64", seed=9, action_source="greedy", with_obs=True)
    env.reset_tensor()
    oracle = orc.OracleEnv(cfg, {kk: v for kk, v in pop.items() if kk != "perlin_seed"})
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), 0.0)
    zeros = np.zeros(n_envs)
    power = oracle.s["cap"] / flat.hvac_cop
    on_frac = []
    for _ in range(60):
        act = orc.greedy_myopic_actions(oracle.s["t_air"], oracle.s["target"], power, oracle.s["lockout"], oracle.s["signal"])
        o_obs, o_rew, o_p, o_s = oracle.step(act, zeros, zeros)
        obs, rew, p, s = env.step_tensor(None)
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"])
        on_frac.append(float(act.mean()))
    torch.cuda.synchronize()
    assert 0.02 < np.mean(on_frac) < 0.98  # the controller actually switches things
    assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"])
    assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"])
    assert np.array_equal(p.cpu().numpy(), o_p)
    np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=0, atol=1e-9)
    np.testing.assert_allclose(obs.cpu().numpy(), o_obs, rtol=0, atol=1e-9)


def _interp_env(n_envs, n, precision, source, signal, temp_std=None, seed=5):
    import mdr_b200
    import golden_util as gu
    cfg = _cfg(n, signal, temp_std=temp_std)
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "interpolation"
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=seed, action_source=source, with_obs=False,
                                        interp_table=gu.synthetic_table())
    env.reset_tensor()
    return cfg, flat, pop, env


@pytest.mark.parametrize("precision,n_envs,n,source,signal", [
    ("fp64", 5, 40, "bangbang", "perlin"), ("fp32", 200, 100, "bangbang", "perlin"), ("fp64", 3, 100, "random", "sinusoidals"),
    ("fp32", 7, 13, "random", "flat")])
def test_fused_with_interpolated_base_power_equals_single_steps(precision, n_envs, n, source, signal):
    """Interpolated base power inside the fused kernel (refresh every 75 steps from the houses' state): 230 fused steps
    (three refreshes, at different offsets for the two calls) == 230 single-step launches."""
    import torch
    _, _, _, a = _interp_env(n_envs, n, precision, source, signal)
    _, _, _, b = _interp_env(n_envs, n, precision, source, signal)
    a.enable_metrics()
    a.run(100)
    _, rew_a, p_a, s_a = a.run(130)
    _, rew_b, p_b, s_b = _per_step(b, 230)
    torch.cuda.synchronize()
    assert torch.equal(a.hvac, b.hvac) and torch.equal(a.t_epoch, b.t_epoch) and torch.equal(p_a, p_b)
    assert torch.equal(a.time_since_interp, b.time_since_interp)
    tol = dict(rtol=0, atol=1e-9) if precision == "fp64" else dict(rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(a.temps, b.temps, **tol)
    torch.testing.assert_close(rew_a, rew_b, **tol)
    rs = 1e-12 if precision == "fp64" else 1e-5
    torch.testing.assert_close(a.env["base_power"], b.env["base_power"], rtol=rs, atol=1e-6)
    torch.testing.assert_close(s_a, s_b, rtol=rs, atol=1e-6)
    assert float(a.metrics[:, 0].min()) == 230


def test_fused_interpolation_fp64_vs_oracle():
    import torch
    import golden_util as gu
    from oracle import mdr_oracle as orc
    n_envs, n, k = 3, 31, 200
    cfg, flat, pop, env = _interp_env(n_envs, n, "fp64", "bangbang", "sinusoidals", temp_std=0.0)
    table = gu.synthetic_table()
    oracle = orc.OracleEnv(cfg, {kk: v for kk, v in pop.items() if kk != "perlin_seed"},
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS))
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), 0.0)
    zeros = np.zeros(n_envs)
    for _ in range(k):
        act = (oracle.s["t_air"] > oracle.s["target"]).astype(np.uint8)
        _, o_rew, o_p, o_s = oracle.step(act, zeros, zeros)
    _, rew, p, s = env.run(k)
    torch.cuda.synchronize()
    assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"])
    assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"])
    assert np.array_equal(p.cpu().numpy(), o_p)
    np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(env.env["base_power"].cpu().numpy(), oracle.s["base_power"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(s.cpu().numpy(), o_s, rtol=1e-12, atol=0)
    np.testing.assert_allclose(rew.cpu().numpy(), o_rew, rtol=0, atol=1e-9)
