"""Clusters larger than one CTA (env/MA_DemandResponse.py:1005-1055 loops over any nb_agents; :1209-1215 samples 100
houses when N > 100): an env of more than 224 houses is split over the CTAs of a thread-block cluster (per-CTA
totals of power / penalties / metrics through distributed shared memory, neighbour messages read from the owning
CTA), beyond 16 384 houses over plain CTAs with a two-pass reduction.  Compared with the oracle on every output."""
import numpy as np
import pytest

import golden_util as gu
from oracle import mdr_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {"fp64": dict(rtol=0.0, atol=1e-9), "fp32": dict(rtol=1e-4, atol=2e-4)}
TOL_W = {"fp64": dict(rtol=1e-12, atol=1e-9), "fp32": dict(rtol=1e-4, atol=1e-2)}


def _case(n_envs, n, seed, interp=False, penalty="individual_L2", signal="perlin", comm="neighbours", solar=False,
          defect=0.0, flags=()):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["cluster_prop"]["agents_comm_mode"] = comm
    ep["cluster_prop"]["comm_defect_prob"] = defect
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    ep["power_grid_prop"]["signal_mode"] = signal
    ep["reward_prop"]["temp_penalty_mode"] = penalty
    for k in flags:
        ep["state_properties"][k] = True
    cfg["default_house_prop"]["solar_gain_bool"] = solar
    cfg["default_house_prop"]["deadband"] = 0.5
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    return cfg, flat, pop


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n_envs,n,kw", [
    (2, 225, dict()),                                                     # smallest split: 2 CTAs
    (301, 300, dict(interp=True)),                                        # many envs: persistent split kernel, 2 CTAs per env, sampled refresh
    (160, 1000, dict(solar=True, signal="sinusoidals")),                  # c3big's tile shape (5 x 200), solar gain on
    (9, 1790, dict(interp=True)),                                         # largest env of the split pipelined kernel (8 x 224)
    (3, 1000, dict(interp=True)),                                         # config 1's shape: 5 CTAs x 200, sampled interpolation
    (2, 1025, dict(penalty="common_L2", signal="sinusoidals")),          # just past the one-CTA limit
    (2, 4096, dict(interp=True, penalty="mixture", solar=True)),         # 9 CTAs x 456
    (1, 4096, dict(comm="random_fixed", penalty="common_max", defect=0.2, flags=("thermal", "hour", "day"))),
    (1, 16384, dict(interp=True, signal="regular_steps")),               # 16 CTAs x 1024 (non-portable cluster size)
    (1, 5000, dict(comm="neighbours_2D", signal="flat")),
])
def test_split_env_matches_oracle(n_envs, n, kw, precision):
    import random
    import mdr_b200
    cfg, flat, pop = _case(n_envs, n, 700 + n, **kw)
    interp = kw.get("interp", False)
    steps = 80 if interp else 12
    table = gu.synthetic_table() if interp else None
    comm = None
    if flat.comm_mode_name == "random_fixed":
        rnd = random.Random(5)
        comm = mdr_b200.comm_table("random_fixed", n, flat.nb_agents_comm, sampler=lambda possible, k: rnd.sample(possible, k=k))
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, interp_table=table, comm_table=comm)
    geo = env.launch_geometry()
    # (fp64 with many envs of <= 1024 houses keeps one CTA per env: more, smaller CTAs of the non-persistent kernel only add waves)
    assert (geo["cluster_size"] >= 2 or (precision == "fp64" and n <= 1024 and n_envs >= 148)) and geo["kernel"].startswith("mdr::step_"), geo
    if precision == "fp32" and not kw.get("penalty") and not kw.get("comm") and n <= 1792:
        assert geo["kernel"].startswith("mdr::step_pipe_split_kernel"), geo
    with_metrics = "penalty" in kw or "comm" in kw or n == 300   # (metrics are an epilogue variant of every kernel)
    if with_metrics:
        env.enable_metrics()
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"}, comm_table=comm,
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    rng = np.random.default_rng(800 + n)
    c = flat.n_comm
    defect = kw.get("defect", 0.0)
    draw = lambda: dict(sgn=rng.uniform(-0.5, 0.5, n_envs), ids=rng.integers(0, n, (n_envs, flat.interp_nb_agents)).astype(np.int32),
                        keep=(rng.random((n_envs, n, c)) > defect).astype(np.uint8))
    d = draw()
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), d["sgn"][e], d["ids"][e])
    obs = env.reset_tensor(signal_noise=d["sgn"], interp_ids=d["ids"], msg_keep=d["keep"])
    tol, tolw = TOL[precision], TOL_W[precision]
    np.testing.assert_allclose(env.env["signal"].cpu().numpy(), oracle.s["signal"], **tolw)
    np.testing.assert_allclose(obs.cpu().numpy(), oracle.obs(d["keep"]), **tol)
    sum_rew = np.zeros(n_envs)
    for t in range(steps):
        d = draw()
        act = rng.integers(0, 2, (n_envs, n)).astype(np.uint8)
        odn = rng.normal(0, 0.5, n_envs)
        o_obs, o_rew, o_p, o_s = oracle.step(act, odn, d["sgn"], d["ids"], d["keep"])
        obs, rew, p, s = env.step_tensor(act, od_noise=odn, signal_noise=d["sgn"], interp_ids=d["ids"], msg_keep=d["keep"])
        sum_rew += o_rew.mean(axis=1)
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"]), t
        assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"]), t
        assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"]), t
        assert np.array_equal(p.cpu().numpy(), o_p), t
        np.testing.assert_allclose(s.cpu().numpy(), o_s, err_msg="signal %d" % t, **tolw)
        np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], err_msg="t_air %d" % t, **tol)
        np.testing.assert_allclose(env.t_mass.cpu().numpy(), oracle.s["t_mass"], err_msg="t_mass %d" % t, **tol)
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, err_msg="reward %d" % t, **tol)
        np.testing.assert_allclose(obs.cpu().numpy(), o_obs, err_msg="obs %d" % t, **tol)
    if with_metrics:
        m = env.metrics.cpu().numpy()
        assert np.array_equal(m[:, 0], np.full(n_envs, steps))
        np.testing.assert_allclose(m[:, 1], sum_rew, rtol=1e-4 if precision == "fp32" else 1e-9, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n_envs,n,source,kw", [
    (70, 225, "array", dict()),                                             # smallest env of the wide kernel
    (64, 1000, "bangbang", dict(solar=True)),                               # c3big's shape: on-device bang-bang, solar gain on
    (80, 777, "array", dict(interp=True, signal="sinusoidals")),            # sampled interpolation ids (N > 100), 3.5 sub-tiles
    (64, 4096, "array", dict(interp=True, signal="regular_steps")),         # 18.3 sub-tiles per CTA
])
def test_wide_kernel_without_observation_matches_oracle(n_envs, n, source, kw, precision):
    """No observation, clusters of 225 .. 8 192 houses: one CTA walks a whole env (mdr::step_wide_kernel) -- against the
    oracle on state, power, signal, base power and reward, and bit for bit against the generic kernel."""
    import torch
    import mdr_b200
    cfg, flat, pop = _case(n_envs, n, 900 + n, **kw)
    interp = kw.get("interp", False)
    steps = 80 if interp else 12
    table = gu.synthetic_table() if interp else None
    mk = lambda: mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, interp_table=table, action_source=source,
                                               with_obs=False)
    env, ref = mk(), mk()
    env.set_launch_options(no_fused=True)
    ref.set_launch_options(no_fused=True, no_pipeline=True)   # generic kernel (split over a cluster)
    assert env.launch_geometry()["kernel"].startswith("mdr::step_wide_kernel"), env.launch_geometry()
    assert ref.launch_geometry()["kernel"] == "mdr::step_kernel", ref.launch_geometry()
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"},
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    rng = np.random.default_rng(950 + n)
    draw = lambda: dict(sgn=rng.uniform(-0.5, 0.5, n_envs), ids=rng.integers(0, n, (n_envs, flat.interp_nb_agents)).astype(np.int32))
    d = draw()
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), d["sgn"][e], d["ids"][e])
    env.reset_tensor(signal_noise=d["sgn"], interp_ids=d["ids"])
    ref.reset_tensor(signal_noise=d["sgn"], interp_ids=d["ids"])
    tol, tolw = TOL[precision], TOL_W[precision]
    for t in range(steps):
        d = draw()
        odn = rng.normal(0, 0.5, n_envs)
        if source == "array":
            act = rng.integers(0, 2, (n_envs, n)).astype(np.uint8)
        else:
            # bangbang_controllers.py:50-61 on the kernel's own operands (the target as stored in the working precision)
            act = (env.temps[..., 0] > env.coef_b[..., 3]).to(torch.uint8).cpu().numpy()
        o_obs, o_rew, o_p, o_s = oracle.step(act, odn, d["sgn"], d["ids"])
        a_in = act if source == "array" else None
        _, rew, p, s = env.step_tensor(a_in, od_noise=odn, signal_noise=d["sgn"], interp_ids=d["ids"])
        _, rew_g, p_g, s_g = ref.step_tensor(a_in, od_noise=odn, signal_noise=d["sgn"], interp_ids=d["ids"])
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"]), t
        assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"]), t
        assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"]), t
        assert np.array_equal(p.cpu().numpy(), o_p), t
        np.testing.assert_allclose(s.cpu().numpy(), o_s, err_msg="signal %d" % t, **tolw)
        np.testing.assert_allclose(env.env["base_power"].cpu().numpy(), oracle.s["base_power"], err_msg="base %d" % t, **tolw)
        np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], err_msg="t_air %d" % t, **tol)
        np.testing.assert_allclose(env.t_mass.cpu().numpy(), oracle.s["t_mass"], err_msg="t_mass %d" % t, **tol)
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, err_msg="reward %d" % t, **tol)
        assert np.array_equal(env.time_since_interp.cpu().numpy(), oracle.s["time_since_interp"]), t
        # the same arithmetic as the generic kernel, operation for operation
        assert torch.equal(env.hvac, ref.hvac) and torch.equal(env.temps, ref.temps), t
        assert torch.equal(p, p_g) and torch.equal(s, s_g) and torch.equal(rew, rew_g), t
        assert torch.equal(env.t_epoch, ref.t_epoch) and torch.equal(env.env["od_temp"], ref.env["od_temp"]), t


def test_split_equals_single_cta_bitwise():
    """225..1024 houses run either way: one 1024-thread CTA (MDR_FLAG_NO_CLUSTER) or a cluster; integer state and power
    identical, reals to rounding (the penalty mean is summed in a different order)."""
    import torch
    import mdr_b200
    cfg, flat, pop = _case(5, 777, 3, interp=True, penalty="mixture", signal="sinusoidals")
    table = gu.synthetic_table()
    a = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64", interp_table=table, seed=3)
    b = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64", interp_table=table, seed=3)
    b.set_launch_options(no_cluster=True)
    assert a.launch_geometry()["cluster_size"] >= 2 and b.launch_geometry()["cluster_size"] == 1
    a.reset_tensor()
    b.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(4)
    for t in range(160):
        act = (torch.rand(5, 777, device="cuda", generator=g) < 0.5).to(torch.uint8)
        oa, ob = a.step_tensor(act), b.step_tensor(act)   # device Philox noise / sampled ids: same counters on both
        assert torch.equal(a.hvac, b.hvac) and torch.equal(oa[2], ob[2]), t
        assert torch.equal(a.temps, b.temps), t
        torch.testing.assert_close(oa[3], ob[3], rtol=1e-13, atol=0)
        torch.testing.assert_close(oa[1], ob[1], rtol=0, atol=1e-12)
        torch.testing.assert_close(oa[0], ob[0], rtol=0, atol=1e-12)


def test_million_house_single_cluster_runs():
    """SURVEY 8d's stress shape: ONE cluster of 10^6 houses (beyond a thread-block cluster: plain CTAs, per-CTA
    partials reduced in a second pass).  Checked against the oracle on the integer state, power and a sample of houses."""
    import torch
    import mdr_b200
    n = 1_000_000
    cfg, flat, pop = _case(1, n, 11, interp=True, signal="sinusoidals")
    table = gu.synthetic_table()
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table)
    assert env.launch_geometry()["cluster_size"] == 0   # plain CTAs + workspace
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"},
                           comm_table=mdr_b200.comm_table("neighbours", n, 10),
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS))
    rng = np.random.default_rng(12)
    ids = rng.integers(0, n, (1, flat.interp_nb_agents)).astype(np.int32)
    oracle.grid_step(0, orc.to_datetime(oracle.s["t_epoch"][0]), 0.0, ids[0])
    env.reset_tensor(interp_ids=ids)
    np.testing.assert_allclose(env.env["signal"].cpu().numpy(), oracle.s["signal"], **TOL_W["fp32"])
    for t in range(3):
        act = rng.integers(0, 2, (1, n)).astype(np.uint8)
        odn = rng.normal(0, 0.5, 1)
        o_obs, o_rew, o_p, o_s = oracle.step(act, odn, None, ids)
        obs, rew, p, s = env.step_tensor(act, od_noise=odn, interp_ids=ids)
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"]), t
        assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"]), t
        assert np.array_equal(p.cpu().numpy(), o_p), t
        np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], **TOL["fp32"])
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, **TOL["fp32"])
        sel = np.r_[0:64, n // 2 - 32:n // 2 + 32, n - 64:n]
        np.testing.assert_allclose(obs[0, sel].cpu().numpy(), o_obs[0, sel], **TOL["fp32"])
    assert torch.isfinite(obs).all()
