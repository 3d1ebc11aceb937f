"""GPU parity tests (run with -m gpu on the B200 box): the CUDA step path, called through the
C ABI, against (1) the golden traces recorded from the unmodified reference and (2) the numpy
oracle on seeded random inputs.

Tolerances (BASELINE.json north_star): HVAC on/off, lockout, seconds_since_off and neighbour
indices bit-exact; temperatures, power, signal, rewards, observations within 1e-9 absolute in
fp64 mode and 1e-4 relative in fp32 mode (with an absolute floor of 1e-4 x the quantity's
natural scale where the value itself is ~0).
"""
import copy

import numpy as np
import pytest

import golden_util as gu
from oracle import mdr_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {"fp64": dict(rtol=0.0, atol=1e-9), "fp32": dict(rtol=1e-4, atol=2e-4)}
# Grid signal / base power (watts, up to 6e6): with interpolation they are sums of ~100 table lookups
# whose slope reaches 2e4 W/K on the synthetic table, so the ~1e-13 K rounding difference between the
# affine update and the reference's closed form shows up as ~1e-8 W, i.e. 1e-13 relative.  They are not
# in the north star's 1e-9-absolute list (temperatures, power, rewards); checked at 1e-12 relative.
TOL_W = {"fp64": dict(rtol=1e-12, atol=1e-9), "fp32": dict(rtol=1e-4, atol=1e-2)}


def _env_from_golden(g, precision, **kw):
    import mdr_b200
    pop = g.batched_snap()
    pop["perlin_seed"] = np.zeros(1)
    table = gu.synthetic_table() if g.uses_interp else None
    return mdr_b200.VecDemandResponseEnv(g.config, pop, precision=precision, interp_table=table, comm_table=g.comm, **kw)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", gu.names())
def test_trace_matches_reference(name, precision):
    g = gu.Golden(name)
    env = _env_from_golden(g, precision)
    tol, tolw = TOL[precision], TOL_W[precision]
    init_comm = g.z.get("init_comm")
    obs0 = env.observe_tensor(msg_keep=g.init_keep[None], comm=None if init_comm is None else init_comm[None])
    np.testing.assert_allclose(obs0[0].cpu().numpy(), g.obs0, **tol)
    ci = oi = 0
    for t in range(g.steps):
        ids = g.interp_ids[t][None] if g.interp_ids[t][0] >= 0 else None
        comm = None if g.comm_t is None else g.comm_t[t][None]
        obs, rew, p, s = env.step_tensor(g.actions[t][None], od_noise=g.od_noise[t:t + 1],
                                         signal_noise=g.sig_noise[t:t + 1], interp_ids=ids,
                                         msg_keep=g.msg_keep[t][None], comm=comm)
        assert float(p[0]) == g.power[t], (t, float(p[0]), g.power[t])
        np.testing.assert_allclose(float(s[0]), g.signal[t], err_msg="signal step %d" % t, **tolw)
        np.testing.assert_allclose(float(env.env["od_temp"][0]), g.od_temp[t], rtol=0, atol=1e-12)
        if t in g.check_steps:
            assert np.array_equal(env.hvac_on[0].cpu().numpy(), g.on[ci]), t
            assert np.array_equal(env.hvac_lockout[0].cpu().numpy(), g.lockout[ci]), t
            assert np.array_equal(env.seconds_since_off[0].cpu().numpy(), g.sso[ci]), t
            np.testing.assert_allclose(env.t_air[0].cpu().numpy(), g.t_air[ci], err_msg="t_air step %d" % t, **tol)
            np.testing.assert_allclose(env.t_mass[0].cpu().numpy(), g.t_mass[ci], err_msg="t_mass step %d" % t, **tol)
            np.testing.assert_allclose(rew[0].cpu().numpy(), g.reward[ci], err_msg="reward step %d" % t, **tol)
            ci += 1
        if t in g.obs_steps:
            np.testing.assert_allclose(obs[0].cpu().numpy(), g.obs[oi], err_msg="obs step %d" % t, **tol)
            oi += 1
    assert ci == len(g.check_steps) and oi == len(g.obs_steps)


@pytest.mark.parametrize("name", ["interp_150_sinus_solar", "interp_40_perlin", "c1_1000_fp64", "c0_bangbang_50",
                                  "hetero_37_solar_lockout"])
def test_reset_reproduces_initial_signal(name):
    """mdr_reset = PowerGrid.step(start_datetime) of build_environment (:133): from the pre-reset state
    (signal 0, time_since_last_interp = period + 1) it must land on the reference's snapshot."""
    import mdr_b200
    g = gu.Golden(name)
    pop = g.batched_snap()
    pop["perlin_seed"] = np.zeros(1)
    period = g.config["default_env_prop"]["power_grid_prop"]["base_power_parameters"]["interpolation"]["interp_update_period"]
    pop["signal"] = np.zeros(1)
    pop["base_power"] = np.zeros(1)
    pop["time_since_interp"] = np.array([period + 1])
    env = mdr_b200.VecDemandResponseEnv(g.config, pop, precision="fp64", comm_table=g.comm,
                                        interp_table=gu.synthetic_table() if g.uses_interp else None)
    ids = g.init_interp_ids[None] if g.init_interp_ids[0] >= 0 else None
    init_comm = g.z.get("init_comm")
    obs = env.reset_tensor(signal_noise=np.array([float(g.init_sig_noise)]), interp_ids=ids, msg_keep=g.init_keep[None],
                           comm=None if init_comm is None else init_comm[None])
    np.testing.assert_allclose(float(env.env["signal"][0]), float(g.snap["signal"]), **TOL_W["fp64"])
    np.testing.assert_allclose(float(env.env["base_power"][0]), float(g.snap["base_power"]), **TOL_W["fp64"])
    if g.uses_interp:
        assert int(env.time_since_interp[0]) == 0
    np.testing.assert_allclose(obs[0].cpu().numpy(), g.obs0, rtol=0, atol=1e-9)


def _random_case(n_envs, n, seed, interp, solar, steps, penalty="individual_L2", signal="sinusoidals", defect=0.0):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["cluster_prop"]["comm_defect_prob"] = defect
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    ep["power_grid_prop"]["signal_mode"] = signal
    ep["reward_prop"]["temp_penalty_mode"] = penalty
    cfg["default_house_prop"]["solar_gain_bool"] = solar
    cfg["default_house_prop"]["deadband"] = 0.5
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    rng = np.random.default_rng(seed + 1)
    c = flat.n_comm
    draws = dict(
        actions=rng.integers(0, 2, (steps, n_envs, n)).astype(np.uint8),
        od_noise=rng.normal(0, 0.5, (steps, n_envs)),
        sig_noise=rng.uniform(-0.5, 0.5, (steps, n_envs)),
        ids=rng.integers(0, n, (steps, n_envs, flat.interp_nb_agents)).astype(np.int32),
        keep=(rng.random((steps, n_envs, n, c)) > defect).astype(np.uint8),
    )
    return cfg, flat, pop, draws


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n_envs,n,interp,solar,penalty,signal", [
    (7, 50, False, False, "individual_L2", "perlin"),       # c2 shape (several envs per CTA, ragged last CTA)
    (5, 100, True, True, "individual_L2", "sinusoidals"),   # c4 shape with on-device interpolation
    (3, 160, True, False, "common_L2", "regular_steps"),    # N > interp_nb_agents: sampled ids
    (2, 1000, True, False, "mixture", "perlin"),            # c1 shape, one env per CTA, 1024 threads
    (33, 3, False, True, "common_max", "flat"),             # tiny clusters, clipped neighbour count
    (1, 1024, False, False, "individual_L2", "perlin"),     # maximum houses per env
    (4, 1, False, False, "individual_L2", "perlin"),        # single-house clusters: no messages
])
def test_batched_envs_match_oracle(n_envs, n, interp, solar, penalty, signal, precision):
    import mdr_b200
    steps = 90 if interp else 24
    cfg, flat, pop, d = _random_case(n_envs, n, 100 + n, interp, solar, steps, penalty, signal, defect=0.2 if n == 50 else 0.0)
    table = gu.synthetic_table() if interp else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, interp_table=table)
    snap = {k: v for k, v in pop.items() if k != "perlin_seed"}
    oracle = orc.OracleEnv(cfg, snap, interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    # reset: initial signal on both sides
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), d["sig_noise"][0, e], d["ids"][0, e])
    obs = env.reset_tensor(signal_noise=d["sig_noise"][0], interp_ids=d["ids"][0], msg_keep=d["keep"][0])
    tol, tolw = TOL[precision], TOL_W[precision]
    np.testing.assert_allclose(env.env["signal"].cpu().numpy(), oracle.s["signal"], **tolw)
    np.testing.assert_allclose(obs.cpu().numpy(), oracle.obs(d["keep"][0]), **tol)
    for t in range(steps):
        o_obs, o_rew, o_p, o_s = oracle.step(d["actions"][t], d["od_noise"][t], d["sig_noise"][t], d["ids"][t], d["keep"][t])
        obs, rew, p, s = env.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t],
                                         interp_ids=d["ids"][t], msg_keep=d["keep"][t])
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"]), t
        assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"]), t
        assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"]), t
        assert np.array_equal(p.cpu().numpy(), o_p), t
        np.testing.assert_allclose(s.cpu().numpy(), o_s, err_msg="signal %d" % t, **tolw)
        np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], err_msg="t_air %d" % t, **tol)
        np.testing.assert_allclose(env.t_mass.cpu().numpy(), oracle.s["t_mass"], err_msg="t_mass %d" % t, **tol)
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, err_msg="reward %d" % t, **tol)
        np.testing.assert_allclose(obs.cpu().numpy(), o_obs, err_msg="obs %d" % t, **tol)
    assert np.array_equal(env.t_epoch.cpu().numpy(), oracle.s["t_epoch"])


def test_fp64_long_trace_stays_within_1e9():
    """2 500 steps (~2.8 h simulated): rounding differences between the affine 2x2 update and the
    reference's closed form must not accumulate past 1e-9."""
    import mdr_b200
    steps = 2500
    cfg, flat, pop, d = _random_case(2, 64, 77, False, True, steps)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64")
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"})
    env.reset_tensor(signal_noise=d["sig_noise"][0])
    worst = 0.0
    for t in range(steps):
        oracle.step(d["actions"][t], d["od_noise"][t], d["sig_noise"][t])
        env.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
        if t % 250 == 249 or t == steps - 1:
            worst = max(worst, float(np.abs(env.t_air.cpu().numpy() - oracle.s["t_air"]).max()),
                        float(np.abs(env.t_mass.cpu().numpy() - oracle.s["t_mass"]).max()))
            assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"])
    assert worst < 1e-9, worst


def test_host_buffer_entry_point_matches_tensor_path():
    import mdr_b200
    cfg, flat, pop, d = _random_case(6, 50, 9, False, False, 5)
    a = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32")
    b = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32")
    a.reset_tensor(signal_noise=d["sig_noise"][0])
    b.reset_tensor(signal_noise=d["sig_noise"][0])
    for t in range(5):
        obs, rew, p, s = a.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
        h_obs, h_rew, h_p, h_s = b.step_host(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
        assert np.array_equal(obs.cpu().numpy(), h_obs) and np.array_equal(rew.cpu().numpy(), h_rew)
        assert np.array_equal(p.cpu().numpy(), h_p) and np.array_equal(s.cpu().numpy(), h_s)


@pytest.mark.parametrize("n_envs,n,interp,precision,threads", [
    (700, 50, False, "fp32", 4), (333, 100, True, "fp32", 3), (40, 300, True, "fp64", 2), (5, 1000, True, "fp32", 1),
    (100, 12, False, "fp64", 5)])
def test_pipelined_host_path_is_bit_identical(n_envs, n, interp, precision, threads):
    """mdr_step_host with an MdrHostCtx (env-axis slices over two streams, compact 16-real records over PCIe, rows
    expanded by host threads) returns exactly the bytes of the serial path and of the device tensors (utils.py:842-868)."""
    import mdr_b200
    cfg, flat, pop, d = _random_case(n_envs, n, 50 + n, interp, False, 6, signal="perlin")
    table = gu.synthetic_table() if interp else None
    mk = lambda: mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, interp_table=table, seed=11)
    a, b, c = mk(), mk(), mk()
    c.host_pipeline(True, n_threads=threads, n_slices=5)
    assert c.host_pipeline_info()["threads"] == threads
    for x in (a, b, c):
        x.reset_tensor()
        if interp:
            x.stagger_interp_clock(seed=4)
    assert c.host_transfer_bytes() < b.host_transfer_bytes()
    for t in range(6):
        obs, rew, p, s = a.step_tensor(d["actions"][t])            # device Philox noise: keys must not depend on slicing
        s_obs, s_rew, s_p, s_s = [x.copy() for x in b.step_host(d["actions"][t])]
        h_obs, h_rew, h_p, h_s = c.step_host(d["actions"][t])
        assert np.array_equal(obs.cpu().numpy(), s_obs) and np.array_equal(rew.cpu().numpy(), s_rew)
        assert np.array_equal(h_obs, s_obs), t
        assert np.array_equal(h_rew, s_rew) and np.array_equal(h_p, s_p) and np.array_equal(h_s, s_s), t
        assert np.array_equal(c.hvac.cpu().numpy(), a.hvac.cpu().numpy()) and np.array_equal(c.temps.cpu().numpy(), a.temps.cpu().numpy())
    c.host_pipeline(False)


def test_checkpoint_and_deepcopy():
    import mdr_b200
    cfg, flat, pop, d = _random_case(3, 40, 21, False, False, 8)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp64")
    env.reset_tensor(signal_noise=d["sig_noise"][0])
    for t in range(4):
        env.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
    sd = env.state_dict()
    twin = copy.deepcopy(env)
    outs = []
    for e_ in (env, twin):
        for t in range(4, 8):
            o = e_.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
        outs.append([x.clone() for x in o] + [e_.temps.clone(), e_.hvac.clone()])
    env.load_state_dict(sd)
    for t in range(4, 8):
        o = env.step_tensor(d["actions"][t], od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t])
    outs.append([x.clone() for x in o] + [env.temps.clone(), env.hvac.clone()])
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert (x == y).all()


@pytest.mark.parametrize("n_envs,n,interp,signal,nb_comm", [
    (9, 50, False, "perlin", 10),          # c2 shape, 4 envs per tile, ragged last tile
    (7, 100, True, "perlin", 10),          # c4 shape: deferred interpolation refresh, all houses sampled
    (5, 160, True, "sinusoidals", 10),     # N > interp_nb_agents: replayed sample ids in the deferred refresh
    (11, 30, True, "regular_steps", 4),    # generic message count (kC = 0 instantiation)
    (40, 7, False, "flat", 10),            # neighbour count clipped to N - 1, many envs per tile
    (3, 224, False, "perlin", 10),         # largest cluster the pipelined kernel takes
])
def test_pipelined_kernel_matches_oracle(n_envs, n, interp, signal, nb_comm):
    """The persistent pipelined fp32 kernel (no message drops, default flags -> `pipelined` geometry) against the
    oracle, including steps on which the interpolation refresh is due (deferred post-pass), and against the generic
    kernel on the same inputs (MDR_FLAG_NO_PIPELINE): integer state bit-exact, reals within the fp32 tolerance."""
    import os
    import mdr_b200
    steps = 160 if interp else 30
    cfg, flat, pop, d = _random_case(n_envs, n, 300 + n, interp, False, steps, "individual_L2", signal)
    cfg["default_env_prop"]["cluster_prop"]["nb_agents_comm"] = nb_comm
    flat = mdr_b200.FlatConfig(cfg)
    table = gu.synthetic_table() if interp else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table)
    twin = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", interp_table=table)
    twin.set_launch_options(no_pipeline=True)
    assert env.launch_geometry()["kernel"].startswith("mdr::step_pipe_kernel")
    assert twin.launch_geometry()["kernel"] == "mdr::step_kernel"
    oracle = orc.OracleEnv(cfg, {k: v for k, v in pop.items() if k != "perlin_seed"},
                           interp=orc.PowerInterp(table, gu.INTERP_GRID, gu.INTERP_KEYS) if interp else None)
    for e in range(n_envs):
        oracle.grid_step(e, orc.to_datetime(oracle.s["t_epoch"][e]), d["sig_noise"][0, e], d["ids"][0, e])
    env.reset_tensor(signal_noise=d["sig_noise"][0], interp_ids=d["ids"][0])
    twin.reset_tensor(signal_noise=d["sig_noise"][0], interp_ids=d["ids"][0])
    tol, tolw = TOL["fp32"], TOL_W["fp32"]
    refreshes = 0
    for t in range(steps):
        base_before = oracle.s["base_power"].copy()
        o_obs, o_rew, o_p, o_s = oracle.step(d["actions"][t], d["od_noise"][t], d["sig_noise"][t], d["ids"][t])
        refreshes += int((oracle.s["base_power"] != base_before).any())
        kw = dict(od_noise=d["od_noise"][t], signal_noise=d["sig_noise"][t], interp_ids=d["ids"][t])
        obs, rew, p, s = env.step_tensor(d["actions"][t], **kw)
        g_obs, g_rew, g_p, g_s = twin.step_tensor(d["actions"][t], **kw)
        assert np.array_equal(env.hvac.cpu().numpy(), twin.hvac.cpu().numpy()), t
        assert np.array_equal(env.hvac_on.cpu().numpy(), oracle.s["on"]), t
        assert np.array_equal(env.hvac_lockout.cpu().numpy(), oracle.s["lockout"]), t
        assert np.array_equal(env.seconds_since_off.cpu().numpy(), oracle.s["sso"]), t
        assert np.array_equal(p.cpu().numpy(), o_p) and np.array_equal(g_p.cpu().numpy(), o_p), t
        assert np.array_equal(env.time_since_interp.cpu().numpy(), twin.time_since_interp.cpu().numpy()), t
        np.testing.assert_allclose(s.cpu().numpy(), o_s, err_msg="signal %d" % t, **tolw)
        np.testing.assert_allclose(s.cpu().numpy(), g_s.cpu().numpy(), err_msg="signal vs generic %d" % t, **tolw)
        np.testing.assert_allclose(env.t_air.cpu().numpy(), oracle.s["t_air"], err_msg="t_air %d" % t, **tol)
        np.testing.assert_allclose(rew.cpu().numpy(), o_rew, err_msg="reward %d" % t, **tol)
        np.testing.assert_allclose(obs.cpu().numpy(), o_obs, err_msg="obs %d" % t, **tol)
        np.testing.assert_allclose(obs.cpu().numpy(), g_obs.cpu().numpy(), err_msg="obs vs generic %d" % t, **tol)
    assert np.array_equal(env.t_epoch.cpu().numpy(), oracle.s["t_epoch"])
    if interp:
        assert refreshes >= 2  # the deferred refresh pass ran at least twice


# order of golden `deploy_acc` (oracle/make_golden.py DEPLOY_KEYS, main-deploy.py:85-97) -> MDR_M_* column
_DEPLOY_TO_M = ("sum_mean_temp_offset", "sum_mean_temp_error", "max_temp_error", "sum_signal_offset", "sum_signal_error",
                "sum_od_temp", "sum_signal", "sum_consumption", "sum_sq_signal_error", "sum_sq_temp_error",
                "sum_sq_max_temp_error")


def check_metrics_against_reference(g, m, precision):
    """`m` = one env's MDR_M_* accumulators after the golden trace; golden `deploy_acc` / `train_acc` are the values the
    reference's own code produced on its own trace (main-deploy.py:124-149 executed from source, metrics.Metrics.update)."""
    from mdr_b200 import _lib
    col = {k: float(m[i]) for i, k in enumerate(_lib.METRIC_NAMES)}
    tol = dict(rtol=1e-9, atol=1e-9) if precision == "fp64" else dict(rtol=2e-4, atol=1e-3)
    assert col["steps"] == g.steps
    for ref, key in zip(g.deploy_acc, _DEPLOY_TO_M):
        np.testing.assert_allclose(col[key], ref, err_msg=key, **tol)
    n = g.n
    train = g.train_acc  # Metrics: cumul_avg_reward, temp_offset, temp_error, signal_offset, signal_error (metrics.py:22-30)
    np.testing.assert_allclose(col["sum_mean_reward"], train[0], err_msg="cumul_avg_reward", **tol)
    np.testing.assert_allclose(col["sum_mean_temp_offset"], train[1], **tol)
    np.testing.assert_allclose(col["sum_mean_temp_error"], train[2], **tol)
    np.testing.assert_allclose(col["sum_signal_offset"] / n, train[3], **tol)
    np.testing.assert_allclose(col["sum_signal_error"] / n, train[4], **tol)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", gu.names())
def test_device_metrics_match_reference_accumulators(name, precision):
    """the on-device accumulators after replaying a reference trace == the reference's own deploy-loop
    accumulators and training Metrics on that trace.  fp32 configurations that the pipelined kernel takes keep it with
    metrics enabled (accumulators are an epilogue of that kernel too)."""
    g = gu.Golden(name)
    env = _env_from_golden(g, precision)
    plain = env.launch_geometry()["kernel"]
    env.enable_metrics()
    assert env.launch_geometry()["kernel"] == plain   # metrics do not change which kernel runs
    env.precompute()
    for t in range(g.steps):
        ids = g.interp_ids[t][None] if g.interp_ids[t][0] >= 0 else None
        comm = None if g.comm_t is None else g.comm_t[t][None]
        env.step_tensor(g.actions[t][None], od_noise=g.od_noise[t:t + 1], signal_noise=g.sig_noise[t:t + 1],
                        interp_ids=ids, msg_keep=g.msg_keep[t][None], comm=comm)
    check_metrics_against_reference(g, env.metrics[0].cpu().numpy(), precision)
