"""Pins the numpy oracle (oracle/mdr_oracle.py) against the reference:
(1) known-answer vectors produced by the reference,
(2) golden traces recorded from the unmodified reference (tests/golden, oracle/make_golden.py),
(3) scipy's interpn for the interpolation restatement."""
import datetime as dt
import os

import numpy as np
import pytest

import golden_util as gu
from oracle import mdr_oracle as orc


# ---------------------------------------------------------------- known answers
DEF = dict(ua=2.18e02, cm=3.45e06, ca=9.08e05, hm=2.84e03)


def _three_steps(q_a, t_od=30.0):
    ta, tm = np.float64(20.0), np.float64(20.0)
    out = []
    for _ in range(3):
        ta, tm = orc.etp_update(ta, tm, t_od, q_a, DEF["ua"], DEF["ca"], DEF["hm"], DEF["cm"], 4)
        out.append((float(ta), float(tm)))
    return out


def test_kat_hvac_off():
    got = _three_steps(0.0)
    exp_a = [20.009539192872296, 20.018951131080087, 20.028237900639283]
    exp_m = [20.000015723031936, 20.00006254307698, 20.00013994301952]
    for (a, m), ea, em in zip(got, exp_a, exp_m):
        assert abs(a - ea) < 1e-12 and abs(m - em) < 1e-12


def test_kat_hvac_on():
    got = _three_steps(-15000 / 1.35)
    exp_a = [19.960919453461884, 19.922360248918494, 19.884313840296272]
    exp_m = [19.99993558525466, 19.99974377102353, 19.999426675846053]
    for (a, m), ea, em in zip(got, exp_a, exp_m):
        assert abs(a - ea) < 1e-12 and abs(m - em) < 1e-12


def test_kat_solar():
    g = orc.house_solar_gain(dt.datetime(2021, 6, 15, 12, 0), 7.175, 0.67)
    assert g == pytest.approx(915.336717041842, abs=1e-9)
    ta, tm = orc.etp_update(np.float64(20), np.float64(20), 30.0, g, DEF["ua"], DEF["ca"], DEF["hm"], DEF["cm"], 4)
    assert abs(ta - 20.013544501811282) < 1e-12 and abs(tm - 20.000022324806423) < 1e-12
    assert orc.house_solar_gain(dt.datetime(2021, 6, 15, 7, 29), 7.175, 0.67) == 0.0
    assert orc.house_solar_gain(dt.datetime(2021, 6, 15, 7, 31), 7.175, 0.67) == pytest.approx(1044.0256593371194, abs=1e-9)
    assert orc.house_solar_gain(dt.datetime(2021, 12, 21, 12, 30), 7.175, 0.67) == pytest.approx(1161.7031020346528, abs=1e-9)
    # reference unit test window (env/unit_tests_MA_DemandResponse.py:113-128): 0 before 7:30 / after 17:30
    assert orc.house_solar_gain(dt.datetime(2021, 3, 1, 17, 31), 7.175, 0.67) == 0.0
    assert orc.house_solar_gain(dt.datetime(2021, 3, 1, 12, 0), 7.175, 0.67) > 0


def test_kat_hvac_q_and_power():
    # env/unit_tests_MA_DemandResponse.py:36-44
    assert -1 * 15000 / (1 + 0.35) == -15000 / 1.35
    assert 15000 / 2.5 == 6000


def test_kat_lockout_trace():
    """SURVEY 8c lockout trace (lockout 40 s, dt 4): cmd -> on, lockout, sso."""
    cmds = [1, 1, 0] + [1] * 10 + [1, 0] + [0] * 10 + [1]
    exp = [(1, 0, 0), (1, 0, 0), (0, 1, 0)]
    exp += [(0, 1, 4 * k) for k in range(1, 10)]
    exp += [(1, 0, 0), (1, 0, 0), (0, 1, 0)]
    exp += [(0, 1, 4 * k) for k in range(1, 10)]
    exp += [(0, 0, 40), (1, 0, 0)]
    on, lock, sso, dur = np.array([0]), np.array([0]), np.array([40]), np.array([40])
    got = []
    for c in cmds:
        on, lock, sso = orc.hvac_step(on, lock, sso, dur, np.array([c]), 4)
        got.append((int(on[0]), int(lock[0]), int(sso[0])))
    assert got == exp


def test_kat_lockout_reference_unit_test():
    """env/unit_tests_MA_DemandResponse.py:46-77: lockout 12 s, dt 4 s."""
    on, lock, sso, dur = np.array([0]), np.array([0]), np.array([12]), np.array([12])
    seq = [(1, (1, 0, 0)), (0, (0, 1, 0)), (1, (0, 1, 4)), (1, (0, 1, 8)), (1, (1, 0, 0))]
    for c, e in seq:
        on, lock, sso = orc.hvac_step(on, lock, sso, dur, np.array([c]), 4)
        assert (int(on[0]), int(lock[0]), int(sso[0])) == e


def test_kat_deadband():
    assert float(orc.deadband_l2(20, 0, 21)) == 1.0
    assert float(orc.deadband_l2(7500, 0, 5625)) == 3515625.0
    assert float(orc.deadband_l2(20, 2, 20.5)) == 0.0
    assert float(orc.deadband_l2(20, 2, 18)) == 1.0


def test_kat_neighbour_tables():
    """SURVEY appendix A.3 (probed on the reference)."""
    t = orc.comm_links("neighbours", 12, 10)
    assert t[0].tolist() == [7, 8, 9, 10, 11, 1, 2, 3, 4, 5]
    assert t[11].tolist() == [6, 7, 8, 9, 10, 0, 1, 2, 3, 4]
    t = orc.comm_links("neighbours", 12, 3)
    assert t[0].tolist() == [11, 1, 2] and t[5].tolist() == [4, 6, 7]
    assert orc.comm_links("neighbours", 5, 10)[0].tolist() == [3, 4, 1, 2]
    assert orc.comm_links("neighbours", 37, 10)[0].tolist() == [32, 33, 34, 35, 36, 1, 2, 3, 4, 5]
    t = orc.comm_links("closed_groups", 12, 3)
    assert t[0].tolist() == [1, 2, 3] and t[3].tolist() == [0, 1, 2] and t[4].tolist() == [5, 6, 7] and t[11].tolist() == [8, 9, 10]
    t = orc.comm_links("closed_groups", 10, 3)
    assert t[8].tolist() == [6, 7, 9] and t[9].tolist() == [6, 7, 8]
    t = orc.comm_links("neighbours_2D", 25, 10, row_size=5, distance_comm=2)
    assert t[0].tolist() == [3, 24, 4, 9, 15, 20, 5, 10, 21, 1, 6, 2]
    assert orc.comm_links("no_message", 7, 10).shape == (7, 0)
    with pytest.raises(ValueError):
        orc.comm_links("bogus", 7, 10)
    with pytest.raises(ValueError):
        orc.comm_links("neighbours_2D", 24, 10, row_size=5)


def test_kat_interpolation():
    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS)
    point = {"Ua_ratio": 1.04, "Cm_ratio": 0.96, "Ca_ratio": 1.0, "Hm_ratio": 1.05, "air_temp": 0.37,
             "mass_temp": -1.3, "OD_temp": 9.71, "HVAC_power": 15000, "hour": 40000, "date": 100}
    assert interp.fast(point) == pytest.approx(4222.708561747741, abs=1e-9)


def test_interpolation_vs_scipy():
    from scipy.interpolate import interpn

    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS)
    rng = np.random.default_rng(3)
    grid = [np.asarray(gu.INTERP_GRID[k], float) for k in gu.INTERP_KEYS]
    for trial in range(40):
        point = {}
        for k, g in zip(gu.INTERP_KEYS, grid):
            v = rng.uniform(g.min() - 0.2 * (g.max() - g.min()), g.max() + 0.2 * (g.max() - g.min()))
            if trial % 5 == 0:  # exactly on grid points, incl. the last one
                v = g[rng.integers(len(g))]
            point[k] = float(v)
        point = interp.clip(point)
        c = [point[k] for k in gu.INTERP_KEYS]
        near = [int(np.argmin(np.abs(grid[i] - c[i]))) for i in range(4)]
        ih = int(np.argmin(np.abs(grid[7] - c[7])))
        sub = interp.values[near[0], near[1], near[2], near[3]][:, :, :, ih, :, :]
        exp = interpn([grid[4], grid[5], grid[6], grid[8], grid[9]], sub, [c[4], c[5], c[6], c[8], c[9]])[0]
        assert interp.fast(point) == exp  # bit-exact restatement of scipy's linear path


def test_grid_point_identity():
    """monteCarlo/unit_tests_interp.py:74-98: at grid points the interpolation returns the table entry."""
    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS)
    rng = np.random.default_rng(4)
    for _ in range(10):
        idx = [int(rng.integers(len(gu.INTERP_GRID[k]))) for k in gu.INTERP_KEYS]
        point = {k: float(gu.INTERP_GRID[k][i]) for k, i in zip(gu.INTERP_KEYS, idx)}
        assert interp.fast(point) == pytest.approx(interp.values[tuple(idx)], rel=1e-13)


# ---------------------------------------------------------------- golden traces
@pytest.mark.parametrize("name", gu.names())
def test_oracle_matches_reference_trace(name):
    g = gu.Golden(name)
    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS) if g.uses_interp else None
    env = orc.OracleEnv(g.config, g.batched_snap(), comm_table=g.comm, interp=interp)
    # initial observation
    init_comm = g.z.get("init_comm")
    obs0 = env.obs(msg_keep=g.init_keep[None], comm=None if init_comm is None else init_comm[None])
    np.testing.assert_allclose(obs0[0], g.obs0, rtol=0, atol=1e-12)
    ci, oi = 0, 0
    for t in range(g.steps):
        ids = g.interp_ids[t][None] if g.interp_ids[t][0] >= 0 else None
        comm = None if g.comm_t is None else g.comm_t[t][None]
        obs, rew, p, s = env.step(g.actions[t][None], g.od_noise[t:t + 1], g.sig_noise[t:t + 1], ids,
                                  g.msg_keep[t][None], comm)
        assert p[0] == g.power[t]
        assert abs(s[0] - g.signal[t]) <= 1e-9 * max(1.0, abs(g.signal[t])), (t, s[0], g.signal[t])
        assert abs(env.s["od_temp"][0] - g.od_temp[t]) < 1e-12
        assert abs(env.s["solar_gain"][0] - g.solar[t]) < 1e-9
        if t in g.check_steps:
            assert np.array_equal(env.s["on"][0], g.on[ci])
            assert np.array_equal(env.s["lockout"][0], g.lockout[ci])
            assert np.array_equal(env.s["sso"][0], g.sso[ci])
            np.testing.assert_allclose(env.s["t_air"][0], g.t_air[ci], rtol=0, atol=1e-10)
            np.testing.assert_allclose(env.s["t_mass"][0], g.t_mass[ci], rtol=0, atol=1e-10)
            np.testing.assert_allclose(rew[0], g.reward[ci], rtol=1e-10, atol=1e-10)
            ci += 1
        if t in g.obs_steps:
            np.testing.assert_allclose(obs[0], g.obs[oi], rtol=0, atol=1e-10)
            oi += 1
    assert ci == len(g.check_steps) and oi == len(g.obs_steps)


@pytest.mark.parametrize("name", ["interp_150_sinus_solar", "interp_40_perlin", "c1_1000_fp64", "c0_bangbang_50"])
def test_oracle_initial_grid_step(name):
    """PowerGrid.step(start_datetime) in build_environment (:133): from signal 0 and
    time_since_last_interp = period + 1 the oracle must land on the reference's snapshot."""
    g = gu.Golden(name)
    snap = g.batched_snap()
    period = g.config["default_env_prop"]["power_grid_prop"]["base_power_parameters"]["interpolation"]["interp_update_period"]
    snap["signal"], snap["base_power"] = np.zeros(1), np.zeros(1)
    snap["time_since_interp"] = np.array([period + 1])
    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS) if g.uses_interp else None
    env = orc.OracleEnv(g.config, snap, comm_table=g.comm, interp=interp)
    ids = g.init_interp_ids if g.init_interp_ids[0] >= 0 else None
    sig = env.grid_step(0, orc.to_datetime(env.s["t_epoch"][0]), float(g.init_sig_noise), ids)
    assert sig == pytest.approx(float(g.snap["signal"]), rel=1e-14)
    assert env.s["base_power"][0] == pytest.approx(float(g.snap["base_power"]), rel=1e-14)


@pytest.mark.parametrize("name", ["c0_bangbang_50", "tiny_3_comm_clipped", "interp_40_perlin", "c1_1000_fp64"])
def test_scalar_port_matches_reference_trace(name):
    """The per-object python port used as the CPU baseline walks the reference trajectory (incl. the interpolated base
    power with sampled houses)."""
    from oracle import mdr_oracle_scalar as sc
    g = gu.Golden(name)
    interp = orc.PowerInterp(gu.synthetic_table(), gu.INTERP_GRID, gu.INTERP_KEYS) if g.uses_interp else None
    env = sc.ScalarEnv(g.config, g.snap, interp)
    ci = oi = 0
    for t in range(g.steps):
        ids = g.interp_ids[t] if g.interp_ids[t][0] >= 0 else None
        obs, rew, done, info = env.step({i: bool(g.actions[t][i]) for i in range(g.n)}, g.od_noise[t], g.sig_noise[t], ids)
        assert info["cluster_hvac_power"] == g.power[t]
        assert env.signal == pytest.approx(g.signal[t], rel=1e-13)
        if t in g.check_steps:
            np.testing.assert_allclose([env.houses[i].t_air for i in range(g.n)], g.t_air[ci], rtol=0, atol=1e-10)
            assert [int(env.houses[i].hvac.seconds_since_off) for i in range(g.n)] == list(g.sso[ci])
            np.testing.assert_allclose([rew[i] for i in range(g.n)], g.reward[ci], rtol=1e-10, atol=1e-10)
            ci += 1
        if t in g.obs_steps:
            np.testing.assert_allclose(np.array([env.norm_state(obs[i]) for i in range(g.n)]), g.obs[oi], rtol=0, atol=1e-10)
            oi += 1


def test_greedy_myopic_oracle_matches_reference_controller():
    """oracle.greedy_myopic_actions vs decisions recorded from the unmodified agents/greedy_myopic_controller.py
    (oracle/make_greedy_golden.py): 24 random clusters, incl. signal = 0 and signal above the cluster's maximum."""
    g = np.load(os.path.join(gu.GOLDEN_DIR, "mc_greedy_myopic.npz"))
    for k in range(int(g["n_cases"])):
        c = {name: g["%d_%s" % (k, name)] for name in ("t_air", "target", "cap", "cop", "lockout", "signal", "action")}
        act = orc.greedy_myopic_actions(c["t_air"][None], c["target"][None], (c["cap"] / c["cop"])[None], c["lockout"][None],
                                        np.array([float(c["signal"])]))
        assert np.array_equal(act[0], c["action"]), k
