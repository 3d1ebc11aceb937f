"""Drop-in proof with the reference's REAL classes (north star: "agents and controllers drop in unchanged").
`tests/golden/dropin_c0.pkl` was recorded on the B200 box by tools/record_dropin_fixture.py: the obs dicts the drop-in
`MADemandResponseEnv` returned on BASELINE config 0's shape.  Here -- in the build container, where /root/reference
exists -- they are fed to the unmodified `BangBangController`, `DeadbandBangBang`, `GreedyMyopic`, `get_actions` and
`normStateDict` (imported through oracle/ref_stubs.py); skipped where the reference is absent (the GPU box)."""
import os
import pickle

import numpy as np
import pytest

from oracle import ref_stubs

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin_c0.pkl")
pytestmark = pytest.mark.skipif(not (ref_stubs.reference_available() and os.path.isfile(FIXTURE)),
                                reason="needs /root/reference and the recorded fixture")


@pytest.fixture(scope="module")
def rec():
    with open(FIXTURE, "rb") as f:
        return pickle.load(f)


@pytest.fixture(scope="module")
def ref():
    Env, norm, cfg, utils = ref_stubs.import_reference()
    import agents.bangbang_controllers as bb
    import agents.greedy_myopic_controller as gm
    return dict(norm=norm, utils=utils, bb=bb, gm=gm)


def test_unmodified_bangbang_and_get_actions_reproduce_the_recorded_actions(rec, ref):
    cfg, obs0 = rec["config"], rec["obs"][0]
    num_state = len(ref["norm"](obs0[0], cfg))
    assert num_state == rec["obs_tensor"][0].shape[1] == 51
    actors = {k: ref["bb"].BangBangController({"id": k}, cfg, num_state=num_state) for k in obs0}   # main-deploy.py:68-77
    for t, act in enumerate(rec["actions"]):
        got = ref["utils"].get_actions(actors, rec["obs"][t])                                      # utils.py:713
        assert {k: bool(v) for k, v in got.items()} == act, t
    # the deadband variant and the trivial controllers run on the same dicts (agents/bangbang_controllers.py:14-38, 64-88)
    for cls in ("DeadbandBangBangController", "BasicController", "AlwaysOnController"):
        if hasattr(ref["bb"], cls):
            ctrl = getattr(ref["bb"], cls)({"id": 3}, cfg, num_state=num_state)
            assert ctrl.act(rec["obs"][5]) in (True, False, 0, 1)


def test_unmodified_normstatedict_equals_device_observation(rec, ref):
    cfg = rec["config"]
    for t in (0, 1, 7, len(rec["obs"]) - 1):
        rows = np.stack([ref["norm"](rec["obs"][t][k], cfg) for k in sorted(rec["obs"][t])])
        np.testing.assert_allclose(rec["obs_tensor"][t], rows, rtol=0, atol=1e-9, err_msg="step %d" % t)


def test_unmodified_greedy_myopic_runs_on_the_obs_dicts(rec, ref):
    """main-deploy.py's flow with --agent GreedyMyopic: one controller object per house, get_actions on every obs dict
    (the first agent of a step sorts the pandas frame built from OUR dicts, agents/greedy_myopic_controller.py:29-49)."""
    from oracle import mdr_oracle as orc
    cfg = rec["config"]
    num_state = rec["obs_tensor"][0].shape[1]
    ref["gm"].global_myopic_memory[0] = None
    actors = {k: ref["gm"].GreedyMyopic({"id": k}, cfg, num_state=num_state) for k in rec["obs"][0]}
    for t in range(8):
        obs = rec["obs"][t]
        ids = sorted(obs)
        got = ref["utils"].get_actions(actors, obs)
        picked = np.array([bool(got[k]) for k in ids])
        col = lambda key: np.array([[obs[k][key] for k in ids]], dtype=np.float64)
        expect = orc.greedy_myopic_actions(col("house_temp"), col("house_target_temp"),
                                           col("hvac_cooling_capacity") / col("hvac_COP"), col("hvac_lockout"),
                                           np.array([obs[0]["reg_signal"]]))[0]
        assert np.array_equal(picked, expect.astype(bool)), t


def test_recorded_trajectory_is_the_reference_trajectory(rec):
    """Same seed, same config through the unmodified reference env: identical obs dict keys / values (1e-9) and rewards."""
    import random
    Env, norm, cfg0, utils = ref_stubs.import_reference()
    cfg = rec["config"]
    random.seed(1)
    env = Env(cfg)
    obs = env.reset()
    for t, act in enumerate(rec["actions"]):
        mine = rec["obs"][t]
        assert list(mine.keys()) == list(obs.keys()) and list(mine[0].keys()) == list(obs[0].keys()), t
        for k in obs:
            for key in ("house_temp", "house_mass_temp", "OD_temp", "reg_signal", "cluster_hvac_power"):
                assert abs(mine[k][key] - obs[k][key]) < 1e-9, (t, k, key)
            assert mine[k]["hvac_turned_on"] == bool(obs[k]["hvac_turned_on"]) and mine[k]["hvac_lockout"] == obs[k]["hvac_lockout"]
            assert mine[k]["hvac_seconds_since_off"] == obs[k]["hvac_seconds_since_off"] and mine[k]["datetime"] == obs[k]["datetime"]
            assert len(mine[k]["message"]) == len(obs[k]["message"])
            for a, b in zip(mine[k]["message"], obs[k]["message"]):
                assert a.keys() == b.keys() and all(abs(a[x] - b[x]) < 1e-9 for x in a)
        obs, rew, done, info = env.step(act)
        assert all(abs(rew[k] - rec["rewards"][t][k]) < 1e-9 for k in rew), t
