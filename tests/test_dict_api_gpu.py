"""GPU tests of the drop-in dict API (`MADemandResponseEnv(config).reset()/step(action_dict)`):
seeded like the reference (`random.seed(s)`), it must walk the same trajectory as the reference
did when the golden traces were recorded -- same population, same start date, same random
draws in the same order."""
import copy
import random
import warnings

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

REFERENCE_KEYS = ['OD_temp', 'datetime', 'house_temp', 'house_mass_temp', 'hvac_turned_on', 'hvac_seconds_since_off',
                  'hvac_lockout', 'house_target_temp', 'house_deadband', 'house_Ua', 'house_Cm', 'house_Ca', 'house_Hm',
                  'house_solar_gain', 'hvac_COP', 'hvac_cooling_capacity', 'hvac_latent_cooling_fraction',
                  'hvac_lockout_duration', 'message', 'reg_signal', 'cluster_hvac_power']


@pytest.mark.parametrize("name", ["c0_bangbang_50", "hetero_37_solar_lockout", "interp_150_sinus_solar",
                                  "modes_12_allflags_defect_max", "random_fixed_20", "random_sample_15", "n2d_25",
                                  "tiny_3_comm_clipped", "no_message_5_small", "modes_10_closed_mixture_flat"])
def test_seeded_dict_env_walks_the_reference_trajectory(name, monkeypatch):
    import mdr_b200
    from oracle import mdr_oracle as orc
    g = gu.Golden(name)
    random.seed(g.seed)
    # the golden generator's recorder replaced np.random.rand before the env was built: mirror it
    rng = np.random.default_rng(g.seed + 1000)
    monkeypatch.setattr(np.random, "rand", lambda *a: rng.random())
    env = mdr_b200.MADemandResponseEnv(g.config, interp_table=gu.synthetic_table() if g.uses_interp else None)
    obs = env.reset()
    assert list(obs.keys()) == list(range(g.n)) and list(obs[0].keys()) == REFERENCE_KEYS
    assert env.nb_agents == g.n and env.agent_ids == list(range(g.n)) and list(env.cluster.houses.keys()) == env.agent_ids
    assert orc.from_datetime(env.datetime) == int(g.snap["t_epoch"])
    np.testing.assert_allclose([obs[i]["house_temp"] for i in range(g.n)], g.snap["t_air"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(obs[0]["reg_signal"], float(g.snap["signal"]), rtol=1e-12, atol=1e-9)
    # the golden generator rebuilt the initial observation once more: mirror its draws
    env._message_draws()
    ci = 0
    for t in range(g.steps):
        act = {i: bool(g.actions[t][i]) for i in range(g.n)}
        obs, rew, done, info = env.step(act)
        assert info["cluster_hvac_power"] == g.power[t]
        np.testing.assert_allclose(obs[0]["reg_signal"], g.signal[t], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(obs[0]["OD_temp"], g.od_temp[t], rtol=0, atol=1e-12)
        assert all(v is False for v in done.values())
        if t in g.check_steps:
            assert [obs[i]["hvac_turned_on"] for i in range(g.n)] == [bool(x) for x in g.on[ci]]
            assert [obs[i]["hvac_lockout"] for i in range(g.n)] == [bool(x) for x in g.lockout[ci]]
            assert [obs[i]["hvac_seconds_since_off"] for i in range(g.n)] == [int(x) for x in g.sso[ci]]
            np.testing.assert_allclose([obs[i]["house_temp"] for i in range(g.n)], g.t_air[ci], rtol=0, atol=1e-9)
            np.testing.assert_allclose([obs[i]["house_mass_temp"] for i in range(g.n)], g.t_mass[ci], rtol=0, atol=1e-9)
            np.testing.assert_allclose([rew[i] for i in range(g.n)], g.reward[ci], rtol=0, atol=1e-9)
            ci += 1
        if t in g.obs_steps:
            oi = g.obs_steps.index(t)
            np.testing.assert_allclose(env.obs_tensor().cpu().numpy(), g.obs[oi], rtol=0, atol=1e-9)
    assert len(obs[0]["message"]) == g.msg_keep.shape[2]


def test_missing_action_warns_and_defaults_to_off():
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = 6
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "constant"
    random.seed(3)
    env = mdr_b200.MADemandResponseEnv(cfg)
    env.reset()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        obs, rew, done, info = env.step({i: True for i in range(5)})  # agent 5 missing
    assert any("did not receive any command" in str(x.message) for x in w)
    assert obs[5]["hvac_turned_on"] is False and obs[0]["hvac_turned_on"] is True
    assert info["cluster_hvac_power"] == 5 * 6000.0


def test_bangbang_controller_and_deepcopy():
    """agents/bangbang_controllers.py:41-61 restated inline: the drop-in env must serve it, and
    copy.deepcopy(env) (used by every utils.test_*_agent) must give an independent twin."""
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = 10
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "constant"
    random.seed(11)
    env = mdr_b200.MADemandResponseEnv(cfg)
    obs = env.reset()
    for _ in range(30):
        obs, rew, _, _ = env.step({i: obs[i]["house_temp"] > obs[i]["house_target_temp"] for i in obs})
    twin = copy.deepcopy(env)
    st = random.getstate()
    a = env.step({i: True for i in obs})
    random.setstate(st)
    b = twin.step({i: True for i in obs})
    assert a[1] == b[1] and a[3] == b[3]
    assert [a[0][i]["house_temp"] for i in obs] == [b[0][i]["house_temp"] for i in obs]
    env.step({i: False for i in obs})
    assert twin.datetime != env.datetime


def test_default_config_needs_the_missing_table():
    """The shipped default (base_power_mode='interpolation') points at a blob the reference does not
    ship: the reference raises FileNotFoundError from np.load; so do we."""
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = 4
    with pytest.raises(FileNotFoundError):
        mdr_b200.MADemandResponseEnv(cfg)
