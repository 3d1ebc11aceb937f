"""device-resident rollout collection -- the observation is written by the kernel straight
into the rollout storage, the transitions must equal what stepping the same env with the same actions
produces, in the layout agents/ppo.py:92-107 consumes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env(seed=3, n_envs=64, n=50):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["power_grid_prop"]["base_power_mode"] = "constant"
    cfg["default_house_prop"]["solar_gain_bool"] = False
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    return mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=seed)


def test_collector_matches_manual_stepping():
    import torch
    import mdr_b200
    torch.manual_seed(1)
    env, twin = _env(), _env()
    f = env.n_features
    # the reference Actor (agents/network.py:14-33): MLP [F, 100, 100, 2] ending in a softmax
    actor = torch.nn.Sequential(torch.nn.Linear(f, 100), torch.nn.ReLU(), torch.nn.Linear(100, 100), torch.nn.ReLU(),
                                torch.nn.Linear(100, 2), torch.nn.Softmax(dim=-1)).cuda()
    col = mdr_b200.DeviceRolloutCollector(env, n_steps=12)
    gen = torch.Generator(device="cuda").manual_seed(7)
    out = col.collect(actor, generator=gen, reset=True)
    assert out["state"].shape == (12, 64, 50, f) and out["next_state"].shape == (12, 64, 50, f)
    assert out["action"].dtype == torch.uint8 and out["done"].sum() == 0
    assert out["next_state"].data_ptr() == col.states[1].data_ptr()  # zero copy view
    # replay the recorded actions on a twin env: identical observations and rewards, bit for bit
    obs = twin.reset_tensor().clone()
    assert torch.equal(obs, out["state"][0])
    for t in range(12):
        probs = actor(out["state"][t].reshape(-1, f))
        chosen = probs.gather(1, out["action"][t].reshape(-1, 1).long()).reshape(64, 50)
        torch.testing.assert_close(chosen, out["a_log_prob"][t], rtol=1e-5, atol=1e-6)
        o, r, p, s = twin.step_tensor(out["action"][t])
        assert torch.equal(o, out["next_state"][t]) and torch.equal(r, out["reward"][t])
        assert torch.equal(p, out["cluster_hvac_power"][t]) and torch.equal(s, out["reg_signal"][t])
    # a second rollout continues the episode
    out2 = col.collect(actor, generator=gen)
    assert torch.equal(out2["state"][0], twin.obs)
    assert int(env.t_epoch[0]) == int(twin.t_epoch[0]) + 12 * 4


def test_output_tensor_validation():
    import torch
    env = _env(n_envs=4)
    env.reset_tensor()
    act = torch.zeros(4, 50, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        env.step_tensor(act, obs_out=torch.empty(4, 50, env.n_features, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        env.step_tensor(act, obs_out=torch.empty(4, 50, env.n_features + 1, device="cuda")[..., :-1])
