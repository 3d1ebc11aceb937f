"""device-resident rollout collection -- policy -> ONE sampling kernel -> step kernel writing the next
observation straight into the rollout storage, captured as a CUDA graph; transitions in the layout agents/ppo.py:92-107
builds, pinned to a batch recorded from the reference's unmodified train_ppo loop (oracle/make_ppo_golden.py)."""
import os

import numpy as np
import pytest

import golden_util as gu

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


@pytest.mark.parametrize("use_graph,claimed", [(True, False), (False, False), (True, True)])
def test_collector_matches_manual_stepping(use_graph, claimed):
    """`claimed`: the collector's env runs the pipelined kernel with tiles claimed in order (one persistent CTA walks all
    16 tiles: grid barrier, claim ring and workspace reuse inside a replayed CUDA graph); the twin keeps the strided lists."""
    import torch
    import mdr_b200
    torch.manual_seed(1)
    env, twin = _env(), _env()
    if claimed:
        env.set_launch_options(max_ctas=1)
    f = env.n_features
    actor = mdr_b200.ActorMLP(f, 2, [100, 100]).cuda()   # agents/network.py:14-33
    col = mdr_b200.DeviceRolloutCollector(env, n_steps=12, seed=7, use_graph=use_graph)
    out = col.collect(actor, reset=True)
    assert out["state"].shape == (12, 64, 50, f) and out["next_state"].shape == (12, 64, 50, f)
    assert out["action"].dtype == torch.uint8 and out["done"].sum() == 0
    assert out["next_state"].data_ptr() == col.states[1].data_ptr()  # zero copy view
    first_actions = out["action"].clone()
    # replay the recorded actions on a twin env (same Philox counters): identical observations and rewards, bit for bit
    obs = twin.reset_tensor().clone()
    assert torch.equal(obs, out["state"][0])
    zero = torch.zeros(1, dtype=torch.int64, device="cuda")
    for t in range(12):
        probs = actor(out["state"][t].reshape(-1, f))
        chosen = probs.gather(1, out["action"][t].reshape(-1, 1).long()).reshape(64, 50)
        torch.testing.assert_close(chosen, out["a_log_prob"][t], rtol=1e-5, atol=1e-6)
        o, r, p, s = twin.step_tensor(out["action"][t], step_counter=zero)
        assert torch.equal(o, out["next_state"][t]) and torch.equal(r, out["reward"][t])
        assert torch.equal(p, out["cluster_hvac_power"][t]) and torch.equal(s, out["reg_signal"][t])
    # a second rollout continues the episode, with fresh draws (the device counter advanced)
    out2 = col.collect(actor)
    assert torch.equal(out2["state"][0], twin.obs)
    assert int(env.t_epoch[0]) == int(twin.t_epoch[0]) + 12 * 4
    assert not torch.equal(out2["action"], first_actions)
    assert env.step_index == 24


def test_sampling_kernel_is_categorical():
    """mdr_sample_actions == Categorical(probs).sample() in distribution; chosen_prob = probs[row, action]; A = 2 and A = 5."""
    import ctypes as C
    import torch
    import mdr_b200
    lib = mdr_b200.load_library()
    m = 400_000
    for a in (2, 5):
        base = torch.tensor([0.1, 0.9] if a == 2 else [0.05, 0.4, 0.25, 0.2, 0.1], device="cuda")
        probs = (base * 3.0).expand(m, a).contiguous()      # un-normalised rows: Categorical divides by the sum
        actions = torch.empty(m, dtype=torch.uint8, device="cuda")
        chosen = torch.empty(m, dtype=torch.float32, device="cuda")
        vp = lambda t: C.c_void_p(t.data_ptr())
        st = lib.mdr_sample_actions(vp(probs), m, a, 5, 9, None, vp(actions), vp(chosen), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert st == 0
        freq = torch.bincount(actions.long(), minlength=a).double() / m
        assert float((freq - base.double()).abs().max()) < 4e-3
        assert torch.equal(chosen, probs.gather(1, actions.long()[:, None]).squeeze(1))
        again = torch.empty_like(actions)
        lib.mdr_sample_actions(vp(probs), m, a, 5, 9, None, vp(again), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert torch.equal(again, actions)                  # counter-based: same (seed, draw index) -> same draw
        lib.mdr_sample_actions(vp(probs), m, a, 5, 10, None, vp(again), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert not torch.equal(again, actions)


def test_ppo_batch_equals_reference_train_loop_batch():
    """The batch `PPO.update` builds (agents/ppo.py:92-107) from the buffers the unmodified train_ppo loop filled
    (train_ppo.py:62-116), reproduced by the collector from the same start state, forced decisions and replayed draws."""
    import json
    import torch
    import mdr_b200
    z = np.load(os.path.join(gu.GOLDEN_DIR, "mc_ppo_rollout.npz"))
    cfg = gu.fix_config(json.loads(str(z["config_json"])))
    steps = int(z["steps"])
    snap = {k[5:]: z[k] for k in z.files if k.startswith("snap_")}
    pop = {k: (np.asarray(v)[None] if np.ndim(v) >= 1 else np.asarray(v).reshape(1)) for k, v in snap.items()}
    pop["perlin_seed"] = np.zeros(1)
    n = z["forced_actions"].shape[1]
    for precision, tol in (("fp64", dict(rtol=0, atol=1e-6)), ("fp32", dict(rtol=1e-4, atol=2e-4))):
        env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision)
        col = mdr_b200.DeviceRolloutCollector(env, n_steps=steps, episode_steps=steps)
        env.step_index = 1   # not a fresh env: the first observation is that of the loaded state (no reset step)
        kw = lambda t: dict(od_noise=z["od_noise"][t:t + 1], signal_noise=z["sig_noise"][t:t + 1])
        col.collect(forced=(z["forced_actions"], z["forced_probs"]), step_kwargs=kw)
        b = col.ppo_batch()
        assert b["state"].shape == z["state"].shape == (n * steps, env.n_features)
        np.testing.assert_allclose(b["state"].cpu().numpy(), z["state"], **tol)          # the reference stores fp32
        np.testing.assert_allclose(b["next_state"].cpu().numpy(), z["next_state"], **tol)
        assert np.array_equal(b["action"].cpu().numpy(), z["action"])
        np.testing.assert_allclose(b["old_action_log_prob"].cpu().numpy(), z["old_action_log_prob"], rtol=1e-6, atol=0)
        np.testing.assert_allclose(b["reward"].cpu().numpy(), z["reward"], **tol)
        assert np.array_equal(b["done"].cpu().numpy(), z["done"])


def test_output_tensor_validation():
    import torch
    env = _env(n_envs=4)
    env.reset_tensor()
    act = torch.zeros(4, 50, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        env.step_tensor(act, obs_out=torch.empty(4, 50, env.n_features, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        env.step_tensor(act, obs_out=torch.empty(4, 50, env.n_features + 1, device="cuda")[..., :-1])


def test_unaligned_step_stride_is_padded():
    """1 x 50 x 51 fp32 rows are 10 200 bytes: the storage pads the per-step stride to 16 bytes."""
    import torch
    import mdr_b200
    env = _env(n_envs=1)
    col = mdr_b200.DeviceRolloutCollector(env, n_steps=5, use_graph=False)
    assert all(col.states[t].data_ptr() % 16 == 0 and col.states[t].is_contiguous() for t in range(6))
    actor = mdr_b200.ActorMLP(env.n_features, 2, [16]).cuda()
    out = col.collect(actor, reset=True)
    assert torch.isfinite(out["next_state"]).all()
