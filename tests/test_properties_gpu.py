"""Full-size, size-independent property checks on the GPU (BASELINE configs c2 / c3 / c4 shapes):
determinism, shard independence, the power checksum, the integer state machine restated with torch
ops as the checker, and structural checks of the observation tensor."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(n_envs, n, interp=False, precision="fp32", seed=5, action_source="array", **kw):
    import mdr_b200
    import golden_util as gu
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if interp else "constant"
    ep["power_grid_prop"]["signal_mode"] = "sinusoidals"
    cfg["default_house_prop"]["solar_gain_bool"] = False
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=seed, action_source=action_source,
                                        interp_table=gu.synthetic_table() if interp else None, **kw)
    return cfg, flat, pop, env


@pytest.mark.parametrize("n_envs,n,interp", [(4096, 50, False), (16384, 100, True), (1000, 1000, False)])
def test_full_size_invariants(n_envs, n, interp):
    import torch
    cfg, flat, pop, env = _make(n_envs, n, interp)
    obs0 = env.reset_tensor().clone()
    assert torch.isfinite(obs0).all()
    g = torch.Generator(device="cuda").manual_seed(1)
    steps = 80 if interp else 12
    dt, c = flat.time_step, flat.n_comm
    for t in range(steps):
        act = (torch.rand(n_envs, n, device="cuda", generator=g) < 0.5).to(torch.uint8)
        hv_before, lockdur = env.hvac.clone(), env.lockout_dur
        obs, rew, p, s = env.step_tensor(act)
        # integer lockout state machine (env/MA_DemandResponse.py:463-492) restated with torch ops
        on, sso = hv_before & 1, hv_before >> 2
        sso = torch.where(on == 0, sso + dt, sso)
        lock = ~((on != 0) | (sso >= lockdur))
        new_on = torch.where(lock, torch.zeros_like(act, dtype=torch.bool), act != 0)
        sso = torch.where(~lock & new_on, torch.zeros_like(sso), sso)
        lock = lock | (~lock & ~new_on & (sso + dt < lockdur))
        expect = (sso << 2) | (lock.int() << 1) | new_on.int()
        assert torch.equal(env.hvac, expect), t
        # power checksum: exact (integer-valued addends)
        p_on = env.coef_b[..., 2].double()
        assert torch.equal(p, (p_on * new_on).sum(1)), t
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(s).all()
    # observation structure, utils.normStateDict order
    ta, tm = env.t_air, env.t_mass
    torch.testing.assert_close(obs[..., 0], (ta - 20) * 0.2, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(obs[..., 1], (tm - 20) * 0.2, rtol=1e-6, atol=1e-6)
    assert torch.equal(obs[..., 5], (env.hvac & 1).float()) and torch.equal(obs[..., 6], ((env.hvac >> 1) & 1).float())
    assert (obs[..., 8] == 1).all()
    torch.testing.assert_close(obs[..., 9].double(), (s / (7500.0 * n))[:, None].expand(-1, n), rtol=1e-6, atol=1e-7)
    # neighbour gather: message k of house i is the own temperature difference of house comm[i, k]
    import mdr_b200
    table = torch.as_tensor(mdr_b200.comm_table("neighbours", n, 10)).cuda().long()
    dT = ((ta - env.coef_b[..., 3]) * 0.2)
    msg_dt = obs[..., 11::4][..., :c]
    torch.testing.assert_close(msg_dt, dT[:, table], rtol=1e-6, atol=1e-6)
    msg_p = obs[..., 13::4][..., :c]
    cur = (env.coef_b[..., 2] * (env.hvac & 1)) / 7500
    torch.testing.assert_close(msg_p, cur[:, table], rtol=1e-6, atol=1e-6)


def test_determinism_and_shard_independence():
    """Same inputs -> identical bits; env e of a big batch == the same env stepped alone (different CTA
    packing), which is what makes sharding over GPUs exact."""
    import torch
    import mdr_b200
    cfg, flat, pop, env = _make(257, 50, False, seed=9)
    _, _, _, env2 = _make(257, 50, False, seed=9)
    sub_ids = [0, 3, 128, 256]
    subs = []
    for e in sub_ids:
        one = {k: np.asarray(v)[e:e + 1] for k, v in pop.items()}
        subs.append(mdr_b200.VecDemandResponseEnv(cfg, one, precision="fp32", seed=9))
    for x in [env, env2] + subs:
        x.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(2)
    for t in range(20):
        act = (torch.rand(257, 50, device="cuda", generator=g) < 0.5).to(torch.uint8)
        noise = torch.randn(257, device="cuda", generator=g, dtype=torch.float64) * 0.5
        o1 = [x.clone() for x in env.step_tensor(act, od_noise=noise)]
        o2 = env2.step_tensor(act, od_noise=noise)
        for a, b in zip(o1, o2):
            assert torch.equal(a, b)
        for e, sub in zip(sub_ids, subs):
            so = sub.step_tensor(act[e:e + 1].contiguous(), od_noise=noise[e:e + 1].contiguous())
            for a, b in zip(o1, so):
                assert torch.equal(a[e:e + 1], b), (t, e)


@pytest.mark.parametrize("window", ["table", True])
def test_l2_residency_window_changes_nothing(window):
    """The optional persisting-L2 access-policy window (over the interpolation table, or over the state arena) is a
    cache hint: identical bits with and without it."""
    import torch
    _, _, _, a = _make(600, 100, True, seed=11)
    _, _, _, b = _make(600, 100, True, seed=11, l2_persist=window)
    for x in (a, b):
        x.reset_tensor()
        x.stagger_interp_clock(seed=4)
    g = torch.Generator(device="cuda").manual_seed(6)
    for t in range(80):
        act = (torch.rand(600, 100, device="cuda", generator=g) < 0.5).to(torch.uint8)
        oa, ob = a.step_tensor(act), b.step_tensor(act)
        for x, y in zip(oa, ob):
            assert torch.equal(x, y), t
    assert torch.equal(a.temps, b.temps) and torch.equal(a.env["base_power"], b.env["base_power"])


def test_on_device_bangbang_million_houses():
    """c3 shape: 1000 envs x 1000 houses, on-device bang-bang, no observation written."""
    import torch
    cfg, flat, pop, env = _make(1000, 1000, False, action_source="bangbang", with_obs=False)
    env.reset_tensor()
    env.run(200)
    torch.cuda.synchronize()
    ta = env.t_air
    assert torch.isfinite(ta).all()
    # a bang-bang cluster must hold its houses near the target once the initial offset is cooled away
    target = env.coef_b[..., 3]
    assert float((ta - target).mean()) < float(torch.as_tensor(pop["t_air"] - pop["target"]).mean())
    assert int(env.t_epoch[0]) == int(pop["t_epoch"][0]) + 200 * 4


def test_production_mode_device_rng_is_reproducible_and_plausible():
    import torch
    cfg, flat, pop, env = _make(512, 50, False, seed=4)
    _, _, _, twin = _make(512, 50, False, seed=4)
    _, _, _, other = _make(512, 50, False, seed=5)
    for x in (env, twin, other):
        x.reset_tensor()
    act = torch.ones(512, 50, dtype=torch.uint8, device="cuda")
    od = []
    for t in range(50):
        env.step_tensor(act)
        twin.step_tensor(act)
        other.step_tensor(act)
        od.append(env.env["od_temp"].clone())
    assert torch.equal(env.env["od_temp"], twin.env["od_temp"]) and torch.equal(env.temps, twin.temps)
    assert not torch.equal(env.env["od_temp"], other.env["od_temp"])
    od = torch.stack(od)  # [T, E]; default temp_std 0.5 around a 28..34 sinusoid
    assert 25.0 < float(od.min()) and float(od.max()) < 37.5
    resid = od[1:] - od[:-1]
    assert 0.4 < float(resid.std()) < 1.0  # difference of two N(0, 0.5) draws: std ~0.71


def test_device_perlin_signal_statistics():
    """Production-mode grid signal (device perlin: utils.Perlin over hashed lattice gradients, parity of the value itself
    is unpinned -- DESIGN.md section 4): bounded, centred on the base power, smooth in time, different per env."""
    import torch
    n_envs, n, steps = 512, 10, 400
    import mdr_b200
    cfg, _, pop, _ = _make(n_envs, n, with_obs=False)
    cfg["default_env_prop"]["power_grid_prop"]["signal_mode"] = "perlin"
    flat = mdr_b200.FlatConfig(cfg)
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", seed=11, with_obs=False)
    env.reset_tensor()
    act = torch.zeros(n_envs, n, dtype=torch.uint8, device="cuda")
    sig = torch.empty(steps, n_envs, dtype=torch.float64, device="cuda")
    for t in range(steps):
        sig[t] = env.step_tensor(act)[3]
    base = flat.avg_power_per_hvac * n
    ratio = env.env["artificial_ratio"]
    noise = sig / (base * ratio) - 1.0          # = amplitude * perlin(t), clipped at -1 and by max_power
    amp = float(flat.signal_params["amplitude_ratios"])
    assert float(noise.min()) >= -1.0 - 1e-9 and float(noise.abs().max()) < 1.5 * amp
    assert abs(float(noise.mean())) < 0.03 * amp                      # centred
    assert float(noise.std()) > 0.05 * amp                            # not degenerate
    d1 = (noise[1:] - noise[:-1]).abs().mean()
    assert float(d1) < 0.25 * float(noise.std())                      # smooth: small step-to-step change
    c = torch.corrcoef(noise[:, :64].T)
    off = c - torch.eye(64, device="cuda", dtype=c.dtype)
    assert float(off.abs().mean()) < 0.2                              # envs are decorrelated (own perlin seed)
