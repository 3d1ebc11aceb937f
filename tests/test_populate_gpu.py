"""Device-side population draw and masked partial reset, through the C ABI.  Parity with the
reference's reset-time randomness is distribution-level by design (python's Mersenne Twister is not
re-implemented): the tests check supports, moments against the closed forms of utils.apply_house_noise /
apply_hvac_noise / HVAC.__init__ / get_random_date_time, exact derived quantities, determinism, and that a
masked reset leaves every byte of the other envs alone while the selected envs restart exactly as a fresh
environment built from the same population would."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(n, noise="big_noise"):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = n
    ep["power_grid_prop"]["base_power_mode"] = "constant"
    cfg["default_house_prop"]["solar_gain_bool"] = False
    cfg["noise_house_prop"]["noise_mode"] = noise
    cfg["noise_hvac_prop"]["noise_mode"] = noise
    cfg["default_hvac_prop"]["lockout_noise"] = 8
    return cfg


def _env(cfg, n_envs, precision="fp64", seed=7, **kw):
    import mdr_b200
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, n_envs, seed=seed)
    return flat, mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=seed, **kw)


def test_device_population_distributions():
    import torch
    cfg = _cfg(50)
    flat, env = _env(cfg, 4000)
    obs = env.reset_envs()
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all()
    nh = flat.noise_house["noise_parameters"][flat.noise_house["noise_mode"]]
    hd, vd = flat.house_def, flat.hvac_def
    half_normal_mean = math.sqrt(2 / math.pi)
    for vals, base, std in ((env.t_air, hd["init_air_temp"], nh["std_start_temp"]), (env.t_mass, hd["init_mass_temp"], nh["std_start_temp"]),
                            (env.raw["target"], hd["target_temp"], nh["std_target_temp"])):
        d = vals.double() - base
        assert (d >= 0).all()  # np.abs(random.gauss(0, std)), utils.py:627-635
        assert float(d.mean()) == pytest.approx(std * half_normal_mean, rel=0.02)
        assert float((d * d).mean()) == pytest.approx(std * std, rel=0.03)
    lo, hi = nh["factor_thermo_low"], nh["factor_thermo_high"]
    for key, name in (("ua", "Ua"), ("cm", "Cm"), ("ca", "Ca"), ("hm", "Hm")):
        f = env.raw[key] / hd[name]
        assert float(f.min()) >= lo and float(f.max()) <= hi
        assert float(f.mean()) == pytest.approx((lo + hi + 1.0) / 3.0, rel=0.005)       # triangular(lo, hi, mode 1)
        assert float(f.var()) == pytest.approx((lo * lo + hi * hi + 1 - lo * hi - lo - hi) / 18.0, rel=0.03)
    caps = flat.noise_hvac["noise_parameters"][flat.noise_hvac["noise_mode"]]["cooling_capacity_list"][vd["cooling_capacity"]]
    cap = env.raw["cap"]
    counts = [float((cap == c).double().mean()) for c in caps]
    assert sum(counts) == pytest.approx(1.0)
    assert all(c == pytest.approx(1.0 / len(caps), rel=0.05) for c in counts)
    dur = env.lockout_dur
    assert int(dur.min()) == vd["lockout_duration"] - 8 and int(dur.max()) == vd["lockout_duration"] + 8
    assert torch.equal(env.hvac, dur << 2)  # off, no lockout, seconds_since_off = lockout duration
    from mdr_b200.config_flatten import epoch_seconds
    t0 = epoch_seconds(flat.start_datetime)
    off = env.t_epoch - t0
    assert int(off.min()) >= 0 and int(off.max()) < 364 * 86400
    assert float(off.double().mean()) == pytest.approx(364 * 86400 / 2, rel=0.05)
    # exact derived quantities
    torch.testing.assert_close(env.env["max_power"], (cap / flat.hvac_cop).sum(1), rtol=1e-13, atol=0)
    assert (env.env["perlin_seed"] > 0).all() and (env.env["perlin_seed"] < 1).all()
    assert (env.time_since_interp == flat.interp_update_period + 1).all()
    od = env.env["od_temp"]
    assert float(od.min()) > flat.night_temp - 5 * max(flat.temp_std, 0.1) and float(od.max()) < flat.day_temp + 5 * max(flat.temp_std, 0.1)


def test_device_population_is_deterministic_and_indexed():
    import torch
    cfg = _cfg(20)
    _, a = _env(cfg, 64, seed=3)
    _, b = _env(cfg, 64, seed=3)
    a.reset_envs(draw_index=5)
    b.reset_envs(draw_index=5)
    assert torch.equal(a.temps, b.temps) and torch.equal(a.raw["ua"], b.raw["ua"]) and torch.equal(a.t_epoch, b.t_epoch)
    assert torch.equal(a.env["signal"], b.env["signal"])
    b.reset_envs(draw_index=6)
    assert not torch.equal(a.raw["ua"], b.raw["ua"]) and not torch.equal(a.t_epoch, b.t_epoch)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_masked_reset_leaves_other_envs_untouched(precision):
    import torch
    import mdr_b200
    n_envs, n = 37, 30
    cfg = _cfg(n)
    flat, env = _env(cfg, n_envs, precision=precision)
    env.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(25):
        env.step_tensor((torch.rand(n_envs, n, device="cuda", generator=g) < 0.5).to(torch.uint8))
    before = {k: v.clone() for k, v in env.state_dict().items() if isinstance(v, torch.Tensor)}
    mask = torch.zeros(n_envs, dtype=torch.bool, device="cuda")
    mask[::3] = True
    obs = env.reset_envs(mask)
    torch.cuda.synchronize()
    after = {k: v for k, v in env.state_dict().items() if isinstance(v, torch.Tensor)}
    keep = ~mask
    for k in ("temps", "hvac", "coef_a", "coef_b", "coef_c", "lockout_dur", "t_epoch", "raw.ua", "raw.cap", "raw.target",
              "env.od_temp", "env.signal", "env.base_power", "env.cluster_power", "env.phase", "env.perlin_seed"):
        assert torch.equal(after[k][keep], before[k][keep]), k
    assert not torch.equal(after["raw.ua"][mask], before["raw.ua"][mask])
    assert torch.equal(env.hvac[mask], env.lockout_dur[mask] << 2)
    # the re-drawn envs start exactly like a fresh environment built from the same population
    pop = {k: env.raw[k].cpu().numpy() for k in env.raw}
    pop.update(t_air=env.t_air.double().cpu().numpy(), t_mass=env.t_mass.double().cpu().numpy(),
               lockout_dur=env.lockout_dur.cpu().numpy(), sso=env.seconds_since_off.cpu().numpy(),
               on=env.hvac_on.cpu().numpy(), lockout=env.hvac_lockout.cpu().numpy(), t_epoch=env.t_epoch.cpu().numpy(),
               time_since_interp=env.time_since_interp.cpu().numpy())
    pop.update({k: v.cpu().numpy() for k, v in env.env.items()})
    fresh = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=precision, seed=7)
    fobs = fresh.observe_tensor()  # same state, same (already initialised) signal -> same observation everywhere
    torch.testing.assert_close(obs, fobs, rtol=0, atol=0)
    sel = {k: v[mask.cpu().numpy()] if np.ndim(v) else v for k, v in pop.items()}
    sel["signal"] = np.zeros_like(sel["signal"])
    fresh2 = mdr_b200.VecDemandResponseEnv(cfg, sel, precision=precision, seed=7)
    fresh2.reset_tensor()
    # initial signal of the selected envs == what a full reset of exactly those envs computes (device perlin is keyed
    # by the env's own perlin_seed, not by its index)
    torch.testing.assert_close(env.env["signal"][mask], fresh2.env["signal"], rtol=1e-12, atol=0)
    # and the rollout simply continues
    for _ in range(5):
        o, r, p, s = env.step_tensor((torch.rand(n_envs, n, device="cuda", generator=g) < 0.5).to(torch.uint8))
    assert torch.isfinite(o).all() and torch.isfinite(r).all() and torch.isfinite(s).all()


def test_masked_reset_with_observation_pointer_is_refused():
    import ctypes as C
    import torch
    import mdr_b200
    from mdr_b200 import _lib
    cfg = _cfg(10)
    _, env = _env(cfg, 4)
    env.reset_tensor()
    m = torch.ones(4, dtype=torch.uint8, device="cuda")
    env._set_inputs(None, None, None, None, None, None)
    env.in_s.env_mask = C.c_void_p(m.data_ptr())
    st = env.lib.mdr_reset(*env._refs, env._stream())
    env.in_s.env_mask = None
    assert st == -6  # MDR_ERR_UNSUPPORTED: masked reset writes no observation (mdr_observe does)
