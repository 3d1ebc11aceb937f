"""CPU tests of the host-side logic: config flattening, neighbour tables, the reference-order
population builder (vs golden snapshots recorded from the reference), the host perlin, and that
the C-ABI library loads and exports every symbol of include/mdr_b200.h (no compute calls)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

import golden_util as gu
import mdr_b200
from mdr_b200 import _lib, config_flatten, population
from mdr_b200.perlin import Perlin
from oracle import mdr_oracle as orc
from oracle import ref_stubs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mdr_b200.h")).read()
    declared = set(re.findall(r"\b(mdr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = mdr_b200.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mdr_version() == _lib.MDR_ABI_VERSION
    assert lib.mdr_strerror(0) == b"ok"
    assert b"NULL" in lib.mdr_strerror(-1)


def test_struct_layout_matches_header():
    """sizeof(MdrConfig) as laid out by ctypes must equal what nvcc compiled (checked through validate)."""
    lib = mdr_b200.load_library()
    flat = mdr_b200.FlatConfig(mdr_b200.make_default_config())
    cfg = flat.to_struct(4, _lib.F32, 0, seed=123)
    assert lib.mdr_validate(ctypes.byref(cfg)) == 0
    assert lib.mdr_obs_width(ctypes.byref(cfg)) == flat.obs_width() == 11  # nb_agents = 1 -> no neighbour
    cfg.seed = 2**64 - 1  # last field: a layout mismatch would shift it
    assert lib.mdr_validate(ctypes.byref(cfg)) == 0
    cfg.abi_version = 99
    assert lib.mdr_validate(ctypes.byref(cfg)) == -7
    cfg.abi_version = _lib.MDR_ABI_VERSION
    cfg.signal_mode = 17
    assert lib.mdr_validate(ctypes.byref(cfg)) == -3
    with pytest.raises(ValueError):
        _lib.check(-3, "x")
    with pytest.raises(_lib.MdrError):
        _lib.check(-2, "x")


@pytest.mark.parametrize("n,expect_f", [(50, 51), (1000, 51), (5, 27), (1, 11)])
def test_obs_width(n, expect_f):
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = n
    flat = mdr_b200.FlatConfig(cfg)
    assert flat.obs_width() == expect_f
    for k in cfg["default_env_prop"]["state_properties"]:
        cfg["default_env_prop"]["state_properties"][k] = True
    for k in cfg["default_env_prop"]["message_properties"]:
        cfg["default_env_prop"]["message_properties"][k] = True
    flat = mdr_b200.FlatConfig(cfg)
    c = min(10, n - 1)
    assert flat.obs_width() == 23 + 11 * c
    lib = mdr_b200.load_library()
    assert lib.mdr_obs_width(ctypes.byref(flat.to_struct(1, _lib.F64, 0))) == 23 + 11 * c


def test_launch_geometry():
    lib = mdr_b200.load_library()
    cfg = mdr_b200.make_default_config()
    out = {}
    for n, e in ((50, 4096), (100, 16384), (1000, 1), (37, 3), (1024, 2)):
        cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = n
        cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "constant"
        flat = mdr_b200.FlatConfig(cfg)
        for prec in (_lib.F32, _lib.F64):
            s = flat.to_struct(e, prec, 0)
            g, t, c, sm = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_size_t()
            pl, cl = ctypes.c_int32(), ctypes.c_int32()
            assert lib.mdr_launch_geometry(ctypes.byref(s), 1, ctypes.byref(g), ctypes.byref(t), ctypes.byref(c),
                                           ctypes.byref(sm), ctypes.byref(pl), ctypes.byref(cl)) == 0
            # persistent pipelined kernel: fp32, whole envs of <= 224 houses per tile -- with solar gain on (the shipped
            # default) or off, it is a run-time variant of the same kernel
            assert pl.value == (1 if prec == _lib.F32 else 0)   # N = 1000 / 1024: the split (cluster) instantiation
            cfg2 = __import__('copy').deepcopy(cfg)
            cfg2['default_house_prop']['solar_gain_bool'] = False
            s2 = mdr_b200.FlatConfig(cfg2).to_struct(e, prec, 0)
            assert lib.mdr_launch_geometry(ctypes.byref(s2), 1, None, None, None, None, ctypes.byref(pl), None) == 0
            assert pl.value == (1 if prec == _lib.F32 else 0)
            assert t.value <= 1024 and t.value % 32 == 0
            if n <= 224:
                assert cl.value == 1 and g.value * n <= t.value and c.value == -(-e // g.value)
            else:  # one env split over a thread-block cluster (e.g. 1000 houses -> 8 CTAs x 128 or 5 x 200)
                assert g.value == 1 and 2 <= cl.value <= 8 and c.value == e * cl.value
                assert n <= cl.value * (t.value - 32)
            assert sm.value <= 227 * 1024
            out[(n, prec)] = (g.value, t.value, c.value, sm.value)
    assert out[(50, 4)][0] == 4 and out[(100, 4)][0] == 2  # CTA start rows 16-byte aligned for the bulk store
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = 2000
    s = mdr_b200.FlatConfig(cfg).to_struct(1, _lib.F32, 0)
    assert lib.mdr_validate(ctypes.byref(s)) == 0   # N > 1024: one env over a thread-block cluster
    s = mdr_b200.FlatConfig(cfg).to_struct(1, _lib.F32, 0, action_source="greedy")
    assert lib.mdr_validate(ctypes.byref(s)) == -6  # the on-device greedy controller sorts inside one CTA


def test_workspace_size_and_wide_kernel_geometry():
    """Host-only ABI answers: the scratch of the pipelined kernel (two launches in flight: due-tile queue, claim header,
    a due word + a due-list entry per tile, a 64-byte record per env) and the kernel a step without observation takes
    on clusters beyond one tile (one CTA walks a whole env)."""
    lib = mdr_b200.load_library()
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["power_grid_prop"]["base_power_mode"] = "constant"
    for n, e, wide in ((100, 16384, False), (1000, 1000, True), (1000, 8, False), (9000, 100, False)):
        cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = n
        s = mdr_b200.FlatConfig(cfg).to_struct(e, _lib.F32, 0)
        need = ctypes.c_size_t()
        assert lib.mdr_workspace_bytes(ctypes.byref(s), ctypes.byref(need)) == 0
        r64 = lambda x: (x + 63) // 64 * 64
        assert need.value >= 2 * (r64(64 + 4 * e) + 64 + 2 * r64(4 * e) + 64 * e), (n, e, need.value)
        pl, cl, c, t = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        assert lib.mdr_launch_geometry(ctypes.byref(s), 0, None, ctypes.byref(t), ctypes.byref(c), None, ctypes.byref(pl),
                                       ctypes.byref(cl)) == 0
        assert (pl.value == 2) == wide, (n, e, pl.value)
        if wide:
            assert c.value == e and t.value == 256 and cl.value == 1
        s.flags = _lib.FLAG_NO_PIPELINE   # tests pin the generic kernel this way
        assert lib.mdr_launch_geometry(ctypes.byref(s), 0, None, None, None, None, ctypes.byref(pl), None) == 0
        assert pl.value == 0


def test_bad_modes_raise_like_the_reference():
    for path, val in ((("cluster_prop", "agents_comm_mode"), "bogus"),
                      (("reward_prop", "temp_penalty_mode"), "bogus"),
                      (("reward_prop", "sig_penalty_mode"), "bogus"),
                      (("power_grid_prop", "base_power_mode"), "bogus")):
        cfg = mdr_b200.make_default_config()
        cfg["default_env_prop"][path[0]][path[1]] = val
        with pytest.raises(ValueError):
            mdr_b200.FlatConfig(cfg)
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["power_grid_prop"]["signal_mode"] = "bogus"
    with pytest.raises((ValueError, KeyError)):
        mdr_b200.FlatConfig(cfg)
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"].update(nb_agents=24, agents_comm_mode="neighbours_2D")
    with pytest.raises(ValueError):
        mdr_b200.FlatConfig(cfg)


@pytest.mark.parametrize("mode,n,c", [("neighbours", 12, 10), ("neighbours", 12, 3), ("neighbours", 5, 10),
                                      ("neighbours", 37, 10), ("closed_groups", 12, 3), ("closed_groups", 10, 3),
                                      ("neighbours_2D", 25, 10), ("no_message", 7, 10), ("neighbours", 1000, 10)])
def test_comm_table_matches_oracle(mode, n, c):
    a = config_flatten.comm_table(mode, n, c)
    b = orc.comm_links(mode, n, c)
    assert a.dtype == np.int32 and np.array_equal(a, b)


@pytest.mark.parametrize("name", gu.names())
def test_population_builder_consumes_rng_like_the_reference(name):
    """random.seed(s); Env(config); env.reset()  ==  two reference-order builds."""
    g = gu.Golden(name)
    flat = mdr_b200.FlatConfig(g.config)
    random.seed(g.seed)
    for _ in range(2):
        pop, table = population.reference_order_population(flat, random)
        if g.uses_interp and flat.n_houses > flat.interp_nb_agents:
            ids = random.choices(list(range(flat.n_houses)), k=flat.interp_nb_agents)  # PowerGrid.step at :133
    for k in ("ua", "cm", "ca", "hm", "cap", "target", "deadband", "t_air", "t_mass", "lockout_dur", "sso", "on",
              "lockout"):
        assert np.array_equal(np.asarray(pop[k][0], dtype=np.float64), np.asarray(g.snap[k], dtype=np.float64)), k
    for k in ("t_epoch", "phase", "od_temp", "artificial_ratio", "max_power"):
        assert float(pop[k][0]) == float(g.snap[k]), k
    if g.comm is not None and table is not None:
        assert np.array_equal(table, g.comm)
    if g.uses_interp and flat.n_houses > flat.interp_nb_agents and "init_interp_ids" in g.z:
        assert list(g.z["init_interp_ids"]) == ids


def test_host_perlin_matches_stub_restatement():
    """Two independent restatements of the published perlin-noise algorithm agree (parity of the
    third-party package itself is unpinned, see oracle/ref_stubs.py)."""
    seed = 0.3791
    mine = Perlin(1, 5, 5, 400, seed)
    octs = [ref_stubs.PerlinNoise(octaves=2 ** i * 5, seed=seed) for i in range(5)]
    for x in (0.0, 3.7, 1234.0, 40000.0, 86396.0):
        exp = orc.perlin_mix([o.noise(x / 400) for o in octs], 5)
        assert mine.calculate_noise(x) == pytest.approx(exp, abs=1e-15)


def test_synthetic_population_shapes_and_shards():
    cfg = mdr_b200.make_default_config()
    cfg["default_env_prop"]["cluster_prop"]["nb_agents"] = 20
    flat = mdr_b200.FlatConfig(cfg)
    pop = population.synthetic_population(flat, 8, seed=5)
    assert pop["t_air"].shape == (8, 20) and pop["t_epoch"].shape == (8,)
    assert (pop["sso"] == pop["lockout_dur"]).all() and (pop["on"] == 0).all()
    assert np.allclose(pop["max_power"], (pop["cap"] / 2.5).sum(1))
    shards = [population.shard_population(pop, r, 4) for r in range(4)]
    assert sum(len(s["t_epoch"]) for s in shards) == 8
    assert np.array_equal(np.concatenate([s["t_air"] for s in shards]), pop["t_air"])


@pytest.mark.skipif(not ref_stubs.reference_available(), reason="reference tree only exists in the build container")
def test_default_config_equals_reference_config():
    import json
    _, _, cfg, _ = ref_stubs.import_reference()
    mine = mdr_b200.make_default_config()
    for k in mine:
        assert mine[k] == cfg[k], k
    ref_grid = json.load(open(os.path.join(ref_stubs.REFERENCE_ROOT, "monteCarlo", "interp_parameters_dict.json")))
    assert ref_grid == mdr_b200.default_config.INTERP_GRID == gu.INTERP_GRID


def test_closed_groups_overrun_fails_like_the_reference():
    """env/MA_DemandResponse.py:834 admits a group that runs one id past the last house when base + nb_comm == nb_agents;
    the reference then raises KeyError while building the first observation.  Here construction fails instead of a kernel
    silently gathering a foreign message."""
    import pytest
    from mdr_b200.config_flatten import comm_table
    with pytest.raises(KeyError):
        comm_table("closed_groups", 9, 4)     # groups of 5: base 5 + 4 == 9
    t = comm_table("closed_groups", 10, 4)    # fits exactly
    assert t.shape == (10, 4) and int(t.max()) == 9
    t = comm_table("closed_groups", 12, 3)
    assert list(t[11]) == [8, 9, 10]


def test_bench_reference_arm_prints_the_contract_line():
    """`python bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the
    contract's keys, no GPU needed."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "30", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "house_steps_per_sec" and line["unit"] == "house-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 30
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_population_spec_mirrors_the_config_entries():
    """MdrPopulationSpec (device-side population draw) carries exactly the entries utils.applyPropertyNoise reads."""
    import pytest
    cfg = mdr_b200.make_default_config()
    cfg["noise_house_prop"]["noise_mode"] = "big_noise"
    cfg["noise_hvac_prop"]["noise_mode"] = "big_noise"
    cfg["default_hvac_prop"]["lockout_noise"] = 8
    flat = mdr_b200.FlatConfig(cfg)
    sp = mdr_b200.population_spec(flat)
    nh = cfg["noise_house_prop"]["noise_parameters"]["big_noise"]
    hd, vd = cfg["default_house_prop"], cfg["default_hvac_prop"]
    assert (sp.std_start_temp, sp.std_target_temp) == (nh["std_start_temp"], nh["std_target_temp"])
    assert (sp.factor_thermo_low, sp.factor_thermo_high) == (nh["factor_thermo_low"], nh["factor_thermo_high"])
    assert (sp.ua, sp.cm, sp.ca, sp.hm) == (hd["Ua"], hd["Cm"], hd["Ca"], hd["Hm"])
    caps = cfg["noise_hvac_prop"]["noise_parameters"]["big_noise"]["cooling_capacity_list"][vd["cooling_capacity"]]
    assert sp.n_cap == len(caps) and [sp.cap_list[i] for i in range(sp.n_cap)] == [float(c) for c in caps]
    assert (sp.lockout_duration, sp.lockout_noise) == (vd["lockout_duration"], 8)
    assert sp.random_start == 1 and sp.interp_update_period == flat.interp_update_period
    from mdr_b200.config_flatten import epoch_seconds
    assert sp.start_epoch == epoch_seconds(flat.start_datetime)
    cfg["default_hvac_prop"]["lockout_noise"] = vd["lockout_duration"] + 1   # HVAC.__init__ :438-461
    with pytest.raises(ValueError):
        mdr_b200.population_spec(mdr_b200.FlatConfig(cfg))
