"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous env shards stepped
independently (here by the numpy oracle standing in for the per-GPU kernel) and one end-of-rollout
metric reduction must reproduce the single-process result exactly -- there is no collective on
the step path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mdr_b200
from mdr_b200 import sharding
from oracle import mdr_oracle as orc

E, N, STEPS = 6, 12, 15


def _case():
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = N
    ep["power_grid_prop"]["base_power_mode"] = "constant"
    ep["power_grid_prop"]["signal_mode"] = "sinusoidals"
    cfg["default_house_prop"]["solar_gain_bool"] = False
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, E, seed=3)
    rng = np.random.default_rng(4)
    actions = rng.integers(0, 2, (STEPS, E, N)).astype(np.uint8)
    noise = rng.normal(0, 0.5, (STEPS, E))
    return cfg, pop, actions, noise


def _rollout(cfg, pop, actions, noise):
    snap = {k: v for k, v in pop.items() if k != "perlin_seed"}
    env = orc.OracleEnv(cfg, snap)
    for e in range(env.E):
        env.grid_step(e, orc.to_datetime(env.s["t_epoch"][e]))
    m = sharding.RolloutMetrics()
    for t in range(actions.shape[0]):
        _, rew, p, s = env.step(actions[t], noise[t])
        m.update(torch.as_tensor(rew), torch.as_tensor(env.s["t_air"]), torch.as_tensor(env.s["target"]),
                 torch.as_tensor(p), torch.as_tensor(s))
    return m


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, pop, actions, noise = _case()
    lo, hi = sharding.shard_range(E, rank, world)
    shard = mdr_b200.shard_population(pop, rank, world)
    assert len(shard["t_epoch"]) == hi - lo
    m = _rollout(cfg, shard, actions[:, lo:hi], noise[:, lo:hi])
    red = sharding.reduce_metrics(m)
    if rank == 0:
        out.update(red)
    dist.destroy_process_group()


def test_shard_ranges_partition_the_env_axis():
    for n, w in ((16384, 8), (10, 3), (7, 8), (1, 1)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_rollout_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        sharded = dict(out)
    cfg, pop, actions, noise = _case()
    single = sharding.reduce_metrics(_rollout(cfg, pop, actions, noise))
    assert sharded["steps"] == single["steps"] == STEPS
    assert sharded["house_steps"] == single["house_steps"] == E * N * STEPS
    for k in ("sum_reward", "sum_abs_temp_error", "sum_sq_signal_error", "mean_reward", "rmse_signal"):
        assert sharded[k] == pytest.approx(single[k], rel=1e-12), k
    assert sharded["max_abs_temp_error"] == single["max_abs_temp_error"]


def _device_metrics_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _fake_device_metrics()
    lo, hi = sharding.shard_range(full.shape[0], rank, world)
    red = sharding.reduce_device_metrics(full[lo:hi])
    if rank == 0:
        out.update(red)
    dist.destroy_process_group()


def _fake_device_metrics():
    """Stand-in for VecDemandResponseEnv.metrics ([E, N_METRICS] fp64, layout MDR_M_* of include/mdr_b200.h)."""
    from mdr_b200 import _lib
    g = torch.Generator().manual_seed(11)
    m = torch.rand(9, _lib.N_METRICS, dtype=torch.float64, generator=g)
    m[:, 0] = 25.0
    return m


def test_device_metric_layout_matches_header():
    from mdr_b200 import _lib
    import re
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mdr_b200.h")).read()
    names = re.findall(r"MDR_M_([A-Z_]+) = (\d+)", header)
    assert [n.lower() for n, _ in names] == list(_lib.METRIC_NAMES)
    assert [int(i) for _, i in names] == list(range(_lib.N_METRICS))
    assert int(re.search(r"MDR_N_METRICS = (\d+)", header).group(1)) == _lib.N_METRICS


def test_two_rank_gloo_device_metric_reduction():
    from mdr_b200 import _lib
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_device_metrics_worker, args=(2, port, out), nprocs=2, join=True)
        sharded = dict(out)
    full = _fake_device_metrics()
    single = sharding.reduce_device_metrics(full)
    assert sharded["envs"] == single["envs"] == 9
    for k in _lib.METRIC_NAMES:
        assert sharded[k] == pytest.approx(single[k], rel=1e-12), k
    assert sharded["max_temp_error"] == float(full[:, _lib.METRIC_NAMES.index("max_temp_error")].max())
