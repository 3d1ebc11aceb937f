"""Multi-GPU use of the step path: env-axis sharding and end-of-rollout metric reduction.

Clusters never interact (neighbour lists, power sums, the signal and the reward are all
intra-cluster: env/MA_DemandResponse.py:816-828, 1042-1050, 234-251), so the env axis is cut into
contiguous shards, one per rank / GPU, and **no collective sits on the step path**.  Only the
rollout metrics (the quantities main-deploy.py:124-209 and metrics.py:22-47 accumulate) are
reduced, once, with a single small all-reduce -- NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

SUM_KEYS = ("steps", "house_steps", "sum_reward", "sum_abs_temp_error", "sum_sq_signal_error", "sum_abs_signal_error")
MAX_KEYS = ("max_abs_temp_error",)


def shard_range(n_envs: int, rank: int, world: int):
    """[lo, hi) of the envs owned by `rank`; contiguous, sizes differ by at most one."""
    return (n_envs * rank) // world, (n_envs * (rank + 1)) // world


class RolloutMetrics:
    """Per-shard accumulators, updated from the tensors a step returns (any device)."""

    def __init__(self, device="cpu"):
        self.sums = torch.zeros(len(SUM_KEYS), dtype=torch.float64, device=device)
        self.maxs = torch.zeros(len(MAX_KEYS), dtype=torch.float64, device=device)

    def update(self, reward, t_air, target, power, signal):
        """reward/t_air/target: [E, N]; power/signal: [E] (the step's cluster power and NEW signal)."""
        e, n = reward.shape
        err = (t_air.double() - target.double()).abs()
        ds = signal.double() - power.double()
        upd = torch.stack([
            torch.tensor(1.0, dtype=torch.float64, device=self.sums.device),
            torch.tensor(float(e * n), dtype=torch.float64, device=self.sums.device),
            reward.double().sum(), err.sum(), (ds * ds).sum() * n, ds.abs().sum() * n,
        ])
        self.sums += upd.to(self.sums.device)
        self.maxs = torch.maximum(self.maxs, err.max().reshape(1).to(self.maxs.device))

    def local(self):
        out = {k: float(v) for k, v in zip(SUM_KEYS, self.sums.tolist())}
        out.update({k: float(v) for k, v in zip(MAX_KEYS, self.maxs.tolist())})
        return out


def reduce_metrics(metrics: RolloutMetrics, group=None):
    """All-reduce (sum / max) over the ranks of `group`; returns a dict of python floats with the
    derived means the reference prints (mean reward, mean |T - target|, RMSE of the signal per agent)."""
    sums, maxs = metrics.sums.clone(), metrics.maxs.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
        world = dist.get_world_size(group)
        sums[0] /= world  # every rank counted the same steps
    out = {k: float(v) for k, v in zip(SUM_KEYS, sums.tolist())}
    out.update({k: float(v) for k, v in zip(MAX_KEYS, maxs.tolist())})
    hs = max(out["house_steps"], 1.0)
    out["mean_reward"] = out["sum_reward"] / hs
    out["mean_abs_temp_error"] = out["sum_abs_temp_error"] / hs
    out["rmse_signal"] = (out["sum_sq_signal_error"] / hs) ** 0.5
    return out


def reduce_device_metrics(metrics, group=None):
    """End-of-rollout reduction of the ON-DEVICE accumulators (VecDemandResponseEnv.metrics, [E_shard, N_METRICS],
    see MDR_M_* in include/mdr_b200.h) over the env axis and over the ranks: one all-reduce(sum) + one
    all-reduce(max) of N_METRICS + 1 doubles.  Returns the totals as a dict (sums over all envs of all shards;
    `max_temp_error` is a max; `envs` counts the envs), the only collective of a sharded rollout."""
    from . import _lib
    imax = _lib.METRIC_NAMES.index("max_temp_error")
    sums = torch.cat([metrics.sum(dim=0), torch.tensor([float(metrics.shape[0])], dtype=metrics.dtype, device=metrics.device)])
    mx = metrics[:, imax].max().reshape(1).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = {k: float(v) for k, v in zip(_lib.METRIC_NAMES, sums[:-1].tolist())}
    out["max_temp_error"] = float(mx.item())
    out["envs"] = float(sums[-1].item())
    return out
