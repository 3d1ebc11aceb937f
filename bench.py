#!/usr/bin/env python
"""Benchmark of the MADemandResponseEnv step path (BASELINE.json metric: house-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3] [--impl reference]

One "step" = one pass of the fused step kernel over one batch of synthetic clusters.  Default
workload (`config.workload`) is BASELINE config 4's per-GPU shard: 16,384 envs x 100 houses, fp32,
F = 51 observation, base power interpolated on device from a synthetic table (refresh every 75
steps), weak scaling (every rank owns its own 16,384 clusters; no collective on the step path).

Printed JSON (rank 0): `value` = whole-job house-steps/s with inputs resident in HBM; `e2e` = the
same metric through the host-buffer C-ABI call (pinned host actions in, observation/reward out);
`roofline` = algorithmic bytes per launch / mean launch duration against the measured HBM peak;
`cpu_baseline` = the per-object python port of the reference (oracle/mdr_oracle_scalar.py) on the
host cores.  `--impl reference` times that CPU port alone, on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "house_steps_per_sec"
UNIT = "house-steps/s"
ALGO_BYTES = {("fp32", True): 271, ("fp32", False): 67, ("fp64", True): 527, ("fp64", False): 119}

WORKLOADS = {
    # name: (envs per GPU, houses per env, base power mode, action source, observation written)
    "c4": dict(envs=16384, houses=100, interp=True, action_source="array", obs=True,
               desc="BASELINE config 4 per-GPU shard: 16384 envs x 100 houses, on-device interpolation"),
    "c1": dict(envs=1, houses=1000, interp=True, action_source="array", obs=True, precision="fp64",
               desc="BASELINE config 1: single env, 1,000 houses, regulation signal, nb_agents_comm=10, fp64 equivalence "
                    "run (ONE CTA: launch-latency-bound, a parity configuration rather than a throughput one)"),
    "c2": dict(envs=4096, houses=50, interp=False, action_source="array", obs=True,
               desc="BASELINE config 2: 4096 envs x 50 houses (PPO/MAPPO rollout shape)"),
    "c3": dict(envs=10000, houses=100, interp=False, action_source="bangbang", obs=False,
               desc="BASELINE config 3: 1M houses (10000 clusters x 100), heterogeneous parameters + lockout, on-device "
                    "bang-bang, no per-step observation"),
    "c3fused": dict(envs=10000, houses=100, interp=False, action_source="bangbang", obs=False, fused=75,
                    desc="BASELINE config 3 through the fused multi-step kernel: 1M houses (10000 clusters x 100), 75 env "
                         "steps (5 simulated minutes) per launch with the house state in registers, on-device bang-bang + "
                         "deploy metrics; scored with the per-step 67 B/house-step, flagged fused-K"),
    "c3big": dict(envs=1000, houses=1000, interp=False, action_source="bangbang", obs=False,
                  desc="BASELINE config 3 as 1000 clusters x 1000 houses (one cluster per CTA, generic kernel)"),
}


def workload_config(w):
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = w["houses"]
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if w["interp"] else "constant"
    cfg["default_house_prop"]["solar_gain_bool"] = False  # both reference CLIs force solar gain off
    return cfg


# ------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process = one 100-house cluster stepped by the per-object python port."""
    houses, steps, warmup, seed = args
    import numpy as np
    import mdr_b200
    from oracle import mdr_oracle_scalar as sc
    w = dict(houses=houses, interp=False)
    cfg = workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, 1, seed=seed)
    snap = {k: np.asarray(v)[0] for k, v in pop.items()}
    snap["signal"] = 4200.0 * houses
    if warmup:
        sc.timed_rollout(cfg, snap, warmup, seed)
    n, dt = sc.timed_rollout(cfg, snap, steps, seed)
    return n, dt


def run_cpu_port(houses, steps, warmup, procs):
    """Aggregate house-steps/s of `procs` independent port processes (the reference is single-threaded
    python, so independent env processes are the only way it can use more than one core)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(houses, steps, warmup, 1000 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    rate = sum(n / dt for n, dt in res)
    one = max(n / dt for n, dt in res)
    return rate, one, wall, max(dt for _, dt in res)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    houses = min(w["houses"], 100)
    rate, one, wall, worst = run_cpu_port(houses, args.steps, args.warmup, procs)
    sample = ("%d processes x 1 cluster x %d houses x %d steps of oracle/mdr_oracle_scalar.py (per-object python "
              "port of env.step + normStateDict; constant base power: the interpolation refresh is omitted, which "
              "favours the CPU arm); best single core %.3g house-steps/s" % (procs, houses, args.steps, one))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * worst / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "cpu_sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-i", str(gpu_id), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for t, l in self.lines if t0 <= t <= t1] or [l for _, l in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mdr_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        # NCCL announces its version on STDOUT when the first communicator is created; the contract is ONE JSON line
        # there, so stdout is pointed at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    w = WORKLOADS[args.workload]
    E = args.envs or w["envs"]
    N = w["houses"]
    cfg = workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, E, seed=1234 + rank)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=args.precision, device=dev, seed=1234 + rank,
                                        interp_table=table, action_source=w["action_source"], with_obs=w["obs"])
    env.reset_tensor()
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    ring = [(torch.rand(E, N, device=dev, generator=gen) < 0.5).to(torch.uint8) for _ in range(8)]
    use_array = w["action_source"] == "array"

    fused = int(w.get("fused", 1))  # env steps per launch (fused multi-step kernel) -- a bench "step" stays ONE env step
    if fused > 1:
        env.enable_metrics()
        args.steps = max(fused, (args.steps // fused) * fused)
        args.warmup = max(fused, (args.warmup // fused) * fused)

    def one_step(i):
        if fused > 1:
            if i % fused == 0:
                env.run(fused)
        else:
            env.step_tensor(ring[i & 7] if use_array else None)

    for i in range(args.warmup):
        one_step(i)
    torch.cuda.synchronize(dev)

    # ---- device-timed region: K launches between two events on the launch stream
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = local
    sampler = ClockSampler(uuid) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t_host0 = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        one_step(i)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    t_host1 = time.perf_counter()
    if distributed:
        dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_host0, t_host1) if sampler else None
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * E * N * args.steps / (max_ms * 1e-3)

    # ---- end-to-end: host-buffer C-ABI call, pinned H2D of the actions + D2H of obs/reward/power/signal
    k_e2e = max(3, min(args.steps, args.e2e_steps))
    host_actions = [r.cpu().numpy() for r in ring[:4]]
    for i in range(2):
        env.step_host(host_actions[i & 3])
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(k_e2e):
        env.step_host(host_actions[i & 3])
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E * N * k_e2e / float(t.item())
    rb = 4 if args.precision == "fp32" else 8
    h2d = E * N
    d2h = (E * N * env.n_features * rb if w["obs"] else 0) + E * N * rb + 16 * E

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, peak_kind = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_kind = 6650.0, "fallback (B200_PROFILING.md)"
        algo = ALGO_BYTES[(args.precision, w["obs"])]
        launch_s = elapsed_ms * 1e-3 / args.steps
        achieved = algo * E * N / launch_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath)).get("%s_%s" % (args.workload, args.precision))
        geom = env.launch_geometry()
        if fused > 1:
            geom["kernel"] = "mdr::run_fused_kernel (%d env steps per launch, state in registers)" % fused
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            procs = os.cpu_count() or 1
            rate, one, wall, _ = run_cpu_port(min(N, 100), args.cpu_steps, 20, procs)
            cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                   "sample": "%d processes x 1 cluster x %d houses x %d steps of oracle/mdr_oracle_scalar.py "
                             "(per-object python port of env.step + normStateDict, constant base power), %.1f s wall; "
                             "best single core %.3g house-steps/s" % (procs, min(N, 100), args.cpu_steps, wall, one)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "envs_per_gpu": E, "houses_per_env": N,
                       "obs_features": env.n_features if w["obs"] else 0, "precision": args.precision,
                       "actions": "ring of 8 pre-generated uint8 [E,N] tensors in HBM" if use_array else w["action_source"],
                       "l2": "per-step working set %.0f MB > 126 MB L2 (no flush needed)" % (algo * E * N / 1e6)
                             if algo * E * N > 126e6 else "working set fits L2: state/params are re-read from L2 every step, "
                             "as in a real rollout; obs writes stream to HBM",
                       "launch": geom, "env_steps_per_launch": fused, "parallelism": "env-sharded x%d, no collective on the step path" % world},
            "clocks": clocks, "gpu_launches": args.steps // fused,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": k_e2e, "api": "VecDemandResponseEnv.step_host -> mdr_step_host (pinned host buffers)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_kind, "algorithmic_bytes_per_house_step": algo,
                         "kernel": geom["kernel"], "launch_us": launch_s * 1e6 * fused, "fused_k": fused},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "fp64"], help="default: the workload's (fp32; c1: fp64)")
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=4000,
                    help="env steps per CPU-baseline process (default: ~3 s per core, ~45 core-seconds on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.precision is None:
        args.precision = WORKLOADS[args.workload].get("precision", "fp32")
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
