#!/usr/bin/env python
"""Benchmark of the MADemandResponseEnv step path (BASELINE.json metric: house-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c0|c1|c2|c2actor|c3|c3fused|c3big] [--impl reference]

One "step" = one pass of the step kernel over one batch of synthetic clusters.  Default workload (`config.workload`)
is BASELINE config 4's per-GPU shard: 16,384 envs x 100 houses, fp32, F = 51 observation, base power interpolated on
device from a synthetic table (refresh clocks staggered over the 75-step period, so every step refreshes ~1/75 of the
clusters), weak scaling (every rank owns its own 16,384 clusters; no collective on the step path).

Printed JSON (rank 0): `value` = whole-job house-steps/s with inputs resident in HBM; `e2e` = the same metric through
the host-buffer C-ABI call (pinned host actions in, observation/reward out); `roofline` = algorithmic bytes per launch /
mean launch duration against the measured HBM peak; `cpu_baseline` = the per-object python port of the reference
(oracle/mdr_oracle_scalar.py, pinned to reference traces) on the host cores.  `--impl reference` times that CPU port
alone on all host cores, on the same workload shape (a bounded sample of it per step).
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
    "c0": dict(envs=1, houses=50, interp=False, action_source="array", obs=True, precision="fp64", dict_api=True,
               desc="BASELINE config 0: main-deploy.py's loop on the drop-in dict API (MADemandResponseEnv.reset/step, python "
                    "bang-bang controller per house, 50 houses, 4 s steps); host-bound by construction"),
    "c1": dict(envs=1, houses=1000, interp=True, action_source="array", obs=True, precision="fp64", launch_batch=25,
               desc="BASELINE config 1: single env, 1,000 houses, regulation signal, nb_agents_comm=10, fp64 equivalence "
                    "run (one thread-block cluster of 5 CTAs: launch-latency-bound, a parity configuration)"),
    "c2": dict(envs=4096, houses=50, interp=False, action_source="array", obs=True,
               desc="BASELINE config 2: 4096 envs x 50 houses (PPO/MAPPO rollout shape), actions from a ring of tensors"),
    "c2actor": dict(envs=4096, houses=50, interp=False, action_source="array", obs=True, actor=True, rollout=16,
                    desc="BASELINE config 2 literally: 4096 envs x 50 houses, actions from the seeded torch policy "
                         "Actor(51, 2, [100, 100]) evaluated in batch on the device, sampled by mdr_sample_actions, "
                         "transitions collected in PPO layout (CUDA graph of 16-step rollouts)"),
    "c3": dict(envs=10000, houses=100, interp=False, action_source="bangbang", obs=False,
               desc="BASELINE config 3: 1M houses (10000 clusters x 100), heterogeneous parameters + lockout, on-device "
                    "bang-bang, no per-step observation"),
    "c3fused": dict(envs=10000, houses=100, interp=False, action_source="bangbang", obs=False, fused=75,
                    desc="BASELINE config 3 through the fused multi-step kernel: 1M houses (10000 clusters x 100), 75 env "
                         "steps (5 simulated minutes) per launch with the house state in registers, on-device bang-bang + "
                         "deploy metrics; scored with the per-step 67 B/house-step, flagged fused-K (not an HBM figure)"),
    "c3big": dict(envs=1000, houses=1000, interp=False, action_source="bangbang", obs=False, launch_batch=25,
                  desc="BASELINE config 3 as 1000 clusters x 1000 houses (one CTA walks a whole cluster; 25 env steps per C "
                       "call: one python call per step is slower than the kernel)"),
}


def workload_config(w, variant=None):
    """`variant` (bench.py --solar / --comm-defect P): non-default settings that stay on the pipelined kernel."""
    import mdr_b200
    cfg = mdr_b200.make_default_config()
    ep = cfg["default_env_prop"]
    ep["cluster_prop"]["nb_agents"] = w["houses"]
    ep["power_grid_prop"]["base_power_mode"] = "interpolation" if w["interp"] else "constant"
    cfg["default_house_prop"]["solar_gain_bool"] = bool(variant and variant.get("solar"))  # both reference CLIs force it off
    if variant and variant.get("comm_defect"):
        ep["cluster_prop"]["comm_defect_prob"] = float(variant["comm_defect"])  # message drops (:992), Philox on the device
    return cfg


def config_block(args, w):
    """`config` of the JSON line: identical for the GPU arm and the reference arm (same workload, same shape)."""
    obs = w["obs"]
    return {"workload": args.workload + ": " + w["desc"], "envs_per_gpu": args.envs or w["envs"], "houses_per_env": w["houses"],
            "obs_features": 51 if obs else 0, "precision": args.precision,
            "base_power": ("interpolation (synthetic table, refresh every 75 steps, %s clocks)"
                           % ("in-phase" if getattr(args, "in_phase", False) else "staggered")) if w["interp"] else "constant",
            "variant": ", ".join(x for x in ("solar gain on" if getattr(args, "solar", False) else "",
                                             "comm_defect_prob %g" % args.comm_defect if getattr(args, "comm_defect", 0) else "",
                                             "device metrics on" if getattr(args, "metrics", False) else "") if x) or "default",
            "actions": ("seeded Actor(51,2,[100,100]) on device" if w.get("actor") else
                        "python bang-bang per house" if w.get("dict_api") else
                        "uint8 [E,N] tensors" if w["action_source"] == "array" else w["action_source"])}


# ------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process = one cluster stepped by the per-object python port (bang-bang policy + per-agent normalisation)."""
    houses, steps, warmup, seed, interp = args
    import numpy as np
    import mdr_b200
    from oracle import mdr_oracle as orc
    from oracle import mdr_oracle_scalar as sc
    w = dict(houses=houses, interp=interp)
    cfg = workload_config(w)
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, 1, seed=seed)
    snap = {k: np.asarray(v)[0] for k, v in pop.items()}
    snap["signal"] = 4200.0 * houses
    table = None
    if interp:
        from mdr_b200.default_config import INTERP_GRID, INTERP_KEYS
        table = orc.PowerInterp(mdr_b200.synthetic_interp_table(), INTERP_GRID, INTERP_KEYS)
        snap["base_power"] = 4200.0 * houses
        snap["time_since_interp"] = (seed * 37) % 300 // 4 * 4   # staggered refresh clocks, like the GPU arm
    if warmup:
        sc.timed_rollout(cfg, snap, warmup, seed, table)
    n, dt = sc.timed_rollout(cfg, snap, steps, seed, table)
    return n, dt


def run_cpu_port(houses, steps, warmup, procs, interp=False):
    """Aggregate house-steps/s of `procs` independent port processes (the reference is single-threaded
    python, so independent env processes are the only way it can use more than one core)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(houses, steps, warmup, 1000 + i, interp) for i in range(procs)])
    wall = time.perf_counter() - t0
    rate = sum(n / dt for n, dt in res)
    one = max(n / dt for n, dt in res)
    return rate, one, wall, max(dt for _, dt in res)


REF_INNER = 150  # env steps of one cluster per reference-arm "step" (two interpolation refreshes at 75 steps)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    houses = w["houses"]
    single = bool(w.get("dict_api")) or w["envs"] == 1   # the reference is single-threaded: one env = one core
    procs = 1 if single else (os.cpu_count() or 1)
    inner = REF_INNER if houses <= 200 else max(1, REF_INNER * 100 // houses)
    steps, warm = args.steps * inner, min(args.warmup, 3) * 10
    rate, one, wall, worst = run_cpu_port(houses, steps, warm, procs, w["interp"])
    sample = ("%d process%s x 1 cluster x %d houses x %d env steps (= %d bench steps of %d env steps) of "
              "oracle/mdr_oracle_scalar.py: per-object python port of env.step + bang-bang + normStateDict, pinned to "
              "reference traces, %s; %.1f s wall, slowest process %.1f s; best single core %.3g house-steps/s"
              % (procs, "es" if procs > 1 else "", houses, steps, args.steps, inner,
                 "interpolated base power refreshed every 75 steps" if w["interp"] else "constant base power", wall, worst, one))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * worst / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args, w),
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


def pin_to_gpu_numa_node(dev_index):
    """Pins this process to the cores of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (first touch then lands on that node): with 8 ranks the observation copies otherwise all stream into node 0."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        dev = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


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
    numa = pin_to_gpu_numa_node(local) if distributed else None
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
    if w.get("dict_api"):
        return dict_api_arm(args, w, dev)
    E = args.envs or w["envs"]
    N = w["houses"]
    variant = {"solar": args.solar, "comm_defect": args.comm_defect, "metrics": args.metrics}
    cfg = workload_config(w, variant)
    flat = mdr_b200.FlatConfig(cfg)
    pop = mdr_b200.synthetic_population(flat, E, seed=1234 + rank)
    table = mdr_b200.synthetic_interp_table() if w["interp"] else None
    env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision=args.precision, device=dev, seed=1234 + rank,
                                        interp_table=table, action_source=w["action_source"], with_obs=w["obs"])
    env.reset_tensor()
    if w["interp"] and not args.in_phase:
        env.stagger_interp_clock(seed=77 + rank)  # rollouts restart clusters at different times: refreshes are spread out
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    ring = [(torch.rand(E, N, device=dev, generator=gen) < 0.5).to(torch.uint8) for _ in range(8)]
    use_array = w["action_source"] == "array"

    fused = int(w.get("fused", 1))       # env steps per launch (fused multi-step kernel) -- a bench "step" stays ONE env step
    batch = int(w.get("launch_batch", 1))  # env steps issued per C call (tiny workloads: keeps the python call out of the way)
    rollout = int(w.get("rollout", 0))   # c2actor: env steps per collected (graph-replayed) rollout
    chunk = max(fused, batch, rollout, 1)
    if chunk > 1:
        args.steps = max(chunk, (args.steps // chunk) * chunk)
        args.warmup = max(chunk, -(-args.warmup // chunk) * chunk)
    if fused > 1 or args.metrics:
        env.enable_metrics()   # the 13 deploy / training accumulators on the device (main-deploy.py:124-209)
    collector = actor = None
    if rollout:
        torch.manual_seed(1)
        actor = mdr_b200.ActorMLP(env.n_features, 2, [100, 100]).to(dev)   # agents/network.py:14-33, BASELINE config 2
        collector = mdr_b200.DeviceRolloutCollector(env, n_steps=rollout, seed=1234 + rank)

    def one_step(i):
        if rollout:
            if i % rollout == 0:
                collector.collect(actor, reset=(i == 0 and collector.total_steps == 0))
        elif fused > 1:
            if i % fused == 0:
                env.run(fused)
        elif batch > 1:
            if i % batch == 0:
                env.step_tensor(ring[(i // batch) & 7] if use_array else None, n_steps=batch)
        else:
            env.step_tensor(ring[i & 7] if use_array else None)

    # ---- device-timed region: K steps between two events on the launch stream
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = local
    sampler = ClockSampler(uuid) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)  # (nvidia-smi needs a moment to start sampling; the GPU idles meanwhile)
    if distributed:
        dist.barrier()
    # The W warm-up steps run back to back with the timed ones, BEHIND the idle gap of the set-up and of the sampler's
    # start: with the warm-up in front of that gap a 20-step window measured 97.8 us per step against 92.1 us for 600
    # steps (now 92.2).  --prewarm-ms (default 0) issues the same step, untimed, for that long before the warm-up.
    n_pre = 0
    if args.prewarm_ms > 0:
        t_pre = time.perf_counter()
        while (time.perf_counter() - t_pre) * 1e3 < args.prewarm_ms:
            for i in range(chunk * 8):
                one_step(n_pre + i)
            n_pre += chunk * 8
            torch.cuda.synchronize(dev)
    for i in range(n_pre, n_pre + args.warmup):
        one_step(i)
    torch.cuda.synchronize(dev)
    if distributed:
        dist.barrier()
        torch.cuda.synchronize(dev)
    t_host0 = time.perf_counter()
    ev0.record(stream)
    for i in range(n_pre + args.warmup, n_pre + args.warmup + args.steps):
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

    env_share = None
    if rollout:
        # the env's share of the rollout: the same number of plain step launches on the same env
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(8):
            env.step_tensor(ring[i & 7])
        e0.record(stream)
        for i in range(args.steps):
            env.step_tensor(ring[i & 7])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        env_share = e0.elapsed_time(e1) / elapsed_ms

    # ---- end-to-end: host-buffer C-ABI call, pinned H2D of the actions + D2H of obs/reward/power/signal
    k_e2e = max(3, min(args.steps, args.e2e_steps))
    host_actions = [r.cpu().numpy() for r in ring[:4]]
    if not args.serial_e2e:
        ncpu = len(os.sched_getaffinity(0))
        # single rank: every core of the affinity mask but one (auto); several ranks: an equal share of the cores
        env.host_pipeline(True, n_threads=0 if not distributed else max(1, min(24, (os.cpu_count() or ncpu) // world - 1)))
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
    d2h = env.host_transfer_bytes() if hasattr(env, "host_transfer_bytes") else \
        (E * N * env.n_features * rb if w["obs"] else 0) + E * N * rb + 16 * E

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
            rate, one, wall, worst = run_cpu_port(min(N, 1000), args.cpu_steps if N <= 200 else max(50, args.cpu_steps * 100 // N),
                                                  20, procs, w["interp"])
            cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                   "sample": "%d processes x 1 cluster x %d houses x %d env steps of oracle/mdr_oracle_scalar.py (per-object "
                             "python port of env.step + bang-bang + normStateDict, %s), %.1f s wall; best single core %.3g "
                             "house-steps/s" % (procs, min(N, 1000), args.cpu_steps if N <= 200 else max(50, args.cpu_steps * 100 // N),
                                                "interpolated base power refreshed every 75 steps" if w["interp"] else
                                                "constant base power", wall, one)}
        config = config_block(args, w)
        config["l2"] = ("per-step working set %.0f MB > 126 MB L2 (no flush needed)" % (algo * E * N / 1e6)
                        if algo * E * N > 126e6 else "working set fits L2: state/params are re-read from L2 every step, "
                        "as in a real rollout; obs writes stream to HBM")
        config["parallelism"] = "env-sharded x%d, no collective on the step path" % world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "prewarm_ms": args.prewarm_ms,  # untimed issue of the same step in front of the W warm-up steps (see above)
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": config, "launch": geom, "env_steps_per_launch": fused, "numa_node": numa,
            "clocks": clocks, "gpu_launches": args.steps // fused * (2 if rollout else 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": k_e2e, "api": "VecDemandResponseEnv.step_host -> mdr_step_host (pinned host buffers)",
                    "pipeline": env.host_pipeline_info()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_kind, "algorithmic_bytes_per_house_step": algo,
                         "kernel": geom["kernel"], "launch_us": launch_s * 1e6 * fused, "fused_k": fused,
                         "note": ("fused-K: the house state never leaves the SM between the K steps; `achieved` is nominal "
                                  "(SURVEY 8d scoring rule), not HBM utilisation") if fused > 1 else
                                 ("per-step working set %.0f MB fits the 126 MB L2: state and parameters are re-read from L2, so "
                                  "`achieved` (algorithmic bytes / launch time) is nominal, not HBM utilisation"
                                  % (algo * E * N / 1e6)) if algo * E * N <= 126e6 else None},
            "cpu_baseline": cpu,
        }
        if rollout:
            line["rollout"] = {"steps_per_graph": rollout, "policy": "ActorMLP(51, 2, [100, 100]), torch.manual_seed(1), fp32",
                               "env_share_of_rollout_time": env_share,
                               "kernels_per_step": "policy (cuBLAS GEMMs + elementwise) + mdr::sample_actions_kernel + "
                                                   "mdr::step_pipe_kernel + 2 scalar copies"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def dict_api_arm(args, w, dev):
    """BASELINE config 0: the reference's deploy loop (main-deploy.py:99-104) on the drop-in dict API."""
    import random

    import numpy as np
    import torch

    import mdr_b200
    cfg = workload_config(w)
    random.seed(1)
    np.random.seed(1)
    env = mdr_b200.MADemandResponseEnv(cfg, precision=args.precision, device=dev)
    obs = env.reset()
    n = env.nb_agents
    act = lambda o: {k: o[k]["house_temp"] > o[k]["house_target_temp"] for k in o}   # agents/bangbang_controllers.py:50-61
    actions = act(obs)
    for _ in range(args.warmup):
        obs, _, _, _ = env.step(actions)
        actions = act(obs)
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        obs, _, _, info = env.step(actions)     # H2D of the actions, the step, D2H of the state, python dicts built
        actions = act(obs)
    ev1.record()
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    value = n * args.steps / (ms * 1e-3)
    rate, one, wall, worst = run_cpu_port(n, args.steps, min(args.warmup, 20), 1, False)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.isfile(peaks_path) else 6650.0
    algo = ALGO_BYTES[(args.precision, False)]
    achieved = algo * n / (ms * 1e-3 / args.steps) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic", "config": config_block(args, w),
        "launch": env._vec.launch_geometry(), "clocks": sampler.stop(t0, t1), "gpu_launches": args.steps,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n * (2 * 8 + 4 + 8) + 40,
                "steps": args.steps, "api": "MADemandResponseEnv.step(action_dict) -> (obs_dict, rewards, dones, info): the "
                                            "timed loop IS the host-buffer path (value == e2e)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "algorithmic_bytes_per_house_step": algo, "kernel": "mdr::step_kernel",
                     "note": "one 50-house cluster per launch: host- and launch-latency-bound, not a roofline configuration"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "1 process x 1 cluster x %d houses x %d steps of oracle/mdr_oracle_scalar.py (same loop: "
                                   "bang-bang per house + env.step + normStateDict), %.1f s" % (n, args.steps, worst)},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--prewarm-ms", type=float, default=0.0,
                    help="untimed: issue the step for this long before the W warm-up steps (clocks back at their loaded level)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "fp64"], help="default: the workload's (fp32; c0/c1: fp64)")
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=3000,
                    help="env steps per CPU-baseline process (default: ~3 s per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-phase", action="store_true", help="keep every cluster's interpolation clock in phase (all refresh on the "
                    "same step, once per 75 steps) instead of staggering them (round-1 behaviour, for comparison)")
    ap.add_argument("--solar", action="store_true", help="solar gain on (config.py's default; the reference CLIs force it off)")
    ap.add_argument("--comm-defect", type=float, default=0.0, help="comm_defect_prob > 0: message drops drawn on the device")
    ap.add_argument("--metrics", action="store_true", help="accumulate the deploy / training metrics on the device in every step")
    ap.add_argument("--serial-e2e", action="store_true", help="e2e leg without the MdrHostCtx pipeline (one H2D, one launch, four D2H)")
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
