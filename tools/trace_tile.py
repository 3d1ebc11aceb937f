"""Per-tile timeline of one CTA of the pipelined step kernel (trace build of the library:
nvcc ... -rdc=true -DMDR_TRACE -o variants/lib_trace.so; run with MDR_LIB_PATH pointing at it)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import mdr_b200

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[name]
cfg = bench.workload_config(w)
flat = mdr_b200.FlatConfig(cfg)
E, N = w["envs"], w["houses"]
pop = mdr_b200.synthetic_population(flat, E, seed=1234)
table = mdr_b200.synthetic_interp_table() if w["interp"] else None
env = mdr_b200.VecDemandResponseEnv(cfg, pop, precision="fp32", device="cuda:0", seed=1234, interp_table=table,
                                    action_source=w["action_source"], with_obs=w["obs"])
env.reset_tensor()
if w["interp"] and os.environ.get("MDR_TRACE_STAGGER", "1") == "1":
    env.stagger_interp_clock(seed=77)   # like bench.py: every step refreshes ~1/75 of the clusters
act = (torch.rand(E, N, device="cuda:0") < 0.5).to(torch.uint8)
a = act if w["action_source"] == "array" else None
for _ in range(20):
    env.step_tensor(a)
TILES, POINTS = 24, 10
n = 8 * TILES * POINTS
buf = (C.c_ulonglong * n)()
lib = env.lib
lib.mdr_debug_trace.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
lib.mdr_debug_trace(buf, n, 1)
env.step_tensor(a)
lib.mdr_debug_trace(buf, n, 0)
t = np.frombuffer(buf, dtype=np.uint64).reshape(8, TILES, POINTS).astype(np.int64)
t0 = t[t > 0].min()
print("absolute ns: first stamp %d, last stamp %d (span %.1f us)" % (t0, t.max(), (t.max() - t0) / 1e3))
rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
names = ["tile start", "issued next+record", "inputs landed", "phase A done", "drain done", "after barrier", "rows done",
         "tile end"]
np.set_printoptions(precision=1, suppress=True, linewidth=200)
for wi in (0, 3, 6):
    print("house warp %d (us since first stamp): columns = %s" % (wi, ", ".join(names)))
    print(rel[wi, :, :8])
    if np.isfinite(rel[wi, :, 9]).any():
        print("  ... within 'issued next+record': [tile start, ring slot of the next tile ready, next tile's copies issued]")
        print(rel[wi, :12][:, [0, 8, 9]])
if np.isfinite(rel[0, 22, 6]):
    print("per-env prologue of all warps (in-order claiming), warp 0, us since its start: "
          "[start, loads back, calendar, draws + shuffles, outdoor temperature (sinpi), due logic, signal | all stores out]")
    pr = t[0, 22, :7].astype(np.float64)
    print(np.round((pr - pr[0]) / 1e3, 2))
if np.isfinite(rel[0, 23, 5]):
    print("own-tile refresh, warps 0 / 3 / 6, us since the refresh started: [start, table walk done, after barrier 1, "
          "signal written (after barrier 2), rows landed, after barrier 3]")
    for wi in (0, 3, 6):
        rr = t[wi, 23, :6].astype(np.float64)
        print(np.round((rr - t[0, 23, 0]) / 1e3, 2))
print("prologue warp: per pass starting at tile it0: [wait for slot, start, end]")
print(rel[7][:, [8, 0, 7]])

# per-CTA span of the traced launch (every CTA): entry, past griddepcontrol.wait, tile loop done, exit
if hasattr(lib, "mdr_debug_cta_span"):
    span = (C.c_ulonglong * (2048 * 8))()
    lib.mdr_debug_cta_span.argtypes = [C.c_void_p, C.c_size_t]
    lib.mdr_debug_cta_span(span, 2048 * 8)
    sp = np.frombuffer(span, dtype=np.uint64).reshape(2048, 8).astype(np.int64)
    sp = sp[sp[:, 0] > 0]
    base = sp[:, 0].min()
    r = (sp - base) / 1e3
    print("CTAs traced: %d; kernel span %.1f us (first entry -> last exit)" % (len(sp), r[:, 3].max()))
    for k, name in enumerate(["entry", "past griddepcontrol.wait", "tile loop done", "exit"]):
        q = np.percentile(r[:, k], [0, 10, 50, 90, 100])
        print("  %-26s min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f us" % (name, *q))
    busy = r[:, 2] - r[:, 1]
    print("  tile loop duration per CTA: min %.1f median %.1f max %.1f us; refresh pass: median %.1f max %.1f us"
          % (busy.min(), np.median(busy), busy.max(), np.median(r[:, 3] - r[:, 2]), (r[:, 3] - r[:, 2]).max()))
    for i in np.argsort(r[:, 3] - r[:, 2])[-6:]:
        print("  CTA %4d refresh pass: loop done %.1f | list ready +%.1f | first tile: thread 0 evaluated +%.1f, all evaluated +%.1f, "
              "patched +%.1f | exit +%.1f" % (i, r[i, 2], r[i, 4] - r[i, 2], r[i, 5] - r[i, 2], r[i, 6] - r[i, 2], r[i, 7] - r[i, 2],
                                            r[i, 3] - r[i, 2]))
    late = np.argsort(r[:, 3])[-8:]
    print("  last CTAs to exit (blockIdx, entry, start, loop done, exit):", [(int(i), *np.round(r[i], 1)) for i in late])
