"""Summarises an ncu report (--set full) of the step kernel into profiles/: headline raw metrics per launch,
warp-stall totals and the hottest SASS instructions with their stall reasons.
Usage: python tools/summarise_ncu.py gpurun_out/prof_c4_r01.ncu-rep profiles/ncu_full_c4_r01.csv [traffic_key]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
want += [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
lines = [["metric", "unit"] + ["launch%d" % i for i in range(len(data))]]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append([w, units[i]] + [r[i] for r in data])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
if starts:
    s = starts[0]
    shdr = srows[s + 1]
    end = starts[1] if len(starts) > 1 else len(srows)
    body = srows[s + 2:end]
    ia = shdr.index("Warp Stall Sampling (All Samples)")
    ie = shdr.index("Instructions Executed")
    cols = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines.append([])
    lines.append(["hottest SASS instructions of launch0 (index, instruction, samples, executed, stall reasons > 5 samples)"])
    top = sorted(range(len(body)), key=lambda i: -int(body[i][ia]))[:30]
    for i in sorted(top):
        r = body[i]
        lines.append([str(i), r[1].strip(), r[ia], r[ie], " ".join("%s=%s" % (shdr[c][6:], r[c]) for c in cols if int(r[c]) > 5)])
with open(out, "w", newline="") as f:
    csv.writer(f).writerows(lines)
if key:
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    tr = sum((float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]) for r in data) / len(data)
    tpath = os.path.join(os.path.dirname(out), "traffic.json")
    t = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    t[key] = tr
    t.pop("source", None)
    t.setdefault("sources", {})[key] = ("%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches "
                                        "of the step kernel)" % (os.path.basename(out), len(data)))
    json.dump(t, open(tpath, "w"), indent=1)
print("wrote", out)
