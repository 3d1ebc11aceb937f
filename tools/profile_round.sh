# Round-end evidence run (one GPU): parity tests, the three bench workloads with the CPU baseline on the default one,
# the reference arm, then the ncu launch list and one `--set full` capture of the hot kernel of the default workload.
# Usage: bash tools/profile_round.sh r01
tag=${1:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err
python bench.py --workload c3fused --steps 1500 --warmup 150 --no-cpu-baseline > gpurun_out/bench_c3fused_$tag.json 2> gpurun_out/bench_c3fused_$tag.err
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 30 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_pipe -s 45 -c 2 -o gpurun_out/prof_c4_$tag -f \
    python bench.py --steps 30 --warmup 30 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
cat gpurun_out/pytest_gpu_$tag.log; tail -2 gpurun_out/smoke_$tag.log
python - "$tag" <<'PY'
import json, sys
tag=sys.argv[1]
for w in ("c4","c2","c3","c3fused","reference"):
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%(w,tag)))
        r=d.get("roofline") or {}
        print("%s value %.4g e2e %.4g frac %s us %s cpu %s" % (w, d["value"], d["e2e"]["value"], r.get("frac"), r.get("launch_us"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as ex: print(w,"ERR",ex)
PY
