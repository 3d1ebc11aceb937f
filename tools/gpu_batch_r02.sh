# One GPU call of round 2: full parity suite, drop-in fixture, every bench workload, the reference arm, c1 launch list.
tag=${1:-r02b}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
timeout 120 python tools/record_dropin_fixture.py gpurun_out/dropin_c0.pkl 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err
for w in c2 c2actor c1 c3 c3big c3fused; do
  timeout 200 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err
done
timeout 200 python bench.py --workload c0 --steps 1000 --warmup 50 > gpurun_out/bench_c0_$tag.json 2> gpurun_out/bench_c0_$tag.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_driver_$tag.json 2> gpurun_out/bench_c4_driver_$tag.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c1_$tag.csv \
    python bench.py --workload c1 --steps 50 --warmup 25 --no-cpu-baseline > gpurun_out/ncu_c1_$tag.log 2>&1
python - "$tag" <<'PY'
import json, sys
tag=sys.argv[1]
for w in ("c4","c4_driver","c2","c2actor","c1","c3","c3big","c3fused","c0","reference"):
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%(w,tag)))
        r=d.get("roofline") or {}
        print("%s value %.4g us/step %.2f e2e %.4g frac %s cpu %s %s" % (w, d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], r.get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("rollout","")))
    except Exception as ex: print(w,"ERR",ex)
PY
