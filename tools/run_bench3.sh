# pytest -m gpu + the three bench workloads; prints a one-line summary per workload
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tag=${1:-cur}
for w in c4 c2 c3; do
  python bench.py --workload $w --steps 600 --warmup 50 --no-cpu-baseline > gpurun_out/bench_${w}_${tag}.json 2> gpurun_out/bench_${w}_${tag}.err
done
python - "$tag" <<'PY'
import json, sys
tag=sys.argv[1]
for w in ("c4","c2","c3"):
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%(w,tag))); print("%s %.1fus frac %.3f value %.3g e2e %.3g"%(w,d["roofline"]["launch_us"],d["roofline"]["frac"],d["value"],d["e2e"]["value"]))
    except Exception as ex: print(w,"ERR",ex)
PY
