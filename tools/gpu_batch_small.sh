# quick GPU iteration: parity suite + the main bench lines.  Usage: bash tools/gpu_batch_small.sh <tag> [workloads]
tag=${1:-x}
wl=${2:-"c4 c2 c1 c3 c3big"}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
for w in $wl; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_driver_$tag.json 2> gpurun_out/bench_c4_driver_$tag.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --serial-e2e > gpurun_out/bench_c4_serial_$tag.json 2> gpurun_out/bench_c4_serial_$tag.err
python - "$tag" $wl c4_driver c4_serial <<'PY'
import json, sys
tag=sys.argv[1]
for w in sys.argv[2:]:
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%(w,tag)))
        r=d.get("roofline") or {}
        print("%s value %.4g us/step %.2f e2e %.4g frac %.3f d2h %s pipe %s" % (w, d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], r.get("frac"), d["e2e"].get("d2h_bytes_per_step"), d["e2e"].get("pipeline")))
    except Exception as ex: print(w,"ERR",ex, open("gpurun_out/bench_%s_%s.err"%(w,tag)).read()[-600:])
PY
# host expansion bandwidth of this box (compact observation records -> [E,N,51] rows)
g++ -O3 -std=c++17 -pthread -o /tmp/expand_bw tools/microbench/expand_bw.cpp 2>/dev/null && {
  nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread" ;
  for t in 1 4 8 16 32; do /tmp/expand_bw $t | tail -2 | head -1; done
  for t in 8 16; do MDR_HOST_NT=0 /tmp/expand_bw $t | tail -2 | head -1; done
} > gpurun_out/expand_bw_$tag.log 2>&1
cat gpurun_out/expand_bw_$tag.log
