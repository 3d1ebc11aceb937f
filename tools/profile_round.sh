# Round-end evidence run (one GPU, one gpurun call): parity suite, every bench workload, the reference arm, the ncu launch
# list and one `--set full` capture of the hot kernel of the default workload (summarised ON the box: the .ncu-rep files
# are too large to bring back), per-tile / per-CTA traces, host expansion bandwidth.  Usage: bash tools/profile_round.sh r02
tag=${1:-r02}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
timeout 400 python bench.py > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err
timeout 300 python bench.py --in-phase --no-cpu-baseline > gpurun_out/bench_c4_inphase_$tag.json 2> gpurun_out/bench_c4_inphase_$tag.err
# the pipelined kernel's other schedule (fixed strided tile list per CTA instead of in-order claiming), same box
MDR_STATIC_TILES=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c4_strided_$tag.json 2> gpurun_out/bench_c4_strided_$tag.err
MDR_STATIC_TILES=1 timeout 300 python bench.py --in-phase --no-cpu-baseline > gpurun_out/bench_c4_strided_inphase_$tag.json 2> gpurun_out/bench_c4_strided_inphase_$tag.err
timeout 300 python bench.py --serial-e2e --no-cpu-baseline --steps 100 --warmup 10 > gpurun_out/bench_c4_serial_e2e_$tag.json 2> gpurun_out/bench_c4_serial_e2e_$tag.err
for w in c2 c2actor c1 c3 c3big c3fused; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err
done
timeout 300 python bench.py --workload c0 --steps 1000 --warmup 50 > gpurun_out/bench_c0_$tag.json 2> gpurun_out/bench_c0_$tag.err
timeout 300 python bench.py --workload c1 --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_c1_$tag.json 2> gpurun_out/bench_reference_c1_$tag.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_driver_$tag.json 2> gpurun_out/bench_c4_driver_$tag.err
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -4 gpurun_out/smoke_$tag.log
timeout 120 python tools/record_dropin_fixture.py gpurun_out/dropin_c0.pkl 2>&1 | tail -1
timeout 300 python tools/launch_probe.py c1 c3big c3 > gpurun_out/launch_probe_$tag.log 2>&1
# ncu: launch list of the default bench command, then full captures summarised here
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 30 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_pipe_kernel -s 45 -c 2 -o /tmp/prof_c4 -f \
    python bench.py --steps 30 --warmup 30 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
python tools/summarise_ncu.py /tmp/prof_c4.ncu-rep gpurun_out/ncu_full_c4_$tag.csv c4_fp32
for w in c2 c3 c3big c1; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 2 -o /tmp/prof_$w -f \
      python tools/launch_probe.py $w > /tmp/ncu_$w.log 2>&1
  python tools/summarise_ncu.py /tmp/prof_$w.ncu-rep gpurun_out/ncu_full_${w}_$tag.csv ${w}_fp32
done
cp profiles/traffic.json gpurun_out/traffic_$tag.json 2>/dev/null
# traces (needs variants/lib_trace.so: nvcc ... -DMDR_TRACE -DMDR_TRACE_CTA=100)
if [ -f variants/lib_trace.so ]; then
  for w in c4 c2 c3big; do MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py $w > gpurun_out/trace_${w}_$tag.log 2>&1; done
  MDR_STATIC_TILES=1 MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py c4 > gpurun_out/trace_c4_strided_$tag.log 2>&1
  MDR_TRACE_STAGGER=0 MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py c4 > gpurun_out/trace_c4_inphase_$tag.log 2>&1
fi
g++ -O3 -std=c++17 -pthread -o /tmp/expand_bw tools/microbench/expand_bw.cpp 2>/dev/null && {
  nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread" ;
  for t in 1 4 8 15 16 32; do /tmp/expand_bw $t | tail -2 | head -1; done
  for t in 8 16; do MDR_HOST_NT=0 /tmp/expand_bw $t | tail -2 | head -1; done
} > gpurun_out/expand_bw_$tag.log 2>&1
python - "$tag" <<'PY'
import json, sys
tag=sys.argv[1]
for w in ("c4","c4_inphase","c4_strided","c4_strided_inphase","c4_serial_e2e","c4_driver","c2","c2actor","c1","c3","c3big","c3fused","c0","reference","reference_c1"):
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%(w,tag)))
        r=d.get("roofline") or {}
        print("%s value %.4g us/step %.2f e2e %.4g frac %s cpu %s %s" % (w, d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], r.get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("rollout","")))
    except Exception as ex: print(w,"ERR",ex)
PY
