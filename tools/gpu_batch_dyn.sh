timeout 900 python -m pytest tests/test_pipe_steady_gpu.py tests/test_parity_gpu.py tests/test_fused_gpu.py -m gpu -x -q 2>&1 | tail -3
bash tools/ab_env.sh "- MDR_STATIC_TILES=1" "c4 c4:--in-phase c2 c3"
bash tools/ab_env.sh "MDR_DYN_MIN_TILES=4" "c2 c3"
for m in 0; do MDR_STATIC_TILES=$m MDR_TRACE_STAGGER=0 MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py c4 > gpurun_out/trace_c4_dyn_static$m.log 2>&1; done
MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py c4 > gpurun_out/trace_c4_dyn_stag.log 2>&1
ncu -k regex:"env_pro|step_pipe" --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/launches_dyn.csv python bench.py --workload c4 --steps 10 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
