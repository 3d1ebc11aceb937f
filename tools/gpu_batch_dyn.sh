timeout 900 python -m pytest tests/test_pipe_steady_gpu.py tests/test_parity_gpu.py tests/test_fused_gpu.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab_env.sh "- MDR_STATIC_TILES=1" "c4 c4:--in-phase"
MDR_LIB_PATH=$PWD/variants/lib_trace.so timeout 200 python tools/trace_tile.py c4 > gpurun_out/trace_c4_dyn_stag.log 2>&1
