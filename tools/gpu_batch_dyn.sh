timeout 900 python -m pytest tests/test_large_cluster_gpu.py tests/test_properties_gpu.py tests/test_fuzz_gpu.py tests/test_fused_gpu.py tests/test_montecarlo_gpu.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab_env.sh "-" "c3big c3"
timeout 200 python tools/simulate_day.py 2>&1 | tail -4
