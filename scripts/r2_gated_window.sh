set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "window_gated or gated_local_large or gated_large_tree" 2>&1 | tail -15 > gpurun_out/r2_gw_tests.log
cat gpurun_out/r2_gw_tests.log
timeout 600 python scripts/gated_throughput.py 16384 8 ${1:-1,2,3} > gpurun_out/r2_gw_throughput.txt 2>&1
cat gpurun_out/r2_gw_throughput.txt
