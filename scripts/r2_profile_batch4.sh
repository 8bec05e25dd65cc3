set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests2.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sa_window -s 3 -c 2 --csv --log-file gpurun_out/r2_traffic_bench_shape_packed.csv python bench.py --steps 2 --warmup 3 --cpu-flips 0 --e2e-steps 0 --accept-sweep 0 --other-configs 0 > gpurun_out/r2_traffic_packed.log 2>&1
python scripts/window_probe.py 8192 3 > /dev/null 2>&1
cat gpurun_out/r2_gpu_tests2.log; tail -c 2500 gpurun_out/r2_bench_c.json; tail -4 gpurun_out/r2_traffic_bench_shape_packed.csv
