# final captures of round 2: every ncu command after a plain run of the same command that exited 0
set -x
python bench.py --steps 2 --warmup 3 --cpu-flips 0 --e2e-steps 0 --accept-sweep 0 --other-configs 0 > gpurun_out/r2c_traffic_plain.json 2> gpurun_out/r2c_traffic_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sa_window -s 3 -c 2 --csv --log-file gpurun_out/r2c_traffic_bench_shape_packed.csv python bench.py --steps 2 --warmup 3 --cpu-flips 0 --e2e-steps 0 --accept-sweep 0 --other-configs 0 > gpurun_out/r2c_traffic.log 2>&1
bash scripts/r2_capture_window.sh r2c_window24
bash scripts/r2_capture_gated.sh r2c_gw 2
python bench.py --chains 8192 --steps 2 --warmup 3 --cpu-flips 0 > gpurun_out/r2c_launches_plain.json 2> gpurun_out/r2c_launches_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --chains 8192 --steps 2 --warmup 3 --cpu-flips 0 > gpurun_out/r2c_launches.log 2>&1
tail -4 gpurun_out/r2c_traffic_bench_shape_packed.csv
