set -x
python scripts/analytical_throughput.py > gpurun_out/r2_analytical.txt 2>&1
python scripts/gated_throughput.py 65536 8 > gpurun_out/r2_gated_base.txt 2>&1
python scripts/small_lattice_throughput.py > gpurun_out/r2_small_base.txt 2>&1
# bench-shape traffic of the hot kernel (65,536 chains), counters only
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sa_window -s 3 -c 2 --csv --log-file gpurun_out/r2_traffic_bench_shape.csv python bench.py --steps 2 --warmup 3 --cpu-flips 0 --e2e-steps 0 --accept-sweep 0 > gpurun_out/r2_traffic_bench_shape.log 2>&1
# descriptor kernels
for d in 12 96; do
  n=65536; if [ $d = 12 ]; then n=1048576; fi
  python scripts/descriptor_probe.py $d $n > gpurun_out/r2_desc_plain_$d.txt 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"descriptors_bits|flip_delta|descriptors_kernel" -c 12 -o gpurun_out/r2_desc_$d python scripts/descriptor_probe.py $d $n > gpurun_out/r2_desc_ncu_$d.log 2>&1
done
# launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --chains 8192 --steps 2 --warmup 3 --cpu-flips 0 > gpurun_out/r2_launches.log 2>&1
tail -3 gpurun_out/r2_analytical.txt gpurun_out/r2_gated_base.txt gpurun_out/r2_desc_plain_*.txt; tail -5 gpurun_out/r2_traffic_bench_shape.csv
