set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
for d in 12 96; do
  n=65536; if [ $d = 12 ]; then n=1048576; fi
  ncu --set full --clock-control none --import-source on -k regex:flip_delta -c 2 -o gpurun_out/r2_flipdelta_$d python scripts/descriptor_probe.py $d $n > gpurun_out/r2_flipdelta_ncu_$d.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:descriptors_kernel -c 1 -o gpurun_out/r2_desckernel_$d python scripts/descriptor_probe.py $d $n > gpurun_out/r2_desckernel_ncu_$d.log 2>&1
done
cat gpurun_out/r2_gpu_tests.log; tail -c 3000 gpurun_out/r2_bench_a.json
