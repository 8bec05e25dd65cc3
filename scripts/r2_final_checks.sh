# full GPU test suite, smoke, default bench line, reference arm (bounded)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2f_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> gpurun_out/r2f_bench.err
cat gpurun_out/r2f_gpu_tests.log; tail -12 gpurun_out/r2f_smoke.log; tail -3 gpurun_out/r2f_bench.err; tail -c 1500 gpurun_out/r2f_bench.json
