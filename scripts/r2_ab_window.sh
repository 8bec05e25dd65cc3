# same-box A/B of the window kernel: parity tests that run through it, then alternating sweeps old / new library
set -x
python -m pytest tests -m gpu -x -q -k "window or compact or flat_local or near_tie or golden or config2 or config4 or tscale or checkpoint or type_hist or accept_vectors or variants_agree or philox or quench" 2>&1 | tail -8 > gpurun_out/r2_ab_tests.log
cat gpurun_out/r2_ab_tests.log
OLD=$PWD/structure-optimization_b200/_ab/libso_old.so
for i in 1 2 3; do
  SO_B200_LIB=$OLD python scripts/window_probe.py 32768 3 2>&1 | grep sweep | sed 's/^/old /' >> gpurun_out/r2_ab_window.txt
  python scripts/window_probe.py 32768 3 2>&1 | grep sweep | sed 's/^/new /' >> gpurun_out/r2_ab_window.txt
done
for T in 1.55 0.87; do
  SO_B200_LIB=$OLD python scripts/window_probe.py 32768 2 $T 2>&1 | grep "sweep 1" | sed "s/^/old T=$T /" >> gpurun_out/r2_ab_window.txt
  python scripts/window_probe.py 32768 2 $T 2>&1 | grep "sweep 1" | sed "s/^/new T=$T /" >> gpurun_out/r2_ab_window.txt
done
cat gpurun_out/r2_ab_window.txt
