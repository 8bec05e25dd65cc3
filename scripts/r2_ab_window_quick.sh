# same-box A/B of two builds of the library on the headline kernel (old = structure-optimization_b200/_ab/libso_old.so)
OLD=$PWD/structure-optimization_b200/_ab/libso_old.so
OUT=gpurun_out/${1:-r2_ab_quick}.txt
: > $OUT
python -m pytest tests -m gpu -x -q -k "window or compact or golden or near_tie or accept_vectors or config4" 2>&1 | tail -3 >> $OUT
for i in 1 2 3; do
  SO_B200_LIB=$OLD python scripts/window_probe.py 32768 3 2>&1 | grep "sweep 2" | sed 's/^/old /' >> $OUT
  python scripts/window_probe.py 32768 3 2>&1 | grep "sweep 2" | sed 's/^/new /' >> $OUT
done
SO_B200_LIB=$OLD python scripts/window_probe.py 32768 2 1.55 2>&1 | grep "sweep 1" | sed "s/^/old T=1.55 /" >> $OUT
python scripts/window_probe.py 32768 2 1.55 2>&1 | grep "sweep 1" | sed "s/^/new T=1.55 /" >> $OUT
cat $OUT
