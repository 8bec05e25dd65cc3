OLD=$PWD/structure-optimization_b200/_ab/libso_old.so
OUT=gpurun_out/${1:-r2_gated_ab}.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "window_gated or gated_local_large or gated_large_tree" 2>&1 | tail -3 >> $OUT
for i in 1 2; do
  SO_B200_LIB=$OLD python scripts/gated_throughput.py 16384 8 2 2>&1 | grep "sweep" | sed 's/^/old /' >> $OUT
  python scripts/gated_throughput.py 16384 8 2 2>&1 | grep "sweep" | sed 's/^/new /' >> $OUT
done
python scripts/gated_throughput.py 16384 8 1 2>&1 | grep "sweep" | sed 's/^/new /' >> $OUT
cat $OUT
