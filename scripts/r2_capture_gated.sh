set -x
TAG=${1:-r2_gw}
CASE=${2:-2}
python scripts/gated_throughput.py 8192 8 $CASE > gpurun_out/${TAG}_plain.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sa_window_gated -s 2 -c 1 -f -o gpurun_out/${TAG} python scripts/gated_throughput.py 8192 8 $CASE > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_plain.txt
