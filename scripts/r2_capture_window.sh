# plain run first, then one --set full capture of the third sweep of the packed window kernel (8,192 chains, d = 96)
set -x
TAG=${1:-r2b_window24}
python scripts/window_probe.py 8192 3 > gpurun_out/${TAG}_plain.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sa_window -s 2 -c 1 -f -o gpurun_out/${TAG} python scripts/window_probe.py 8192 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_plain.txt
