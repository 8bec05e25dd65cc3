set -x
python scripts/profile_resident.py > gpurun_out/r2c_res_gated_plain.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sa_resident -s 1 -c 1 -f -o gpurun_out/r2c_res_gated python scripts/profile_resident.py > gpurun_out/r2c_res_gated_ncu.log 2>&1
tail -2 gpurun_out/r2c_res_gated_plain.txt
