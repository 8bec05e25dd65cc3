set -x
python scripts/profile_resident.py > gpurun_out/r2_res_gated_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_resident -s 1 -c 1 -o gpurun_out/r2_res_gated python scripts/profile_resident.py > gpurun_out/r2_res_gated_ncu.log 2>&1
python scripts/profile_gate_index.py > gpurun_out/r2_gate_index_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_resident -s 2 -c 1 -o gpurun_out/r2_gate_index python scripts/profile_gate_index.py > gpurun_out/r2_gate_index_ncu.log 2>&1
python scripts/window_probe.py 8192 3 > gpurun_out/r2_window_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_window -s 2 -c 1 -o gpurun_out/r2_window python scripts/window_probe.py 8192 3 > gpurun_out/r2_window_ncu.log 2>&1
python scripts/descriptor_probe.py 96 65536
cat gpurun_out/r2_res_gated_plain.txt gpurun_out/r2_gate_index_plain.txt gpurun_out/r2_window_plain.txt
