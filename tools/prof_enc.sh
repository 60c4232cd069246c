set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_tc_encoder2$' -s 1 -c 1 -f -o gpurun_out/prof_enc2 python tools/profile_step.py 2 > gpurun_out/prof_enc2.log 2>&1
tail -2 gpurun_out/prof_enc2.log
MSHGNN_ENCODER=v1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_tc_encoder$' -s 1 -c 1 -f -o gpurun_out/prof_enc1 python tools/profile_step.py 2 > gpurun_out/prof_enc1.log 2>&1
tail -2 gpurun_out/prof_enc1.log
