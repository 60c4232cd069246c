set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_encoder_stream' -s 1 -c 1 -f -o gpurun_out/prof_encq python tools/profile_step.py 2 > gpurun_out/prof_encq.log 2>&1
tail -2 gpurun_out/prof_encq.log
