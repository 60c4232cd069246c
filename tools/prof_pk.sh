set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_rowgemm_persistent -s 60 -c 6 -f -o gpurun_out/prof_pk python tools/profile_step.py 3 > gpurun_out/prof_pk.log 2>&1
tail -2 gpurun_out/prof_pk.log
cp morphsym-hgnn_b200/lib/libmshgnn_b200.so gpurun_out/libmshgnn_b200.prof.so
