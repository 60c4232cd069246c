#!/bin/bash
# round-2 analysis pass: role timing of the stack kernel, one-step launch list with DRAM counters, full ncu capture of the
# stack launches (with source), for the 16384-graph bench workload.
mkdir -p gpurun_out
STACK_TIMING_COMPACT=1 timeout 200 python tools/stack_timing.py 16384 tc 2>&1 | tail -4 | tee gpurun_out/r2b_stack_timing.txt
STACK_TIMING_COMPACT=1 timeout 200 python tools/stack_timing.py 16384 tc1x 2>&1 | tail -4 | tee gpurun_out/r2b_stack_timing_1x.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2b_launches.csv python tools/profile_step.py 4 > gpurun_out/r2b_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_stack -s 2 -c 2 -f -o gpurun_out/r2b_stack_prof python tools/profile_step.py 2 > gpurun_out/r2b_stack_prof.log 2>&1
echo "ncu stack rc=$?"; tail -2 gpurun_out/r2b_stack_prof.log
