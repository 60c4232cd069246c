#!/bin/bash
# One GPU-box visit: parity tests, launch list, full ncu captures of the hot kernels, bench.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
if [ "${1:-}" != "noprof" ]; then
  # steady-state step = launches [2*LPS, 3*LPS): list every launch of it
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py 4 > gpurun_out/launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_rowgemm -s 30 -c 6 -f -o gpurun_out/prof_rowgemm \
      python tools/profile_step.py 2 > gpurun_out/prof_rowgemm.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_encoder|k_tc_reducegemm' -s 9 -c 3 -f -o gpurun_out/prof_enc \
      python tools/profile_step.py 2 > gpurun_out/prof_enc.log 2>&1
fi
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-400 gpurun_out/bench_ref.json
