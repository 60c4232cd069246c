#!/bin/bash
# stack tests (bounded) + ablation runs of the CTA-pair kernel; usage: gpu_r2_quickab.sh <debug masks...>
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stack.py -x -q -m gpu > gpurun_out/r2c_stack_tests.log 2>&1; rc=$?; tail -3 gpurun_out/r2c_stack_tests.log
if [ $rc -ne 0 ]; then echo "stack tests failed"; grep -n "Error\|assert" gpurun_out/r2c_stack_tests.log | head; exit 1; fi
bash tools/gpu_r2_ablate.sh "$@"
