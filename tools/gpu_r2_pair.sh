#!/bin/bash
# CTA-pair stack kernel: correctness first (bounded by timeout: a wrong barrier protocol hangs), then A/B timing against the one-CTA kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stack.py -x -q -m gpu -k "cta_pair" > gpurun_out/r2c_pair_tests.log 2>&1; rc=$?
tail -15 gpurun_out/r2c_pair_tests.log
if [ $rc -ne 0 ]; then echo "pair tests failed (rc=$rc)"; exit 1; fi
timeout 600 python -m pytest tests/test_gpu_stack.py -x -q -m gpu > gpurun_out/r2c_stack_tests.log 2>&1; echo "stack tests rc=$?"; tail -3 gpurun_out/r2c_stack_tests.log
STACK_TIMING_COMPACT=1 timeout 200 python tools/stack_timing.py 16384 tc 2>&1 | tail -3
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4),
          {k: ks.get(k) for k in ("stack_fwd", "stack_bwd", "dw_layers", "encoder_fwd", "dw_encoder")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
i=0
for envs in "MSHGNN_STACK_2CTA=1" "MSHGNN_STACK_2CTA=0" "MSHGNN_STACK_2CTA=1" "MSHGNN_STACK_2CTA=0" "$@"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2c_pair_$i.json 2> gpurun_out/r2c_pair_$i.err
  short gpurun_out/r2c_pair_$i.json "$envs"
done
