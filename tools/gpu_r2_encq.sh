#!/bin/bash
# persistent encoder: bounded parity tests, then same-box alternating A/B against the one-item-per-CTA kernel.  usage: gpu_r2_encq.sh [tests]
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py > gpurun_out/encq_smoke.log 2>&1; rc=$?; tail -4 gpurun_out/encq_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc"; exit 1; fi
timeout ${TEST_TIMEOUT:-150} python -m pytest ${1:-tests/test_gpu_parity.py tests/test_gpu_stack.py} -x -q -m gpu > gpurun_out/encq_tests.log 2>&1; rc=$?; tail -3 gpurun_out/encq_tests.log
if [ $rc -ne 0 ]; then echo "tests failed rc=$rc"; grep -n "Error\|assert\|FAILED" gpurun_out/encq_tests.log | head -20; exit 1; fi
for r in ${REPEATS:-1 2}; do
  i=0
  for envs in "MSHGNN_ENCODER=pair" "MSHGNN_ENCODER=stream" ${EXTRA_ENVS:-}; do
    i=$((i+1))
    env $envs timeout 300 python bench.py --skip-cpu --skip-e2e --skip-extra --skip-strong > gpurun_out/encq_ab_$i$r.json 2>gpurun_out/encq_ab_$i$r.err
    python - "$envs" gpurun_out/encq_ab_$i$r.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = {x['kernel']: x for x in d['kernels']}
e = k['encoder_fwd']
print(f"{sys.argv[1]:28s} step {d['ms_per_step']:.4f} inf {d['inference']['ms_per_step']:.4f} enc_fwd {e['ms_per_step']:.4f} hbm {e.get('frac_hbm_peak')} dw_enc {k['dw_encoder']['ms_per_step']:.4f} stack_fwd {k['stack_fwd']['ms_per_step']:.4f}")
PY
  done
done
