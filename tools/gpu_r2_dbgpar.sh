#!/bin/bash
for envs in "MSHGNN_STACK_PRIVATE_DH=0" "MSHGNN_STACK_PRIVATE_DH=1" "MSHGNN_STACK_PRIVATE_DH=1 MSHGNN_STACK_2CTA=2" "MSHGNN_STACK_PRIVATE_DH=0 MSHGNN_STACK_2CTA=2"; do
  echo "== $envs"
  env $envs timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "tensor_core_mode" 2>&1 | grep -E "^parity|passed|failed" | cut -c1-220
done
