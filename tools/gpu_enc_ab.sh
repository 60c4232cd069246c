#!/bin/bash
# same-box A/B, alternating, three repeats each.  usage: gpu_enc_ab.sh "ENV=a" "ENV=b" ...
mkdir -p gpurun_out
for r in ${REPEATS:-1 2 3}; do
  i=0
  for envs in "$@"; do
    i=$((i+1))
    env $envs timeout 300 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/ab_$i$r.json 2>/dev/null
    python - "$envs" gpurun_out/ab_$i$r.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
k = {x['kernel']: x['ms_per_step'] for x in d['kernels']}
print(f"{sys.argv[1]:28s} step {d['ms_per_step']:.4f} inf {d['inference']['ms_per_step']:.4f} enc_fwd {k['encoder_fwd']:.4f} dw_enc {k['dw_encoder']:.4f} conv {k['conv_fwd']:.4f} dx {k['dx_bwd']:.4f} dw {k['dw_layers']:.4f}")
PY
  done
done
