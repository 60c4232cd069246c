#!/bin/bash
# same-box timing of encoder variants (no tests: the debug switches give wrong results).  usage: gpu_r2_encq_ab.sh "ENV=a,ENV2=b" ...   (commas separate variables of one run)
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env ${envs//,/ } timeout 300 python bench.py --skip-cpu --skip-e2e --skip-extra --skip-strong --steps 10 > gpurun_out/encq_v_$i.json 2>gpurun_out/encq_v_$i.err
  python - "$envs" gpurun_out/encq_v_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    k = {x['kernel']: x for x in d['kernels']}
    e = k['encoder_fwd']
    print(f"{sys.argv[1]:44s} step {d['ms_per_step']:.4f} inf {d['inference']['ms_per_step']:.4f} enc_fwd {e['ms_per_step']:.4f} hbm {e.get('frac_hbm_peak'):.3f} dw_enc {k['dw_encoder']['ms_per_step']:.4f}")
except Exception as ex:
    print(sys.argv[1], "failed", ex)
PY
done
