#!/bin/bash
# strong-scaling block (per-rank compute of a 16384-graph global batch at 2 / 4 / 8 GPUs, measured on one GPU) for a list of env settings
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env ${envs//,/ } timeout 300 python bench.py --skip-cpu --skip-e2e --skip-extra --steps 10 > gpurun_out/strong_v_$i.json 2>gpurun_out/strong_v_$i.err
  python - "$envs" gpurun_out/strong_v_$i.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
rows = d["strong"]["per_rank_compute_at_n_gpus"]
print(f"{sys.argv[1]:24s} step {d['ms_per_step']:.4f}", " | ".join(f"B={r['per_gpu_batch']} eager {r['ms_per_step']:.4f} graph {r['ms_per_step_cuda_graph']:.4f}" for r in rows))
PY
done
