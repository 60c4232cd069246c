#!/bin/bash
# same-box A/B at small batches: each argument is an env list; runs at 2048 and 16384 graphs
mkdir -p gpurun_out
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "launches", d["gpu_launches"], {k: ks.get(k) for k in ("reduce_partials", "derive_weights", "memset", "optimizer", "decoder_bwd", "stack_fwd", "stack_bwd")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
i=0
for envs in "$@"; do
  for b in 2048 16384; do
    i=$((i+1))
    env $envs timeout 200 python bench.py --batch $b --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-extra --skip-strong > gpurun_out/r2m_$i.json 2> gpurun_out/r2m_$i.err
    short gpurun_out/r2m_$i.json "B=$b $envs"
  done
done
