#!/bin/bash
# small-batch A/B: usage gpu_r2_small.sh "<env list>" ... ; each run at 2048 and 4096 graphs
mkdir -p gpurun_out
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4), ks)
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
i=0
for envs in "$@"; do
  for b in 2048 4096; do
    i=$((i+1))
    env $envs timeout 200 python bench.py --batch $b --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-extra --skip-strong > gpurun_out/r2f_small_$i.json 2> gpurun_out/r2f_small_$i.err
    short gpurun_out/r2f_small_$i.json "B=$b $envs"
  done
done
