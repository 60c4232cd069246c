#!/bin/bash
# same-box A/B of library builds / switches: each argument is an env list for one short bench run
mkdir -p gpurun_out
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4), {k: ks.get(k) for k in ("stack_fwd", "stack_bwd", "dw_layers", "encoder_fwd")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra --skip-strong > gpurun_out/r2h_$i.json 2> gpurun_out/r2h_$i.err
  short gpurun_out/r2h_$i.json "$envs"
done
