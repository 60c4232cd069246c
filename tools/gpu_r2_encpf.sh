#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stack.py -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; rc=$?; tail -3 gpurun_out/r2g_tests.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/r2g_tests.log | head; exit 1; fi
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4), {k: ks.get(k) for k in ("encoder_fwd", "dw_encoder", "stack_fwd", "stack_bwd", "dw_layers")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra --skip-strong > gpurun_out/r2g_$i.json 2> gpurun_out/r2g_$i.err
  short gpurun_out/r2g_$i.json "$envs"
done
