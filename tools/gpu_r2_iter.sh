#!/bin/bash
# quick iteration: stack tests, role timing, short bench runs.  usage: tools/gpu_r2_iter.sh ["ENV=val ENV2=val" ...] (one bench run per argument)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stack.py -x -q -m gpu > gpurun_out/r2_it_tests.log 2>&1; rc=$?
tail -4 gpurun_out/r2_it_tests.log
if [ $rc -ne 0 ]; then echo "stack tests failed (rc=$rc): skipping the timed runs"; grep -n "Error\|error" gpurun_out/r2_it_tests.log | head -10; exit 1; fi
timeout 200 python tools/stack_timing.py 16384 tc 2>&1 | tail -21
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
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_it_$i.json 2> gpurun_out/r2_it_$i.err
  short gpurun_out/r2_it_$i.json "$envs"
done
