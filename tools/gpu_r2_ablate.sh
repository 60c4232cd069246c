#!/bin/bash
# ablation of the CTA-pair stack kernel (MSHGNN_STACK_DEBUG bit mask: 1 no A loads, 2 no W loads, 4 no MMAs, 8 bare epilogue); timing only
mkdir -p gpurun_out
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4),
          {k: ks.get(k) for k in ("stack_fwd", "stack_bwd")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
for arg in "$@"; do
  dbg=${arg%%:*}; la=1; case "$arg" in *:*) la=${arg##*:};; esac
  MSHGNN_STACK_LOOKAHEAD=$la MSHGNN_STACK_DEBUG=$dbg timeout 200 python bench.py --steps 6 --warmup 3 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2d_abl_$dbg.json 2> gpurun_out/r2d_abl_$dbg.err
  short gpurun_out/r2d_abl_$dbg.json "debug=$dbg lookahead=$la"
done
