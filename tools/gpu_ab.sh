#!/bin/bash
# One GPU-box visit for an A/B of two switches: MSHGNN_ENCODER (v1 = one CTA per SM, 64-column K blocks; default = two CTAs
# per SM) and MSHGNN_DW_FIT (0 = one split count for every weight-gradient launch; default = wave-fitted per layer).
# Every step runs under its own timeout; a failing smoke() of the default build switches the rest of the visit to v1.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
SMOKE=$?
echo "smoke rc=$SMOKE"; tail -4 gpurun_out/smoke.log
if [ $SMOKE -ne 0 ]; then
  export MSHGNN_ENCODER=v1
  timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v1.log 2>&1
  echo "smoke(v1) rc=$?"; tail -3 gpurun_out/smoke_v1.log
fi
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
short() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], {k: d[k] for k in ('value', 'ms_per_step')}, 'inf', d['inference']['ms_per_step'])
for k in d['kernels'][:8]: print(f"  {k['kernel']:18s} {k['ms_per_step']:.4f} n={k['launches_per_step']}", {x: round(k[x], 3) for x in k if x.startswith('frac')})
PY
}
timeout 400 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; short gpurun_out/bench_new.json
MSHGNN_ENCODER=v1 timeout 400 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/bench_encv1.json 2> gpurun_out/bench_encv1.err; short gpurun_out/bench_encv1.json
MSHGNN_DW_FIT=0 timeout 400 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/bench_nofit.json 2> gpurun_out/bench_nofit.err; short gpurun_out/bench_nofit.json
echo "t=$(( $(date +%s) - T0 ))s"
