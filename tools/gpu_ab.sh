#!/bin/bash
# One GPU-box visit: smoke, the GPU parity suite, then bench A/B runs.  usage: tools/gpu_ab.sh ["ENV=val ENV2=val" ...]
# (each argument is an environment assignment list for one extra short bench run next to the default one)
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
short() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], {k: d[k] for k in ('value', 'ms_per_step')}, 'inf', d['inference']['ms_per_step'])
for k in d['kernels']: print(f"  {k['kernel']:18s} {k['ms_per_step']:.4f} n={k['launches_per_step']}", {x: round(k[x], 3) for x in k if x.startswith('frac')})
PY
}
timeout 400 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; short gpurun_out/bench_new.json
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 400 python bench.py --skip-cpu --skip-e2e --skip-extra > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  echo "--- $envs"; short gpurun_out/bench_ab$i.json | head -8
done
echo "t=$(( $(date +%s) - T0 ))s"
