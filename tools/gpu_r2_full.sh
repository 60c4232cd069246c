#!/bin/bash
# whole GPU suite + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -s -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gpu_tests.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
print("train ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "infer ms", round(d["inference"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print({k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]})
print("strong", json.dumps(d["strong"])[:900])
print("cpu", d["cpu_baseline"]); print("other", json.dumps(d["other_configs"])[:600])
PY
