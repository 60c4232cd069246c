#!/bin/bash
# first GPU pass of round 2: stack-kernel tests, the whole GPU suite, then A/B bench (per-layer launches vs stack kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stack.py -x -q -s -m gpu > gpurun_out/r2_stack_tests.log 2>&1; echo "stack tests rc=$?" | tee -a gpurun_out/r2_stack_tests.log
tail -30 gpurun_out/r2_stack_tests.log
timeout 900 python -m pytest tests -q -s -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
tail -15 gpurun_out/r2_gpu_tests.log
MSHGNN_STACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_bench_layer.json 2> gpurun_out/r2_bench_layer.err; echo "bench(per-layer) rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_bench_stack.json 2> gpurun_out/r2_bench_stack.err; echo "bench(stack) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_layer.json", "gpurun_out/r2_bench_stack.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "train ms", round(d["ms_per_step"], 4), "infer ms", round(d["inference"]["ms_per_step"], 4), "launches", d["gpu_launches"])
        print("   ", {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]})
    except Exception as e:
        print(f, "unreadable:", e)
PY
