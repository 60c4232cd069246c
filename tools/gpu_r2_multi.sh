#!/bin/bash
# N-GPU pass: NCCL multi-rank correctness test, the bench line under torchrun (weak value + strong block), and the
# per-kernel breakdown of one rank's share of a strong-scaling step.  usage: tools/gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -s -m gpu 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-extra > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2_bench_n{n}.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "train ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "e2e_windowed", d["e2e_windowed"] and round(d["e2e_windowed"]["value"]))
print("strong", json.dumps(d["strong"]))
print("allreduce", d["allreduce"])
PY
for b in 2048 4096; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_bench_b$b.json 2> gpurun_out/r2_bench_b$b.err
  python - $b <<'PY'
import json, sys
b = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2_bench_b{b}.json").read().strip().splitlines()[-1])
print("B", b, "train ms", round(d["ms_per_step"], 4), "profiled", round(d["profiled_ms_per_step"], 4), "infer ms", round(d["inference"]["ms_per_step"], 4))
print("   ", {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]})
PY
done
