#!/bin/bash
# round-2 evidence pass on one box: GPU suite, smoke, default bench line + reference arm, launch lists with DRAM counters of a
# train step and of an inference forward, full ncu capture of the stack launches (forward + backward) and of the encoder kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -s -m gpu > gpurun_out/r2e_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/r2e_gpu_tests.log | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 700 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2e_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2e_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 36 -c 40 --csv --log-file gpurun_out/r2e_launches.csv python tools/profile_step.py 4 > gpurun_out/r2e_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/r2e_launches_infer.csv python tools/profile_infer.py 5 > gpurun_out/r2e_launches_infer.log 2>&1; echo "infer launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_stack2 -s 2 -c 2 -f -o gpurun_out/r2e_stack_prof python tools/profile_step.py 2 > gpurun_out/r2e_stack_prof.log 2>&1; echo "ncu stack rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_encoder|k_tc_reducegemm' -s 3 -c 3 -f -o gpurun_out/r2e_enc_prof python tools/profile_step.py 2 > gpurun_out/r2e_enc_prof.log 2>&1; echo "ncu enc rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print("train ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "infer ms", round(d["inference"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print({k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]})
print("strong", json.dumps(d["strong"])[:700])
PY
