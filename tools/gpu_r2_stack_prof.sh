#!/bin/bash
# stack kernel: chunk-size sweep (bench, device events) + one ncu --set full capture of the forward and backward stack launches
mkdir -p gpurun_out
short() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"], 4) for k in d["kernels"]}
    print(sys.argv[2], "train", round(d["ms_per_step"], 4), "infer", round(d["inference"]["ms_per_step"], 4),
          {k: ks.get(k) for k in ("stack_fwd", "stack_bwd", "dw_layers", "conv_fwd", "dx_bwd", "base_mlp_fwd", "base_mlp_bwd")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
for rc in 6 12 24 48 100000; do
  MSHGNN_STACK_RC=$rc timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_rc$rc.json 2> gpurun_out/r2_rc$rc.err
  short gpurun_out/r2_rc$rc.json "RC=$rc"
done
MSHGNN_STACK=0 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-extra > gpurun_out/r2_rc_off.json 2> gpurun_out/r2_rc_off.err
short gpurun_out/r2_rc_off.json "per-layer"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_stack -s 2 -c 2 -f -o gpurun_out/r2_stack_prof python tools/profile_step.py 2 > gpurun_out/r2_stack_prof.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_stack_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_stack_launches.csv python tools/profile_step.py 3 > gpurun_out/r2_stack_launches.log 2>&1
echo "ncu launches rc=$?"
