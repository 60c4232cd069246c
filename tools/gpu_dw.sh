set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 4 > gpurun_out/launches.log 2>&1
python tools/launch_traffic.py gpurun_out/launches.csv gpurun_out/step_traffic.json | head -6
python bench.py --skip-cpu --skip-e2e | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), {k['kernel']: round(k['ms_per_step'],4) for k in d['kernels'][:8]})
"
