set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e',d['e2e']['value'],'win',{k:v for k,v in d['e2e_windowed'].items() if k!='note'}); print('inf',d['inference'])
for k in d['kernels']: print(f"{k['kernel']:18s} {k['ms_per_step']:.4f} n={k['launches_per_step']}", {x:round(k[x],3) for x in k if x.startswith('frac')})
PY
