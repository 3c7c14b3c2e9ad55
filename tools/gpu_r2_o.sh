#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | cut -c1-600 > gpurun_out/test_gpu_o.log
cat gpurun_out/test_gpu_o.log
timeout 900 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_o.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_o.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['modes'])
r=d['roofline']; print(r['achieved'], r['frac'], r['model']['frac_of_model'], r['at_scale']); print(d['rollout'])
PY
