#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-600 > gpurun_out/test_gpu_n.log
cat gpurun_out/test_gpu_n.log
timeout 900 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['modes']); p=d['phases_ms_layer0']; print({k:v for k,v in p.items() if 'edge_bwd' in k})
r=d['roofline']; print(r['achieved'], r['frac'], r['model'], r['at_scale'])
PY
