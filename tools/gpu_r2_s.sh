#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'])
p=d['phases_ms_layer0']; print({k:v for k,v in p.items() if '[' not in k})
r=d['roofline']; print(r['achieved'], r['frac'], r['model']['frac_of_model'], r['at_scale'])
PY
