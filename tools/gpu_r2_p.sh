#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-600 > gpurun_out/test_gpu_p.log
cat gpurun_out/test_gpu_p.log
for pdl in 0 1; do
FEGNN_PDL=$pdl timeout 900 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line --no-phases > gpurun_out/bench_p$pdl.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_p.err | cut -c1-300
done
python - <<'PY'
import json
for f in ('bench_p0','bench_p1'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,{k:d[k] for k in ('value','ms_per_step','gpu_launches','cuda_graph')}, d['e2e']['ms_per_step'], d['rollout']['forward_only'])
PY
