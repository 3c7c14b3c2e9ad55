#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_backward_tensor_core" 2>&1 | tail -2 | cut -c1-800
for nb in 0 1 0 1; do
FEGNN_MODE_NODE_BACKWARD=$nb timeout 600 python bench.py --steps 40 --warmup 10 --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line --no-phases > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
print('node_backward=$nb', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks'])
PY
done
