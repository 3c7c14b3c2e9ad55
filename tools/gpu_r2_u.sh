#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_backward_tensor_core" 2>&1 | tail -5 | cut -c1-800
cat gpurun_out/node_bwd_tc_c3_gravity_heavy_l0.txt 2>/dev/null | sort -k3 -g -r | head -6
for nb in 0 1; do
FEGNN_MODE_NODE_BACKWARD=$nb timeout 600 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line > gpurun_out/bench_u$nb.json 2> gpurun_out/bench_u.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for f in ('bench_u0','bench_u1'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    p=d['phases_ms_layer0']
    print(f,{k:d[k] for k in ('ms_per_step',)}, d['e2e']['ms_per_step'], {k:p[k] for k in ('node_h_bwd','node_pre_bwd')})
PY
