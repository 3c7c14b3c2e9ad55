#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py -m gpu -q -k "node_backward_tensor_core" 2>&1 | tail -15 | cut -c1-600
cat gpurun_out/node_bwd_tc_c3_gravity_heavy_l0.txt 2>/dev/null | head -40
timeout 300 python tools/debug_vn.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_vnegnn.py -m gpu -q 2>&1 | tail -8 | cut -c1-400
timeout 900 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step']); p=d['phases_ms_layer0']; print({k:p[k] for k in ('node_h_bwd','node_pre_bwd','edge_bwd','virtual_bwd','node_h_fwd','node_pre_fwd')})
PY
