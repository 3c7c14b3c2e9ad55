#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py tests/test_gpu_fullsize.py -m gpu -q -x -k "node_pre_forward or fullsize or full_size" 2>&1 | tail -3 | cut -c1-400
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
FEGNN_NODE_PRE_TC3=0 timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step[node_pre fp32]', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
