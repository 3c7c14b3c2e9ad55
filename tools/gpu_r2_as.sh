#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_pre_forward" 2>&1 | tail -4 | cut -c1-400
cat gpurun_out/node_pre_fwd_mode3_c3_gravity_heavy_l0.txt gpurun_out/node_pre_fwd_mode3_c8_l1.txt
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-400
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
FEGNN_NODE_PRE_TC3=0 timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step[node_pre fp32]', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 600 $B --no-phases --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step'])"
