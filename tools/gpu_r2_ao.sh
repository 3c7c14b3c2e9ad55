#!/bin/bash
# node_h forward on tcgen05 3xTF32: phase tests, full GPU suite, bench at Water-3D and at 1 M nodes
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_h_forward" 2>&1 | tail -5 | cut -c1-600
cat gpurun_out/node_h_fwd_mode3_*.txt
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | cut -c1-600
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step']); print({k:v for k,v in d['phases_ms_layer0'].items() if 'node' in k})"
FEGNN_MODE_NODE_FORWARD=0 timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step[node_forward=0]', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 600 $B --no-phases --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step'])"
FEGNN_MODE_NODE_FORWARD=0 timeout 600 $B --no-phases --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large[node_forward=0]', d['ms_per_step'])"
