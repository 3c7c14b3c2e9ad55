#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_h_forward" 2>&1 | tail -2 | cut -c1-300
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 600 $B --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step']); print({k:v for k,v in d['phases_ms_layer0'].items() if 'node' in k})"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --no-phases --steps 1 --warmup 3 --no-graph 2>&1 | grep "VTRACE node_h" | tail -3 > gpurun_out/vtrace_nodeh.txt
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --no-phases --workload large --steps 1 --warmup 3 --no-graph 2>&1 | grep "VTRACE node_h" | tail -2 > gpurun_out/vtrace_nodeh_large.txt
