#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-400
B="python bench.py --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
  timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
timeout 600 $B --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step'])"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --steps 1 --warmup 3 --no-graph > gpurun_out/vtrace4_all.txt 2>&1
