#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-400
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], 'prep', d['phases_ms_layer0']['graph_prep'], 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
done
