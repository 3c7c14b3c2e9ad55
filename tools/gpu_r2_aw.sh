#!/bin/bash
set -u
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2 3; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('fused   step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
FEGNN_SORT_FUSED=0 timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('unfused step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
done
