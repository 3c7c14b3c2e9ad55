#!/bin/bash
set -u
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | cut -c1-300
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2 3; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'], 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
done
