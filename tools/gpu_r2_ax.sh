#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_radius_graph.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-300
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2 3; do
timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('fused   step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
FEGNN_SORT_FUSED=0 timeout 600 $B --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('unfused step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
done
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('prep', d['phases_ms_layer0']['graph_prep'])"
