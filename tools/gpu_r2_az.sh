#!/bin/bash
set -u
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('normal', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
FEGNN_ROLLOUT_ORDER=rev timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('reversed', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches_rollout.csv $B2 > /dev/null 2>&1; echo "rc=$?"
