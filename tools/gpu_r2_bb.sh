#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_radius_graph.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"radix|count_rows|scan2" -c 40 --csv --log-file gpurun_out/launches_sort.csv python tools/rollout_probe.py > /dev/null 2>&1; echo "rc=$?"
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), 'rollout', d['rollout']['forward_only']['ms'], d['rollout']['training_forward']['ms'])"
done
