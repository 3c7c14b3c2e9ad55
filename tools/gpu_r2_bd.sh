#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "pipelined" 2>&1 | tail -4 | cut -c1-600
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2 3; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'serial', round(d['e2e']['ms_per_step_serial_copy'],4))"
done
