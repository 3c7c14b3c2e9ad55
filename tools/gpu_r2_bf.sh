#!/bin/bash
set -u
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
timeout 600 $B --workload large --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large default', d['ms_per_step'])"
FEGNN_DENSE_ROWS_MAX_TILES_PER_SM=1000000 timeout 600 $B --workload large --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large rows-kernel', d['ms_per_step'])"
FEGNN_DENSE_ROWS_MAX_TILES_PER_SM=1000000 timeout 600 $B --workload water3d_b20 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('b20 rows-kernel', d['ms_per_step'])"
timeout 600 $B --workload water3d_b20 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('b20 default', d['ms_per_step'])"
