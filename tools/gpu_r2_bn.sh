#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_phases.py tests/test_gpu_fullsize.py tests/test_gpu_vnegnn.py -m gpu -q -x -k "node_h_forward or fullsize or full_size or vnegnn or VNEGNN" 2>&1 | tail -2 | cut -c1-300
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done
timeout 600 $B --workload large --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"node_h_fwd_tc" -c 6 --csv --log-file gpurun_out/launches_nodeh_large.csv $B --workload large --steps 1 --warmup 3 --no-graph > /dev/null 2>&1; grep node_h gpurun_out/launches_nodeh_large.csv | awk -F'","' '{print $NF}' | tail -3
