#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 600 $B --workload large --steps 3 --warmup 3 2>gpurun_out/large.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('large', d['ms_per_step'])"
timeout 600 $B --workload water3d_b20 --steps 5 --warmup 3 2>gpurun_out/b20.err | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('b20', d['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/launches_r2_ae_large.csv $B --workload large --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_large.log 2>&1; echo "large launch list rc=$?"
tail -3 gpurun_out/large.err gpurun_out/ncu_large.log
