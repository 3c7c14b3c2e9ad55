#!/bin/bash
# fresh warm launch list + bench of the build after the virtual-backward work (no kernel change against 6227bbb)
set -u
mkdir -p gpurun_out
B="python bench.py --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
  timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_an_warm.csv $B --steps 2 --warmup 3 --no-graph > /dev/null 2>&1; echo "warm launch list rc=$?"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --steps 1 --warmup 3 --no-graph > gpurun_out/vtrace5_all.txt 2>&1
