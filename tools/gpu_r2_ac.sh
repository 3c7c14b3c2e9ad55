#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | cut -c1-400
B="python bench.py --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for i in 1 2; do
  timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --steps 1 --warmup 3 --no-graph 2>&1 | grep DTRACE | tail -12 > gpurun_out/dtrace2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_ac_warm.csv $B --steps 2 --warmup 3 --no-graph > /dev/null 2>&1; echo "warm launch list rc=$?"
