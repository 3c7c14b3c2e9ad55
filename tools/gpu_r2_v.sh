#!/bin/bash
set -u
mkdir -p gpurun_out
export FEGNN_MODE_NODE_BACKWARD=1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:dense_bwd_tc_rows -s 22 -c 2 -o gpurun_out/ncu_rows -f \
  python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line > gpurun_out/ncu_rows.log 2>&1
echo "ncu rc=$?"
