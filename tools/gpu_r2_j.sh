#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_bwd|node_pre_bwd|node_h_bwd" -c 60 --csv --log-file gpurun_out/ncu_dense_times.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_dense_times.csv | awk -F'","' '{print $5, $9, $NF}' | tail -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_bwd_tc -s 20 -c 3 -o gpurun_out/ncu_dense_bwd -f \
  python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line > gpurun_out/ncu_dense.log 2>&1
echo "ncu rc=$?"
