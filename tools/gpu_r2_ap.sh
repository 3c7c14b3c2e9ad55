#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_ap_warm.csv $B --steps 2 --warmup 3 --no-graph > /dev/null 2>&1; echo "warm launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"node_h|node_pre|dense_bwd" -c 200 --csv --log-file gpurun_out/launches_r2_ap_large.csv $B --workload large --steps 1 --warmup 1 --no-graph > /dev/null 2>&1; echo "large launch list rc=$?"
