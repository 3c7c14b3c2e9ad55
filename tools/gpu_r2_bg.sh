#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_phases.py tests/test_gpu_fullsize.py -m gpu -q -x -k "node_backward or fullsize or full_size or phases" 2>&1 | tail -2 | cut -c1-300
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_r2_bg_large.csv $B --workload large --steps 1 --warmup 3 --no-graph > gpurun_out/bg_large.log 2>&1; echo "large launch list rc=$?"; tail -2 gpurun_out/bg_large.log | cut -c1-300
