#!/bin/bash
# Last measurement pass of round 2: GPU tests, smoke, default bench line (all blocks), reference arm, warm launch list.
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300 > gpurun_out/test_gpu_final6.log; cat gpurun_out/test_gpu_final6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/smoke_final6.txt; cat gpurun_out/smoke_final6.txt
timeout 1500 python bench.py > gpurun_out/bench_final6.json 2> gpurun_out/bench_final6.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final6.json 2> gpurun_out/bench_reference_final6.err; echo "reference arm rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_final6_warm.csv $B > /dev/null 2>&1; echo "warm launch list rc=$?"
