#!/bin/bash
# Final measurement pass of round 2: GPU tests, default bench line (all blocks), reference arm, launch lists (cold = default ncu,
# warm = --cache-control none), ncu --set full of the dominant kernel at Water-3D and at 3.6 M edges.
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300 > gpurun_out/test_gpu_final.log; cat gpurun_out/test_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/smoke_final.txt; cat gpurun_out/smoke_final.txt
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; echo "reference arm rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_final.csv $B > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_final_warm.csv $B > /dev/null 2>&1; echo "warm launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc4 -s 4 -c 1 -o gpurun_out/ncu_edge_bwd4_water3d -f $B > /dev/null 2>&1; echo "ncu water3d rc=$?"
FEGNN_EXP_MODES=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc4 -s 3 -c 1 -o gpurun_out/ncu_edge_bwd4_b20 -f python tools/exp_edge_bwd.py water3d_b20 > /dev/null 2>&1; echo "ncu b20 rc=$?"
