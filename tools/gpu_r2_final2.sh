#!/bin/bash
# Final measurement pass (session 4 of round 2): GPU tests, smoke, default bench line (all blocks), reference arm, warm / cold launch
# lists, ncu --set full of the dominant kernel (Water-3D and 3.6 M edges) and of the new node-side forward kernels.
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300 > gpurun_out/test_gpu_final2.log; cat gpurun_out/test_gpu_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/smoke_final2.txt; cat gpurun_out/smoke_final2.txt
timeout 1500 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final2.json 2> gpurun_out/bench_reference_final2.err; echo "reference arm rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_final2.csv $B > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/launches_r2_final2_warm.csv $B > /dev/null 2>&1; echo "warm launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc4 -s 4 -c 1 -o gpurun_out/ncu_edge_bwd4_water3d_f2 -f $B > /dev/null 2>&1; echo "ncu water3d rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_h_fwd_tc -s 3 -c 1 -o gpurun_out/ncu_node_h_fwd_tc_f2 -f $B > /dev/null 2>&1; echo "ncu node_h rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:node_pre_fwd_tc3 -s 4 -c 1 -o gpurun_out/ncu_node_pre_fwd_tc3_f2 -f $B > /dev/null 2>&1; echo "ncu node_pre rc=$?"
FEGNN_EXP_MODES=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc4 -s 3 -c 1 -o gpurun_out/ncu_edge_bwd4_b20_f2 -f python tools/exp_edge_bwd.py water3d_b20 > /dev/null 2>&1; echo "ncu b20 rc=$?"
