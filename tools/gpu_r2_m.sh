#!/bin/bash
set -u
mkdir -p gpurun_out
FEGNN_EXP_MODES=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc4 -s 3 -c 1 -o gpurun_out/ncu_bwd4_b20 -f python tools/exp_edge_bwd.py water3d_b20 > gpurun_out/ncu_bwd4.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bwd4.log
