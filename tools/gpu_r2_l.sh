#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "edge_backward_modes and (7 or 8)" 2>&1 | tail -5 | cut -c1-900
for m in 7; do for n in c3 small_graphs c3_gravity_heavy; do echo "== mode $m $n"; cat gpurun_out/edge_bwd_mode${m}_${n}_l0.txt 2>/dev/null | sort -k3 -g -r | head -4; done; done
for w in water3d water3d_b20; do FEGNN_EXP_MODES=4,7,8 timeout 300 python tools/exp_edge_bwd.py $w 2>&1 | tail -1; done
for x in 3 7; do FEGNN_EXP=$x FEGNN_EXP_MODES=7 timeout 300 python tools/exp_edge_bwd.py water3d_b20 2>&1 | tail -1; done
