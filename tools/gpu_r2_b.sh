#!/bin/bash
set -u
mkdir -p gpurun_out
for w in water3d water3d_b20; do
for x in 0 1 2 3 4 7; do
  FEGNN_EXP=$x timeout 200 python tools/exp_edge_bwd.py $w 2>&1 | tail -1
done; done | tee gpurun_out/exp_edge_bwd.txt
timeout 900 python -m pytest tests/test_gpu_callers.py -m gpu -q 2>&1 | tail -15 | cut -c1-500
cat gpurun_out/callers_train_single_epoch.txt
