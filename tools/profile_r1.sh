#!/bin/bash
# Run on the GPU box (under gpurun).  Launch list of two un-captured training steps + full ncu captures of the
# dominant kernels.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --no-graph --no-cpu-baseline --no-phases --steps 2 --warmup 3"
# warm-up (3 resident + 3 e2e steps) launches are skipped by counting: ~81 library launches + torch kernels per step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_all.csv $BENCH > gpurun_out/launches_bench.log 2>&1
for k in edge_bwd_tc2_kernel edge_fwd_tc_kernel virtual_bwd_heads_tc_kernel virtual_bwd_trunk_tc_kernel virtual_fwd_tc_kernel node_h_bwd1_kernel node_pre_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f \
      -o gpurun_out/ncu_$k $BENCH > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
