#!/bin/bash
# ncu --set full of the non-edge-backward kernels of the Water-3D step (node-side backward per-tile kernel, virtual backward
# heads / trunk, edge forward, virtual forward), warm caches (the step runs back to back in the product).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for spec in "dense_bwd_tc_rows:10:3:dense_rows" "virtual_bwd_heads:4:1:vbwd_heads" "virtual_bwd_trunk:4:1:vbwd_trunk" "edge_fwd_tc:4:1:edge_fwd" "virtual_fwd_tc:4:1:vfwd"; do
  IFS=: read k s c o <<< "$spec"
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/ncu_${o}_r2z -f $B > /dev/null 2>&1; echo "ncu $o rc=$?"
done
