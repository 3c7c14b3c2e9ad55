#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
for spec in "node_pre_fwd:2:1:node_pre_fwd" "node_h_z:1:1:node_h_z" "node_h_out:1:1:node_h_out"; do
  IFS=: read k s c o <<< "$spec"
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/ncu_${o}_r2ah -f $B > /dev/null 2>&1; echo "ncu $o rc=$?"
done
