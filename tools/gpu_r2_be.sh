#!/bin/bash
set -u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_scatter_fused -s 2 -c 2 -o gpurun_out/ncu_scatter_fused -f python tools/rollout_probe.py > /dev/null 2>&1; echo "rc=$?"
