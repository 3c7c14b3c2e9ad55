#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_bwd_tc3 -s 3 -c 1 -o gpurun_out/ncu_bwd3_water3d -f python tools/exp_edge_bwd.py water3d > gpurun_out/ncu_bwd3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bwd3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"edge_bwd|stats" -c 40 --csv --log-file gpurun_out/ncu_bwd_times.csv python tools/exp_edge_bwd.py water3d > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_bwd_times.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
