#!/bin/bash
set -u
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches_rollout.csv python tools/rollout_probe.py > gpurun_out/rollout_probe.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/rollout_probe.log
