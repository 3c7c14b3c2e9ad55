#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B > gpurun_out/vtrace3_all.txt 2>&1
grep -E "VTRACE" gpurun_out/vtrace3_all.txt | tail -40 > gpurun_out/vtrace3.txt
grep -v VTRACE gpurun_out/vtrace3_all.txt | grep -v DTRACE | tail -8 | cut -c1-300
