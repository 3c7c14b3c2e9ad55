#!/bin/bash
# clock64 stamps of one CTA of the per-tile node-side backward kernel (instrumented build, FEGNN_TRACE)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B 2>&1 | grep DTRACE | tail -40 > gpurun_out/dtrace.txt
wc -l gpurun_out/dtrace.txt
