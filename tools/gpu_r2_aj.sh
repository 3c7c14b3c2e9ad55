#!/bin/bash
# clock64 stamps between the barriers of one CTA of the virtual backward kernels (instrumented build)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line"
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B 2>&1 | grep -E "VTRACE|DTRACE" | tail -16 > gpurun_out/vtrace.txt
wc -l gpurun_out/vtrace.txt
FEGNN_LIB=$PWD/fastegnn_b200/_C/libfegnn_trace.so timeout 600 $B --workload large 2>&1 | grep -E "VTRACE" | tail -4 > gpurun_out/vtrace_large.txt
