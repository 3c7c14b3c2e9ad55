#!/bin/bash
# One GPU-box pass (under gpurun): the whole -m gpu suite and the default bench line.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-a}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_gpu_$TAG.log 2>&1
echo "gpu suite rc=$?"; tail -6 gpurun_out/test_gpu_$TAG.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
