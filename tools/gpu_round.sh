#!/bin/bash
# One GPU-box pass (under gpurun): smoke(), the whole -m gpu suite, the default bench line, protein / FastRF workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-a}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_gpu_$TAG.log 2>&1
echo "gpu suite rc=$?"; tail -6 gpurun_out/test_gpu_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --workload protein --gpu-eager-bar --steps 10 > gpurun_out/bench_protein_$TAG.json 2> gpurun_out/bench_protein_$TAG.err
echo "protein rc=$?"; cut -c1-300 gpurun_out/bench_protein_$TAG.json
timeout 300 python bench.py --workload protein --model fastrf --gpu-eager-bar --steps 10 > gpurun_out/bench_protein_rf_$TAG.json 2> gpurun_out/bench_protein_rf_$TAG.err
echo "protein fastrf rc=$?"; cut -c1-300 gpurun_out/bench_protein_rf_$TAG.json
timeout 300 python bench.py --model fastrf --gpu-eager-bar > gpurun_out/bench_water3d_rf_$TAG.json 2> gpurun_out/bench_water3d_rf_$TAG.err
echo "water3d fastrf rc=$?"; cut -c1-300 gpurun_out/bench_water3d_rf_$TAG.json
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "reference arm rc=$?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
