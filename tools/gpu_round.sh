#!/bin/bash
# One GPU-box pass (under gpurun): the whole -m gpu suite, the default bench line (+ eager-PyTorch bar), other workloads,
# a launch list of two un-captured steps.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-a}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_gpu_$TAG.log 2>&1
echo "gpu suite rc=$?"; tail -6 gpurun_out/test_gpu_$TAG.log
timeout 400 python bench.py --gpu-eager-bar > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --workload water3d_b20 --no-cpu-baseline --steps 10 > gpurun_out/bench_b20_$TAG.json 2> gpurun_out/bench_b20_$TAG.err
echo "b20 rc=$?"
timeout 300 python bench.py --workload nbody100 --no-cpu-baseline --steps 10 > gpurun_out/bench_nbody100_$TAG.json 2> gpurun_out/bench_nbody100_$TAG.err
echo "nbody100 rc=$?"
FEGNN_PRECISION=tf32_all timeout 300 python bench.py --no-cpu-baseline --no-phases > gpurun_out/bench_tf32all_$TAG.json 2> gpurun_out/bench_tf32all_$TAG.err
echo "tf32_all rc=$?"
timeout 600 python bench.py --workload large --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_large1m_$TAG.json 2> gpurun_out/bench_large1m_$TAG.err
echo "large 1M rc=$?"; cut -c1-400 gpurun_out/bench_large1m_$TAG.json
BENCH="python bench.py --no-graph --no-cpu-baseline --no-phases --steps 2 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    $BENCH > gpurun_out/launches_$TAG.log 2>&1
echo "launch list rc=$?"
