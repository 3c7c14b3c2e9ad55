#!/bin/bash
# One GPU-box pass (under gpurun): the whole -m gpu suite, the default bench line (+ eager-PyTorch bar), other workloads,
# a launch list of two un-captured steps and full ncu captures of the edge kernels.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-a}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_gpu_$TAG.log 2>&1
echo "gpu suite rc=$?"; tail -12 gpurun_out/test_gpu_$TAG.log
timeout 400 python bench.py --gpu-eager-bar > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --workload water3d_b20 --no-cpu-baseline --steps 10 > gpurun_out/bench_b20_$TAG.json 2> gpurun_out/bench_b20_$TAG.err
echo "b20 rc=$?"
timeout 300 python bench.py --workload nbody100 --gpu-eager-bar --steps 10 > gpurun_out/bench_nbody100_$TAG.json 2> gpurun_out/bench_nbody100_$TAG.err
echo "nbody100 rc=$?"; cut -c1-300 gpurun_out/bench_nbody100_$TAG.json
BENCH="python bench.py --no-graph --no-cpu-baseline --no-phases --steps 2 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    $BENCH > gpurun_out/launches_$TAG.log 2>&1
echo "launch list rc=$?"
for k in edge_bwd_tc2_kernel edge_fwd_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f \
      -o gpurun_out/ncu_${k}_$TAG $BENCH > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
