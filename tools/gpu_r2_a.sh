#!/bin/bash
# Round-2 GPU pass A (1 GPU): layout probe, mode-5 phase tests under a timeout, the full GPU suite, bench lines, launch list.
set -u
mkdir -p gpurun_out
TAG=${1:-a}
nvidia-smi -L; free -g | head -2; nproc
timeout 120 tools/umma_probe_f16 > gpurun_out/umma_probe_f16_r2.txt 2>&1; echo "probe_f16 rc=$?"; grep -c -i "exact\|ok" gpurun_out/umma_probe_f16_r2.txt; grep "F8" gpurun_out/umma_probe_f16_r2.txt | cut -c1-200
timeout 420 python -m pytest tests/test_gpu_phases.py -m gpu -q -k "test_edge_backward_modes and -5" > gpurun_out/test_mode5_$TAG.log 2>&1
echo "mode5 tests rc=$?"; tail -5 gpurun_out/test_mode5_$TAG.log | cut -c1-600
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/test_gpu_$TAG.log 2>&1
echo "gpu tests rc=$?"; tail -25 gpurun_out/test_gpu_$TAG.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
FEGNN_MODE_EDGE_BACKWARD=5 timeout 300 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line > gpurun_out/bench_mode5_$TAG.json 2> gpurun_out/bench_mode5_$TAG.err
echo "bench mode5 rc=$?"; cut -c1-400 gpurun_out/bench_mode5_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line > gpurun_out/launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
