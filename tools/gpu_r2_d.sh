#!/bin/bash
set -u
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullsize.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/test_gpu_d.log 2>&1
echo "gpu tests rc=$?"; tail -12 gpurun_out/test_gpu_d.log | cut -c1-300
cat gpurun_out/parity_fullsize.txt
for w in water3d water3d_b20; do FEGNN_EXP=0 timeout 200 python tools/exp_edge_bwd.py $w 2>&1 | tail -1; done
timeout 600 python bench.py --workload water3d_b20 --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line --steps 10 > gpurun_out/bench_b20_d.json 2> gpurun_out/bench_b20_d.err; cut -c1-330 gpurun_out/bench_b20_d.json
