#!/bin/bash
# 2-GPU pass: partitioned parity tests + the default N=2 bench line (replicas + partitioned config-5 block)
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2_au.json 2> gpurun_out/bench_n${N}_r2_au.err
echo "bench N=$N rc=$?"; tail -c 3000 gpurun_out/bench_n${N}_r2_au.json; tail -3 gpurun_out/bench_n${N}_r2_au.err | cut -c1-300
