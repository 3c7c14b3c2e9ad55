#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
echo "bench N=$N rc=$?"; tail -c 3800 gpurun_out/bench_n${N}_r2.json; tail -3 gpurun_out/bench_n${N}_r2.err | cut -c1-300
