#!/bin/bash
# N-GPU pass of the final build: the default N>1 bench line (replicas + partitioned config-5 block)
set -u
mkdir -p gpurun_out
N=${1:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2_final4.json 2> gpurun_out/bench_n${N}_r2_final4.err
echo "bench N=$N rc=$?"; tail -c 2500 gpurun_out/bench_n${N}_r2_final4.json; tail -2 gpurun_out/bench_n${N}_r2_final4.err | cut -c1-300
