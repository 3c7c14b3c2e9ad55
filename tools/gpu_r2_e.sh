#!/bin/bash
# 2-GPU pass: partitioned parity (NCCL / peer-memory / fused transports, host and device plans), default bench at N=2.
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m pytest tests/test_gpu_partitioned.py -m gpu -q > gpurun_out/test_part_r2.log 2>&1
echo "partitioned tests rc=$?"; tail -30 gpurun_out/test_part_r2.log | cut -c1-600
cat gpurun_out/dist_check_fused*rank0.txt 2>/dev/null | head -40
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2.json 2> gpurun_out/bench_n${N}_r2.err
echo "bench N=$N rc=$?"; tail -c 4500 gpurun_out/bench_n${N}_r2.json; tail -5 gpurun_out/bench_n${N}_r2.err | cut -c1-300
