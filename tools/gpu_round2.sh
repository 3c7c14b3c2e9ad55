#!/bin/bash
# Two-GPU pass (gpurun --gpus 2): partitioned-path parity with both halo transports, partitioned bench p2p vs nccl.
set -u
mkdir -p gpurun_out
TAG=${1:-e}
timeout 900 python -m pytest tests/test_gpu_partitioned.py -m gpu -q > gpurun_out/test_part_$TAG.log 2>&1
echo "partitioned tests rc=$?"; tail -15 gpurun_out/test_part_$TAG.log | cut -c1-400
for halo in p2p nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
      bench.py --gpus 2 --workload large --nodes 200000 --steps 5 --warmup 3 --halo $halo > gpurun_out/bench_part2_${halo}_$TAG.json 2> gpurun_out/bench_part2_${halo}_$TAG.err
  echo "partitioned 2 $halo rc=$?"; tail -1 gpurun_out/bench_part2_${halo}_$TAG.json | cut -c1-300
done
