#!/bin/bash
# Two-GPU pass (gpurun --gpus 2): smoke(), partitioned-path parity, the dp bench and the partitioned bench at 2 ranks.
set -u
mkdir -p gpurun_out
TAG=${1:-e}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_model.py::test_driver_smoke_entry_point -m gpu -q > gpurun_out/test_part_$TAG.log 2>&1
echo "partitioned tests rc=$?"; tail -4 gpurun_out/test_part_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2_$TAG.json 2> gpurun_out/bench_dp2_$TAG.err
echo "dp2 rc=$?"; cut -c1-300 gpurun_out/bench_dp2_$TAG.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --workload large --nodes 200000 --steps 5 --warmup 3 > gpurun_out/bench_part2_$TAG.json 2> gpurun_out/bench_part2_$TAG.err
echo "partitioned 2 rc=$?"; cut -c1-400 gpurun_out/bench_part2_$TAG.json
