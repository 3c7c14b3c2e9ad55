#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_partitioned.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2_bi.json 2> gpurun_out/bench_n${N}_r2_bi.err
echo "bench N=$N rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_r2_bi.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'part', d['partitioned'].get('ms_per_step'), d['partitioned'].get('parity_vs_one_gpu',{}).get('worst_rel_err'), d['partitioned'].get('error'))"
tail -2 gpurun_out/bench_n${N}_r2_bi.err | cut -c1-200
