#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2_final6.json 2> gpurun_out/bench_n${N}_r2_final6.err
echo "bench N=$N rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_r2_final6.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'part', d['partitioned'].get('ms_per_step'), d['partitioned'].get('one_gpu',{}).get('ms_per_step'), d['partitioned'].get('parity_vs_one_gpu',{}).get('worst_rel_err'), d['partitioned'].get('error'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29586 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
