#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "mse_mmd or mmd or pipelined or adam" 2>&1 | tail -3 | cut -c1-500
B="python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases"
for i in 1 2 3; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['gpu_launches'])"
done
timeout 600 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node 1 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases 2>/dev/null | tail -1 | cut -c1-200
