#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-600 > gpurun_out/test_gpu_k.log
cat gpurun_out/test_gpu_k.log
timeout 600 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"
for nb in 0 1; do
FEGNN_MODE_NODE_BACKWARD=$nb timeout 600 python bench.py --workload water3d_b20 --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line > gpurun_out/bench_k_b20_nb$nb.json 2> gpurun_out/bench_k_b20.err; echo "bench b20 rc=$?"
done
python - <<'PY'
import json
for f in ('bench_k','bench_k_b20_nb0','bench_k_b20_nb1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,{k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['modes']); p=d['phases_ms_layer0']; print({k:p[k] for k in ('node_h_bwd','node_pre_bwd','edge_bwd','virtual_bwd','node_h_fwd','node_pre_fwd','edge_fwd','virtual_fwd')})
    except Exception as e: print(f, 'ERR', e)
PY
