#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "eval_forward or forward_only or smoke" 2>&1 | tail -5 | cut -c1-300
timeout 900 python bench.py --no-per-config > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','cuda_graph')}); print(d['e2e']['ms_per_step'], d['e2e']['ms_per_step_serial_copy']); print(d.get('rollout')); print(d.get('fp32_mode',{}).get('ms_per_step'))
PY
tail -3 gpurun_out/bench_g.err
