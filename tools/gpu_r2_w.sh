#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "node_backward_tensor_core" 2>&1 | tail -2 | cut -c1-800
export FEGNN_MODE_NODE_BACKWARD=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"dense_bwd" -c 40 --csv --log-file gpurun_out/ncu_rows_times.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-phases --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_rows_times.csv | awk -F'","' '{print $5, $9, $NF}' | tail -4
timeout 600 python bench.py --no-per-config --no-cpu-baseline --no-gpu-eager-bar --no-fp32-line --no-phases > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step',)}, d['e2e']['ms_per_step'])
PY
