#!/bin/bash
# sanity of the very last build: GPU tests, smoke, one short bench line
set -u
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -2 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager-bar --no-per-config --no-fp32-line --no-phases 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['clocks'])"
