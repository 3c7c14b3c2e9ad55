#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / tensor-memory kernels (phase tests of the edge and virtual phases)
set -u
mkdir -p gpurun_out
SEL='(edge_backward_modes and (7 or 5 or 4) and c3 and not gravity) or (edge_forward_modes and c3 and not gravity) or (virtual_backward_modes and c3 and not gravity and 1-) or (virtual_forward_modes and c3 and not gravity)'
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_phases.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
