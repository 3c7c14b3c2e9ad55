#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullsize.txt
for pdl in 0 1 1 0 1; do
echo "=== PDL=$pdl"
FEGNN_PDL=$pdl timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "water3d-1.0" 2>&1 | grep -E "passed|failed|over|AssertionError: \(\[" | cut -c1-700
done
cut -c1-330 gpurun_out/parity_fullsize.txt
