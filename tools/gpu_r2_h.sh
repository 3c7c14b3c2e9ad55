#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vnegnn.py -m gpu -q 2>&1 | tail -25 | cut -c1-400
