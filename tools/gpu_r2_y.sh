#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-500
