#!/bin/bash
# Compile check of the work-in-progress kernels (not part of libfegnn.so): prints ptxas resource usage.
set -e
cd "$(dirname "$0")"
cat > /tmp/fegnn_wip_check.cu <<'EOT'
#include "common.cuh"
#include "edge_kernels.cu"
#include "edge_tc.cu"
#include "edge_tc_bwd2.cu"
#include "wip/edge_tc_bwd3.cu"
EOT
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I.. -Xptxas -v -c /tmp/fegnn_wip_check.cu -o /tmp/fegnn_wip_check.o 2>&1 | grep -A2 "edge_bwd_tc3" | head -8
echo "compile check ok"
