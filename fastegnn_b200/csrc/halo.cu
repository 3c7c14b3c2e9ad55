// halo.cu -- halo exchange of the spatially partitioned path by direct peer-memory access over NVLink / NVSwitch.
//
// The partitioned layer (fastegnn_b200/partitioned.py; the multi-GPU form of models/FastEGNN.py:192-223) needs, per
// layer, the owners' (Q_j, x_j) rows of every remote neighbour j (forward) and the sum of the users' (dQ_j, dx_j) at
// the owner (backward).  Instead of pack -> NCCL all-to-all -> unpack, the rows travel in ONE kernel each way:
//   push        warp per sent row: read Q[i] (256 B) and x[i] (12 B) locally, store them straight into the halo rows of
//               the destination rank's Q / x arrays (peer pointers from torch symmetric memory);
//   reduce_push warp per halo row: add this rank's dQ / dx halo rows into the owner's rows with remote atomics.
// The destination addresses are fixed by the partition plan, so they are precomputed once per plan as 64-bit tables.
// Ordering across ranks is the caller's (a symmetric-memory barrier on the same stream after the kernel).
#include "common.cuh"

namespace fegnn {

__global__ void __launch_bounds__(256) halo_push_kernel(int n, const int* __restrict__ src_row,
                                                        const unsigned long long* __restrict__ dst_q,
                                                        const unsigned long long* __restrict__ dst_x,
                                                        const float* __restrict__ Q, const float* __restrict__ x) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int i = src_row[w];
  const float2 q = *reinterpret_cast<const float2*>(Q + (size_t)i * kH + 2 * lane);
  *reinterpret_cast<float2*>(reinterpret_cast<float*>(dst_q[w]) + 2 * lane) = q;
  if (lane < 3) reinterpret_cast<float*>(dst_x[w])[lane] = x[(size_t)i * 3 + lane];
}

__global__ void __launch_bounds__(256) halo_reduce_push_kernel(int n, int row0,
                                                               const unsigned long long* __restrict__ dst_q,
                                                               const unsigned long long* __restrict__ dst_x,
                                                               const float* __restrict__ gQ, const float* __restrict__ gx) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const size_t i = (size_t)row0 + w;
  float* dq = reinterpret_cast<float*>(dst_q[w]);
  atomicAdd(dq + lane, gQ[i * kH + lane]);
  atomicAdd(dq + 32 + lane, gQ[i * kH + 32 + lane]);
  if (lane < 3) atomicAdd(reinterpret_cast<float*>(dst_x[w]) + lane, gx[i * 3 + lane]);
}

cudaError_t launch_halo_push(int n, const int* src_row, const unsigned long long* dst_q, const unsigned long long* dst_x,
                             const float* Q, const float* x, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  halo_push_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, src_row, dst_q, dst_x, Q, x); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_halo_reduce_push(int n, int row0, const unsigned long long* dst_q, const unsigned long long* dst_x,
                                    const float* gQ, const float* gx, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  halo_reduce_push_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, row0, dst_q, dst_x, gQ, gx); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
