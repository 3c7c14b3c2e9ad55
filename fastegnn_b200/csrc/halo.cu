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


// =====================================================================================================================
// Second generation: exchange + inter-rank ordering in ONE kernel, no NCCL and no separate barrier launch.
//
// Every rank owns a signal pad (uint32 [channels][16], symmetric memory) that its peers write with st.release.sys and
// it reads with ld.acquire.sys; `epoch[channel]` is a LOCAL device counter (so a captured CUDA graph replays correctly)
// that every rank advances identically because all ranks run the same sequence of exchanges.  A kernel
//   1. writes its payload into the peers' memory (plain stores over NVLink),
//   2. last CTA done: __threadfence_system(), then signals epoch+1 to EVERY rank and waits until every rank's signal has
//      arrived in its own pad,
// so that when the kernel completes on a rank, all payload destined for that rank is in place: stream order does the
// rest.  Signalling everybody makes each exchange a full barrier, which is what makes buffer reuse safe with two
// alternating receive buffers (a rank cannot be two exchanges ahead of any other rank).  A spin is bounded (~4 s of
// %globaltimer); on timeout the sticky error word is set and the kernel returns instead of hanging the GPU.
//
//   halo_push_signal     forward: owners' (Q_j, x_j) rows -> users' halo rows.  backward: users' (dQ_j, dx_j) halo rows ->
//                        the owner's RECEIVE buffer (one slot per (user, row)), plain stores: no remote atomics.
//   halo_reduce_apply    backward, local: every owned boundary row adds its received slots in FIXED (rank, row) order
//                        -> the reverse halo is deterministic run to run.
//   p2p_allreduce        one-shot all-reduce of <= ar_capacity floats (the per-graph sums, <= ~2 KB): every rank stores
//                        its vector into slot [rank] of every peer, signal / wait, then sums the slots in rank order --
//                        bitwise identical on all ranks, which keeps the replicated per-graph state identical.
constexpr int kP2PMaxWorld = 16;
constexpr int kP2PChannels = 4;

struct P2PArgs {
  unsigned long long sig_peer[kP2PMaxWorld];
  unsigned long long ar_peer[kP2PMaxWorld];
  unsigned* epoch;
  unsigned* done;
  int* err;
  int rank, world, ar_capacity;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// All threads of ONE CTA (blockDim.x >= world) call this after the CTA's (and, through the done counter, the grid's)
// payload stores.  Returns after every rank's signal for this exchange has arrived.
__device__ __forceinline__ void p2p_signal_wait(const P2PArgs& s, int channel) {
  const unsigned e = s.epoch[channel] + 1u;
  __threadfence_system();
  __syncthreads();
  const int p = threadIdx.x;
  if (p < s.world) {
    unsigned* mine_at_peer = reinterpret_cast<unsigned*>(s.sig_peer[p]) + channel * kP2PMaxWorld + s.rank;
    st_release_sys(mine_at_peer, e);
    const unsigned* his_at_me = reinterpret_cast<const unsigned*>(s.sig_peer[s.rank]) + channel * kP2PMaxWorld + p;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(his_at_me) - e) < 0) {
      if (globaltimer_ns() - t0 > 4000000000ull) {
        atomicExch(s.err, 1 + channel);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s.epoch[channel] = e;
}

// src_row == nullptr: rows row0 .. row0 + n - 1 (the halo rows of a gradient array); else rows src_row[k]
__global__ void __launch_bounds__(256) halo_push_signal_kernel(P2PArgs s, int channel, int n, const int* __restrict__ src_row,
                                                               int row0, const unsigned long long* __restrict__ dst_q,
                                                               const unsigned long long* __restrict__ dst_x,
                                                               const float* __restrict__ Q, const float* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += (gridDim.x * blockDim.x) >> 5) {
    const size_t i = src_row != nullptr ? (size_t)src_row[w] : (size_t)row0 + w;
    const float2 q = *reinterpret_cast<const float2*>(Q + i * kH + 2 * lane);
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(dst_q[w]) + 2 * lane) = q;
    if (lane < 3) reinterpret_cast<float*>(dst_x[w])[lane] = x[i * 3 + lane];
  }
  __shared__ int is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&s.done[channel], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) s.done[channel] = 0;
  p2p_signal_wait(s, channel);
}

// warp per owned boundary row r = rows[k]: gQ[r] += sum over slots[ptr[k] .. ptr[k+1]) of rq[slot], same for gx
__global__ void __launch_bounds__(256) halo_reduce_apply_kernel(int n_rows, const int* __restrict__ rows,
                                                                const int* __restrict__ ptr, const int* __restrict__ slots,
                                                                const float* __restrict__ rq, const float* __restrict__ rx,
                                                                float* __restrict__ gQ, float* __restrict__ gx) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= n_rows) return;
  const size_t r = rows[k];
  float2 acc = *reinterpret_cast<const float2*>(gQ + r * kH + 2 * lane);
  float ax = lane < 3 ? gx[r * 3 + lane] : 0.f;
  for (int j = ptr[k]; j < ptr[k + 1]; ++j) {
    const size_t sl = slots[j];
    const float2 v = *reinterpret_cast<const float2*>(rq + sl * kH + 2 * lane);
    acc.x += v.x; acc.y += v.y;
    if (lane < 3) ax += rx[sl * 3 + lane];
  }
  *reinterpret_cast<float2*>(gQ + r * kH + 2 * lane) = acc;
  if (lane < 3) gx[r * 3 + lane] = ax;
}

struct P2PSegs {
  float* ptr[4];
  int count[4];
  int nseg;
};
// ONE CTA.  Total count <= ar_capacity.  parity = epoch & 1 selects one of two slot sets.
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2PArgs s, int channel, P2PSegs g) {
  const unsigned e = s.epoch[channel] + 1u;
  const size_t set = (size_t)(e & 1u) * s.world * s.ar_capacity;
  int off = 0;
  for (int q = 0; q < g.nseg; ++q) {
    for (int i = threadIdx.x; i < g.count[q]; i += blockDim.x) {
      const float v = g.ptr[q][i];
      for (int p = 0; p < s.world; ++p)
        reinterpret_cast<float*>(s.ar_peer[p])[set + (size_t)s.rank * s.ar_capacity + off + i] = v;
    }
    off += g.count[q];
  }
  p2p_signal_wait(s, channel);
  const float* mine = reinterpret_cast<const float*>(s.ar_peer[s.rank]) + set;
  off = 0;
  for (int q = 0; q < g.nseg; ++q) {
    for (int i = threadIdx.x; i < g.count[q]; i += blockDim.x) {
      float acc = 0.f;
      for (int p = 0; p < s.world; ++p) acc += __ldcv(mine + (size_t)p * s.ar_capacity + off + i);   // rank order: identical everywhere
      g.ptr[q][i] = acc;
    }
    off += g.count[q];
  }
}

cudaError_t launch_halo_push_signal(const P2PArgs& s, int channel, int n, const int* src_row, int row0,
                                    const unsigned long long* dst_q, const unsigned long long* dst_x, const float* Q,
                                    const float* x, int sms, cudaStream_t st) {
  int grid = (n + 7) / 8;
  grid = grid < 1 ? 1 : (grid > 2 * sms ? 2 * sms : grid);
  halo_push_signal_kernel<<<grid, 256, 0, st>>>(s, channel, n, src_row, row0, dst_q, dst_x, Q, x); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_halo_reduce_apply(int n_rows, const int* rows, const int* ptr, const int* slots, const float* rq,
                                     const float* rx, float* gQ, float* gx, cudaStream_t st) {
  if (n_rows == 0) return cudaSuccess;
  halo_reduce_apply_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(n_rows, rows, ptr, slots, rq, rx, gQ, gx); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_p2p_allreduce(const P2PArgs& s, int channel, const P2PSegs& g, cudaStream_t st) {
  p2p_allreduce_kernel<<<1, 256, 0, st>>>(s, channel, g); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
