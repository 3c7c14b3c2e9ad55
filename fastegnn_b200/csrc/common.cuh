// common.cuh -- device building blocks shared by every phase kernel.
//
// Geometry of all tile kernels: 256 threads, a tile of TM = 128 rows (edges, nodes or
// (node,channel) pairs) x 64 columns.  Thread (ty = tid>>4, tx = tid&15) owns rows
// ty*8 .. ty*8+7 and columns tx*4 .. tx*4+3 of every GEMM result, so epilogues that
// follow one another (silu, its derivative, products with upstream gradients) stay in
// the same thread and never round-trip through memory.
//
// Shared-memory layouts
//   activation tile  A[128][64]   row-major, unpadded (all access patterns used here
//                                 are conflict-free: quarter-warp-uniform float4 reads,
//                                 contiguous float4 row writes, one-row-per-warp walks)
//   weight tile      W[64][64]    the reference [out=n][in=k] matrix with the 16-byte
//                                 chunk index XOR-swizzled by (n>>2)&7, so that ONE copy
//                                 serves both  y = a W^T  (gemm_nt, forward) and
//                                 g_in = g W  (gemm_nn, backward) without bank conflicts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/fegnn.h"

namespace fegnn {

constexpr int kH = FEGNN_H;
constexpr int kThreads = 256;
constexpr int kTM = 128;        // rows per tile
constexpr int kRT = 8;          // rows per thread
constexpr int kTileFloats = kTM * kH;
constexpr int kWFloats = kH * kH;

__device__ __forceinline__ float sigmoid_f(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float silu_f(float z) { return z * sigmoid_f(z); }
// a = silu(z), d = silu'(z) = s (1 + z (1 - s))
__device__ __forceinline__ void silu_grad_f(float z, float& a, float& d) {
  float s = sigmoid_f(z);
  a = z * s;
  d = s + a * (1.f - s);
}

__device__ __forceinline__ int wswz(int n, int k) { return n * kH + ((((k >> 2) ^ ((n >> 2) & 7)) << 2) | (k & 3)); }

// Stage a 64x64 block of a reference-layout weight into the swizzled tile.
// element (n,k) = g[n*ld + off + k*kstride]
__device__ __forceinline__ void stage_weight(float* __restrict__ Ws, const float* __restrict__ g, int ld, int off,
                                             int kstride) {
  // Every load of a thread is issued before the first store (one global-memory round trip per tile, not 16).
  if (kstride == 1 && ((ld | off) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    float4 w[kWFloats / 4 / kThreads];
#pragma unroll
    for (int j = 0; j < kWFloats / 4 / kThreads; ++j) {
      const int i = threadIdx.x + j * kThreads, n = i >> 4, c = i & 15;
      w[j] = *reinterpret_cast<const float4*>(g + (size_t)n * ld + off + c * 4);
    }
#pragma unroll
    for (int j = 0; j < kWFloats / 4 / kThreads; ++j) {
      const int i = threadIdx.x + j * kThreads, n = i >> 4, c = i & 15;
      *reinterpret_cast<float4*>(Ws + n * kH + ((c ^ ((n >> 2) & 7)) << 2)) = w[j];
    }
    return;
  }
  float w[kWFloats / kThreads];
#pragma unroll
  for (int j = 0; j < kWFloats / kThreads; ++j) {
    const int i = threadIdx.x + j * kThreads, n = i >> 6, k = i & 63;
    w[j] = g[(size_t)n * ld + off + (size_t)k * kstride];
  }
#pragma unroll
  for (int j = 0; j < kWFloats / kThreads; ++j) {
    const int i = threadIdx.x + j * kThreads, n = i >> 6, k = i & 63;
    Ws[wswz(n, k)] = w[j];
  }
}
__device__ __forceinline__ void stage_vec(float* __restrict__ s, const float* __restrict__ g, int n, int stride = 1) {
  for (int i = threadIdx.x; i < n; i += kThreads) s[i] = g[(size_t)i * stride];
}

__device__ __forceinline__ void zero_acc(float (&acc)[kRT][4]) {
#pragma unroll
  for (int i = 0; i < kRT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// acc[i][j] += sum_k A[ty*8+i][k] * W[tx*4+j][k]          (y = a W^T)
__device__ __forceinline__ void gemm_nt(float (&acc)[kRT][4], const float* __restrict__ A,
                                        const float* __restrict__ W, int ty, int tx) {
  const float* a0 = A + ty * kRT * kH;
  const float* w0 = W + tx * 4 * kH;
  const int sw = tx & 7;
#pragma unroll 2
  for (int k4 = 0; k4 < 16; ++k4) {
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(w0 + j * kH + ((k4 ^ sw) << 2));
#pragma unroll
    for (int i = 0; i < kRT; ++i) {
      float4 a = *reinterpret_cast<const float4*>(a0 + i * kH + (k4 << 2));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a.x, w[j].x, acc[i][j]);
        acc[i][j] = fmaf(a.y, w[j].y, acc[i][j]);
        acc[i][j] = fmaf(a.z, w[j].z, acc[i][j]);
        acc[i][j] = fmaf(a.w, w[j].w, acc[i][j]);
      }
    }
  }
}

// acc[i][j] += sum_n G[ty*8+i][n] * W[n][tx*4+j]          (g_in = g W)
__device__ __forceinline__ void gemm_nn(float (&acc)[kRT][4], const float* __restrict__ G,
                                        const float* __restrict__ W, int ty, int tx) {
  const float* g0 = G + ty * kRT * kH;
#pragma unroll 2
  for (int n4 = 0; n4 < 16; ++n4) {
    float4 w[4];
    const int chunk = (tx ^ (n4 & 7)) << 2;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) w[jj] = *reinterpret_cast<const float4*>(W + (n4 * 4 + jj) * kH + chunk);
#pragma unroll
    for (int i = 0; i < kRT; ++i) {
      float4 g = *reinterpret_cast<const float4*>(g0 + i * kH + (n4 << 2));
      acc[i][0] = fmaf(g.x, w[0].x, acc[i][0]); acc[i][1] = fmaf(g.x, w[0].y, acc[i][1]);
      acc[i][2] = fmaf(g.x, w[0].z, acc[i][2]); acc[i][3] = fmaf(g.x, w[0].w, acc[i][3]);
      acc[i][0] = fmaf(g.y, w[1].x, acc[i][0]); acc[i][1] = fmaf(g.y, w[1].y, acc[i][1]);
      acc[i][2] = fmaf(g.y, w[1].z, acc[i][2]); acc[i][3] = fmaf(g.y, w[1].w, acc[i][3]);
      acc[i][0] = fmaf(g.z, w[2].x, acc[i][0]); acc[i][1] = fmaf(g.z, w[2].y, acc[i][1]);
      acc[i][2] = fmaf(g.z, w[2].z, acc[i][2]); acc[i][3] = fmaf(g.z, w[2].w, acc[i][3]);
      acc[i][0] = fmaf(g.w, w[3].x, acc[i][0]); acc[i][1] = fmaf(g.w, w[3].y, acc[i][1]);
      acc[i][2] = fmaf(g.w, w[3].z, acc[i][2]); acc[i][3] = fmaf(g.w, w[3].w, acc[i][3]);
    }
  }
}

// Weight-gradient tile: wg[jn][jk] += sum_r G[r][tn*4+jn] * A[r][tk*4+jk]   (dW = g^T a)
// thread (tn = tid>>4, tk = tid&15) owns the 4x4 block (n = tn*4.., k = tk*4..) of dW[n][k].
__device__ __forceinline__ void wgrad_acc(float (&wg)[4][4], const float* __restrict__ G,
                                          const float* __restrict__ A, int nrows) {
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
#pragma unroll 4
  for (int r = 0; r < nrows; ++r) {
    float4 g = *reinterpret_cast<const float4*>(G + r * kH + tn * 4);
    float4 a = *reinterpret_cast<const float4*>(A + r * kH + tk * 4);
    wg[0][0] = fmaf(g.x, a.x, wg[0][0]); wg[0][1] = fmaf(g.x, a.y, wg[0][1]);
    wg[0][2] = fmaf(g.x, a.z, wg[0][2]); wg[0][3] = fmaf(g.x, a.w, wg[0][3]);
    wg[1][0] = fmaf(g.y, a.x, wg[1][0]); wg[1][1] = fmaf(g.y, a.y, wg[1][1]);
    wg[1][2] = fmaf(g.y, a.z, wg[1][2]); wg[1][3] = fmaf(g.y, a.w, wg[1][3]);
    wg[2][0] = fmaf(g.z, a.x, wg[2][0]); wg[2][1] = fmaf(g.z, a.y, wg[2][1]);
    wg[2][2] = fmaf(g.z, a.z, wg[2][2]); wg[2][3] = fmaf(g.z, a.w, wg[2][3]);
    wg[3][0] = fmaf(g.w, a.x, wg[3][0]); wg[3][1] = fmaf(g.w, a.y, wg[3][1]);
    wg[3][2] = fmaf(g.w, a.z, wg[3][2]); wg[3][3] = fmaf(g.w, a.w, wg[3][3]);
  }
}
__device__ __forceinline__ void zero_wg(float (&wg)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) wg[i][j] = 0.f;
}
// Flush a per-CTA dW block into a reference-layout gradient: dst[n*ld + off + k*kstride] += wg
__device__ __forceinline__ void wgrad_flush(const float (&wg)[4][4], float* __restrict__ dst, int ld, int off,
                                            int kstride) {
  if (dst == nullptr) return;
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
  if (kstride == 1 && ((ld | off) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)      // one red.global.add.v4.f32 per 4 consecutive k
      atomicAdd(reinterpret_cast<float4*>(dst + (size_t)(tn * 4 + jn) * ld + off + tk * 4),
                make_float4(wg[jn][0], wg[jn][1], wg[jn][2], wg[jn][3]));
    return;
  }
#pragma unroll
  for (int jn = 0; jn < 4; ++jn)
#pragma unroll
    for (int jk = 0; jk < 4; ++jk)
      atomicAdd(dst + (size_t)(tn * 4 + jn) * ld + off + (size_t)(tk * 4 + jk) * kstride, wg[jn][jk]);
}

// Sum over the 16 threads that share a row (a half-warp: same ty, tx = 0..15).
__device__ __forceinline__ float rowsum16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float warpsum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return rowsum16(v);
}

// Column sums kept per thread (its 4 columns over its 8 rows, across tiles) are
// reduced over the 16 ty-threads with atomics at kernel end.
__device__ __forceinline__ void colsum_flush(const float (&cs)[4], float* __restrict__ dst, int stride, int tx) {
  if (dst == nullptr) return;
  // the two half-warps of a warp hold the same columns (ty even / odd): add them first
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float v = cs[j] + __shfl_xor_sync(0xffffffffu, cs[j], 16);
    if ((threadIdx.x & 16) == 0) atomicAdd(dst + (size_t)(tx * 4 + j) * stride, v);
  }
}

// Shared-memory form for kernels with many column-sum vectors: vector v of every thread is stashed in
// scratch[v][ty][64] (NV * 1024 floats), then 64 threads per vector add the 16 ty-partials and issue ONE atomic per
// column and CTA.  All threads must call; a __syncthreads() is needed before scratch is reused.
struct ColsumDst {
  float* dst;
  int stride;
};
__device__ __forceinline__ void colsum_stash(float* __restrict__ scratch, int v, const float (&cs)[4]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  *reinterpret_cast<float4*>(scratch + v * (16 * kH) + ty * kH + tx * 4) = make_float4(cs[0], cs[1], cs[2], cs[3]);
}
template <int NV>
__device__ __forceinline__ void colsum_emit(const float* __restrict__ scratch, const ColsumDst (&d)[NV]) {
  __syncthreads();
  const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if ((v & 3) == grp && d[v].dst != nullptr) {
      float sum = 0.f;
#pragma unroll
      for (int ty = 0; ty < 16; ++ty) sum += scratch[v * (16 * kH) + ty * kH + col];
      atomicAdd(d[v].dst + (size_t)col * d[v].stride, sum);
    }
  }
}

// Segmented inclusive sum inside a warp over lanes with equal, contiguous keys.
// Returns the segment total in the LAST lane of each segment (is_tail == true there).
__device__ __forceinline__ float warp_segsum(float v, int key, int lane, bool& is_tail) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float vu = __shfl_up_sync(0xffffffffu, v, o);
    int ku = __shfl_up_sync(0xffffffffu, key, o);
    if (lane >= o && ku == key) v += vu;
  }
  int kd = __shfl_down_sync(0xffffffffu, key, 1);
  is_tail = (lane == 31) || (kd != key);
  return v;
}

// Column-parallel walk over the rows of a tile with a sorted (contiguous) key per row:
// thread (c = tid&63, grp = tid>>6) sums rows grp*32 .. grp*32+31 of column c and
// flushes one atomicAdd per key run:  dst[key*64 + c] += run sum.  key < 0 rows are skipped.
__device__ __forceinline__ void tile_segsum_rows(const float* __restrict__ T, const int* __restrict__ skey,
                                                 float* __restrict__ dst) {
  const int c = threadIdx.x & 63, grp = threadIdx.x >> 6;
  int cur = -1;
  float acc = 0.f;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    int rr = grp * 32 + i;
    int k = skey[rr];
    if (k != cur) {
      if (cur >= 0) atomicAdd(dst + (size_t)cur * kH + c, acc);
      cur = k;
      acc = 0.f;
    }
    if (k >= 0) acc += T[rr * kH + c];
  }
  if (cur >= 0) atomicAdd(dst + (size_t)cur * kH + c, acc);
}

// Load a [rows<=128][64] global block (row stride ld floats) into a tile; rows >= nvalid are zero.
__device__ __forceinline__ void load_tile(float* __restrict__ T, const float* __restrict__ g, size_t ld, int nvalid,
                                          float scale_unused = 1.f) {
  for (int i = threadIdx.x; i < kTM * 16; i += kThreads) {
    int r = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nvalid) v = *reinterpret_cast<const float4*>(g + (size_t)r * ld + c4 * 4);
    *reinterpret_cast<float4*>(T + r * kH + c4 * 4) = v;
  }
}

// "Done once" flag PER DEVICE: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting, so a process that
// drives several GPUs (model on cuda:1 while the current device was 0 earlier) must opt in on each of them.
struct DevOnce {
  bool done[64] = {};
  static int dev() {
    int d = 0;
    return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : -1;
  }
  bool get() const {
    const int d = dev();
    return d >= 0 && done[d];
  }
  void set() {
    const int d = dev();
    if (d >= 0) done[d] = true;
  }
};

// 16 bytes global -> shared without a register stop; bytes = 0 writes zeros (rows past the end)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
               :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Column sums over the 32 lanes of a warp for 32 per-lane values: lane L returns sum over the lanes of v[L].  Each round
// keeps the half of the columns whose index bit matches the lane's and trades the other half with the partner lane:
// 16 + 8 + 4 + 2 + 1 = 31 shuffles (a butterfly per column costs 160).  v is clobbered.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int n = 16; n >= 1; n >>= 1) {
    const bool hi = lane & n;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = hi ? v[j] : v[j + n];
      const float keep = hi ? v[j + n] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
    }
  }
  return v[0];
}

// clock64 stamps between the barriers of ONE thread of ONE CTA (instrumented build only: -DFEGNN_TRACE, tools/)
#ifdef FEGNN_TRACE
#define VTR_DECL() long long tr_t[64]; int tr_l[64]; int tri_ = 0; const bool tr_on = threadIdx.x == 0 && blockIdx.x == 1; \
  if (tr_on) { tr_t[0] = clock64(); tr_l[0] = __LINE__; tri_ = 1; }
#define VTR() do { if (tr_on && tri_ < 64) { tr_t[tri_] = clock64(); tr_l[tri_] = __LINE__; ++tri_; } } while (0)
#define VTR_PRINT(name) do { if (tr_on) { printf("VTRACE %s :", name); \
  for (int i_ = 1; i_ < tri_; ++i_) printf(" L%d:%lld", tr_l[i_], tr_t[i_] - tr_t[i_ - 1]); printf("\n"); } } while (0)
#else
#define VTR_DECL() do { } while (0)
#define VTR() do { } while (0)
#define VTR_PRINT(name) do { } while (0)
#endif

// Number of kernels this library has launched (host counter; read through fegnn_launch_count()).
inline unsigned long long g_launches = 0;

// Programmatic dependent launch.  Every kernel of the main chain starts with pdl_trigger() (the next kernel of the stream
// may begin its prologue: weight staging, barrier init, tensor-memory allocation) and calls pdl_wait() before it touches
// anything a predecessor produced -- or anything a predecessor may still read (griddepcontrol.wait returns when every
// prerequisite grid has completed and flushed).  Before the wait a kernel reads layer weights only (written by the optimizer,
// never inside a forward / backward).  Both are no-ops in a launch without the attribute.  FEGNN_PDL=0 disables it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() {
  static const bool on = !(getenv("FEGNN_PDL") && atoi(getenv("FEGNN_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ++g_launches;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace fegnn
