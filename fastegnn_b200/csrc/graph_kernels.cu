// graph_kernels.cu -- per-graph phases (B x C x H sized; launch-latency work).
//
//  graph_pre   xbar (models/FastEGNN.py:212), centred Gram M (:213-214) and the per-(graph,
//              channel) constant part of phi_ev's first Linear, G1 = V1s S_c + V1m M[:,c] (:112-115).
//  graph_post  Z' = Z + mean_i(D phi_X(u)) (:147-149) and S' = S + phi_hv([S ; mean_i u]) (:168-177).
// One CTA per graph (forward) / a few CTAs striding over graphs (backward, so that weight
// gradients are reduced in registers before they touch memory).
// Spec: oracle/staged.py graph_pre / graph_post / graph_post_bwd / graph_pre_bwd.
#include "common.cuh"

namespace fegnn {

struct GraphArgs {
  int B, C, ldv;
  unsigned flags;
  const float *inv_nb, *Z, *S, *xsum;
  const float *wv1;                          // edge_mlp_virtual.0.weight
  const float *nodev_w0, *nodev_b0, *nodev_w2, *nodev_b2;
  float *M, *Zc, *G1;                        // graph_pre outputs
  const float *Dsum, *Usum;                  // graph_post inputs
  float *Z_new, *S_new;
  // backward
  const float *gZ_new, *gS_new, *gG1;
  float *gZ, *gS, *gDsum, *gUsum, *gxsum;
  float *g_wv1, *g_nodev_w0, *g_nodev_b0, *g_nodev_w2, *g_nodev_b2;
};

constexpr int kPerThread = (FEGNN_MAX_C * kH + kThreads - 1) / kThreads;

struct __align__(16) GraphSmem {
  float a[FEGNN_MAX_C * kH];     // S / at                 (float4-accessed arrays first: 16-byte aligned)
  float b[FEGNN_MAX_C * kH];     // Um / gzt1 / gG1
  float c[FEGNN_MAX_C * kH];     // dsilu(zt1) / at
  float d[FEGNN_MAX_C * kH];     // gzt2
  float Zc[3 * FEGNN_MAX_C];
  float M[FEGNN_MAX_C * FEGNN_MAX_C];
  float gM[FEGNN_MAX_C * FEGNN_MAX_C];
  float gZc[3 * FEGNN_MAX_C];
  float xbar[4];
};

// 64x64 weight block of a reference-layout matrix -> shared memory, row stride 65 (conflict-free both for
// "thread n walks row n" and "thread k walks column k").  All 16 loads of a thread are independent.
constexpr int kWPad = kH + 1;
constexpr int kWPadFloats = kH * kWPad;
__device__ __forceinline__ void stage_plain(float* __restrict__ Wsm, const float* __restrict__ g, int ld, int off) {
#pragma unroll 16
  for (int i = threadIdx.x; i < kWFloats; i += kThreads) {
    const int n = i >> 6, k = i & 63;
    Wsm[n * kWPad + k] = g[(size_t)n * ld + off + k];
  }
}
constexpr size_t kGraphSmemBytes = sizeof(GraphSmem) + 3 * kWPadFloats * sizeof(float);

__global__ void __launch_bounds__(kThreads) graph_pre_fwd_kernel(GraphArgs a) {
  extern __shared__ __align__(16) unsigned char graph_smem_raw[];
  GraphSmem& s = *reinterpret_cast<GraphSmem*>(graph_smem_raw);
  float* V1s = reinterpret_cast<float*>(graph_smem_raw + sizeof(GraphSmem));
  const int b = blockIdx.x, C = a.C, tid = threadIdx.x;
  stage_plain(V1s, a.wv1, a.ldv, kH);
  if (tid < 3) s.xbar[tid] = a.xsum[b * 3 + tid] * a.inv_nb[b];
  for (int i = tid; i < C * kH; i += kThreads) s.a[i] = a.S[(size_t)b * C * kH + i];
  __syncthreads();
  if (tid < 3 * C) {
    float z = a.Z[(size_t)b * 3 * C + tid] - s.xbar[tid / C];
    s.Zc[tid] = z;
    a.Zc[(size_t)b * 3 * C + tid] = z;
  }
  __syncthreads();
  if (tid < C * C) {
    int c = tid / C, d = tid - c * C;
    float m = s.Zc[c] * s.Zc[d] + s.Zc[C + c] * s.Zc[C + d] + s.Zc[2 * C + c] * s.Zc[2 * C + d];
    s.M[tid] = m;
    a.M[(size_t)b * C * C + tid] = m;
  }
  __syncthreads();
  for (int o = tid; o < C * kH; o += kThreads) {
    const int c = o >> 6, n = o & 63;
    float acc = 0.f;
#pragma unroll 16
    for (int k = 0; k < kH; ++k) acc = fmaf(s.a[c * kH + k], V1s[n * kWPad + k], acc);
    const float* wrow = a.wv1 + (size_t)n * a.ldv + 2 * kH + 1;
    for (int d = 0; d < C; ++d) acc = fmaf(wrow[d], s.M[d * C + c], acc);
    a.G1[(size_t)b * C * kH + o] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) graph_post_fwd_kernel(GraphArgs a) {
  extern __shared__ __align__(16) unsigned char graph_smem_raw[];
  GraphSmem& s = *reinterpret_cast<GraphSmem*>(graph_smem_raw);
  float* T1s = reinterpret_cast<float*>(graph_smem_raw + sizeof(GraphSmem));
  float* T1a = T1s + kWPadFloats;
  float* T2 = T1a + kWPadFloats;
  const int b = blockIdx.x, C = a.C, tid = threadIdx.x;
  const float inb = a.inv_nb[b];
  if (tid < 3 * C) a.Z_new[(size_t)b * 3 * C + tid] = a.Z[(size_t)b * 3 * C + tid] + a.Dsum[(size_t)b * 3 * C + tid] * inb;
  if (a.flags & FEGNN_F_LAST) return;
  stage_plain(T1s, a.nodev_w0, 2 * kH, 0);
  stage_plain(T1a, a.nodev_w0, 2 * kH, kH);
  stage_plain(T2, a.nodev_w2, kH, 0);
  for (int i = tid; i < C * kH; i += kThreads) {
    s.a[i] = a.S[(size_t)b * C * kH + i];
    s.b[i] = a.Usum[(size_t)b * C * kH + i] * inb;
  }
  __syncthreads();
  for (int o = tid; o < C * kH; o += kThreads) {
    const int c = o >> 6, n = o & 63;
    float acc = a.nodev_b0[n];
#pragma unroll 16
    for (int k = 0; k < kH; ++k) acc = fmaf(s.a[c * kH + k], T1s[n * kWPad + k], acc);
#pragma unroll 16
    for (int k = 0; k < kH; ++k) acc = fmaf(s.b[c * kH + k], T1a[n * kWPad + k], acc);
    s.c[o] = silu_f(acc);
  }
  __syncthreads();
  for (int o = tid; o < C * kH; o += kThreads) {
    const int c = o >> 6, n = o & 63;
    float acc = a.nodev_b2[n];
#pragma unroll 16
    for (int k = 0; k < kH; ++k) acc = fmaf(s.c[c * kH + k], T2[n * kWPad + k], acc);
    a.S_new[(size_t)b * C * kH + o] = s.a[o] + acc;
  }
}

// thread (tn = tid>>4, tk = tid&15) owns dW[n = tn*4+jn][k = tk*4+jk] += sum_c G[c][n] * A[c][k]
__device__ __forceinline__ void small_outer_acc(float (&wg)[4][4], const float* G, const float* A, int C) {
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
  for (int c = 0; c < C; ++c) {
    float4 g = *reinterpret_cast<const float4*>(G + c * kH + tn * 4);
    float4 v = *reinterpret_cast<const float4*>(A + c * kH + tk * 4);
    wg[0][0] = fmaf(g.x, v.x, wg[0][0]); wg[0][1] = fmaf(g.x, v.y, wg[0][1]);
    wg[0][2] = fmaf(g.x, v.z, wg[0][2]); wg[0][3] = fmaf(g.x, v.w, wg[0][3]);
    wg[1][0] = fmaf(g.y, v.x, wg[1][0]); wg[1][1] = fmaf(g.y, v.y, wg[1][1]);
    wg[1][2] = fmaf(g.y, v.z, wg[1][2]); wg[1][3] = fmaf(g.y, v.w, wg[1][3]);
    wg[2][0] = fmaf(g.z, v.x, wg[2][0]); wg[2][1] = fmaf(g.z, v.y, wg[2][1]);
    wg[2][2] = fmaf(g.z, v.z, wg[2][2]); wg[2][3] = fmaf(g.z, v.w, wg[2][3]);
    wg[3][0] = fmaf(g.w, v.x, wg[3][0]); wg[3][1] = fmaf(g.w, v.y, wg[3][1]);
    wg[3][2] = fmaf(g.w, v.z, wg[3][2]); wg[3][3] = fmaf(g.w, v.w, wg[3][3]);
  }
}

__global__ void __launch_bounds__(kThreads) graph_post_bwd_kernel(GraphArgs a) {
  extern __shared__ __align__(16) unsigned char graph_smem_raw[];
  GraphSmem& s = *reinterpret_cast<GraphSmem*>(graph_smem_raw);
  float* T1s = reinterpret_cast<float*>(graph_smem_raw + sizeof(GraphSmem));
  float* T1a = T1s + kWPadFloats;
  float* T2 = T1a + kWPadFloats;
  const int C = a.C, tid = threadIdx.x;
  const bool last = a.flags & FEGNN_F_LAST;
  if (!last) {
    stage_plain(T1s, a.nodev_w0, 2 * kH, 0);
    stage_plain(T1a, a.nodev_w0, 2 * kH, kH);
    stage_plain(T2, a.nodev_w2, kH, 0);
  }
  float wgT2[4][4], wgT1s[4][4], wgT1a[4][4];
  zero_wg(wgT2); zero_wg(wgT1s); zero_wg(wgT1a);
  float bf1 = 0.f, bf2 = 0.f;   // thread n < 64 owns the bias sums of column n
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float inb = a.inv_nb[b];
    if (tid < 3 * C) {
      float g = a.gZ_new[(size_t)b * 3 * C + tid];
      a.gZ[(size_t)b * 3 * C + tid] = g;
      a.gDsum[(size_t)b * 3 * C + tid] = g * inb;
    }
    if (last) {                 // no phi_hv: S' = S is discarded (top layer, gS_new == null) or passed through (FastRF)
      for (int i = tid; i < C * kH; i += kThreads) {
        a.gS[(size_t)b * C * kH + i] = a.gS_new != nullptr ? a.gS_new[(size_t)b * C * kH + i] : 0.f;
        a.gUsum[(size_t)b * C * kH + i] = 0.f;
      }
      continue;
    }
    __syncthreads();
    for (int i = tid; i < C * kH; i += kThreads) {
      s.a[i] = a.S[(size_t)b * C * kH + i];
      s.b[i] = a.Usum[(size_t)b * C * kH + i] * inb;
      s.d[i] = a.gS_new[(size_t)b * C * kH + i];
    }
    __syncthreads();
    // zt1 -> at (s.c) ; keep silu' in registers of the owning thread via recompute below
    float dts[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int o = tid + q * kThreads;
      dts[q] = 0.f;
      if (o < C * kH) {
        const int c = o >> 6, n = o & 63;
        float acc = a.nodev_b0[n];
#pragma unroll 16
        for (int k = 0; k < kH; ++k) acc = fmaf(s.a[c * kH + k], T1s[n * kWPad + k], acc);
#pragma unroll 16
        for (int k = 0; k < kH; ++k) acc = fmaf(s.b[c * kH + k], T1a[n * kWPad + k], acc);
        float at, dt;
        silu_grad_f(acc, at, dt);
        s.c[o] = at;
        dts[q] = dt;
      }
    }
    __syncthreads();
    // dT2 += gzt2^T at ; df2 += sum_c gzt2
    small_outer_acc(wgT2, s.d, s.c, C);
    if (tid < kH) for (int c = 0; c < C; ++c) bf2 += s.d[c * kH + tid];
    // gzt1[c][k] = (sum_n gzt2[c][n] T2[n][k]) * silu'(zt1[c][k])   -> overwrite s.c after all reads of at
    float gz[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int o = tid + q * kThreads;
      gz[q] = 0.f;
      if (o < C * kH) {
        const int c = o >> 6, k = o & 63;
        float acc = 0.f;
#pragma unroll 16
        for (int n = 0; n < kH; ++n) acc = fmaf(s.d[c * kH + n], T2[n * kWPad + k], acc);
        gz[q] = acc * dts[q];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int o = tid + q * kThreads;
      if (o < C * kH) s.c[o] = gz[q];
    }
    __syncthreads();
    // dT1s += gzt1^T S ; dT1a += gzt1^T Um ; df1 += sum_c gzt1
    small_outer_acc(wgT1s, s.c, s.a, C);
    small_outer_acc(wgT1a, s.c, s.b, C);
    if (tid < kH) for (int c = 0; c < C; ++c) bf1 += s.c[c * kH + tid];
    // gS = gS' + gzt1 T1s ; gUsum = (gzt1 T1a) * inv_nb
    for (int o = tid; o < C * kH; o += kThreads) {
      const int c = o >> 6, k = o & 63;
      float accs = s.d[o], accu = 0.f;
#pragma unroll 16
      for (int n = 0; n < kH; ++n) {
        const float g = s.c[c * kH + n];
        accs = fmaf(g, T1s[n * kWPad + k], accs);
        accu = fmaf(g, T1a[n * kWPad + k], accu);
      }
      a.gS[(size_t)b * C * kH + o] = accs;
      a.gUsum[(size_t)b * C * kH + o] = accu * inb;
    }
  }
  if (!last) {
    wgrad_flush(wgT2, a.g_nodev_w2, kH, 0, 1);
    wgrad_flush(wgT1s, a.g_nodev_w0, 2 * kH, 0, 1);
    wgrad_flush(wgT1a, a.g_nodev_w0, 2 * kH, kH, 1);
    if (tid < kH) {
      if (a.g_nodev_b0 != nullptr) atomicAdd(a.g_nodev_b0 + tid, bf1);
      if (a.g_nodev_b2 != nullptr) atomicAdd(a.g_nodev_b2 + tid, bf2);
    }
  }
}

__global__ void __launch_bounds__(kThreads) graph_pre_bwd_kernel(GraphArgs a) {
  extern __shared__ __align__(16) unsigned char graph_smem_raw[];
  GraphSmem& s = *reinterpret_cast<GraphSmem*>(graph_smem_raw);
  float* V1s = reinterpret_cast<float*>(graph_smem_raw + sizeof(GraphSmem));
  const int C = a.C, tid = threadIdx.x;
  stage_plain(V1s, a.wv1, a.ldv, kH);
  float wgV1s[4][4];
  zero_wg(wgV1s);
  float gV1m[kPerThread];
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) gV1m[q] = 0.f;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < C * kH; i += kThreads) {
      s.a[i] = a.S[(size_t)b * C * kH + i];
      s.b[i] = a.gG1[(size_t)b * C * kH + i];
    }
    if (tid < C * C) s.M[tid] = a.M[(size_t)b * C * C + tid];
    if (tid < 3 * C) s.Zc[tid] = a.Zc[(size_t)b * 3 * C + tid];
    __syncthreads();
    // dV1s += gG1^T S
    small_outer_acc(wgV1s, s.b, s.a, C);
    // dV1m[n][d] += sum_c gG1[c][n] M[d][c]
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int o = tid + q * kThreads;
      if (o < kH * C) {
        const int n = o / C, d = o - n * C;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(s.b[c * kH + n], s.M[d * C + c], acc);
        gV1m[q] += acc;
      }
    }
    // gS[c][k] += sum_n gG1[c][n] V1s[n][k]
    for (int o = tid; o < C * kH; o += kThreads) {
      const int c = o >> 6, k = o & 63;
      float acc = 0.f;
#pragma unroll 16
      for (int n = 0; n < kH; ++n) acc = fmaf(s.b[c * kH + n], V1s[n * kWPad + k], acc);
      a.gS[(size_t)b * C * kH + o] += acc;
    }
    // gM[d][c] = sum_n V1m[n][d] gG1[c][n]
    if (tid < C * C) {
      const int d = tid / C, c = tid - d * C;
      float acc = 0.f;
      for (int n = 0; n < kH; ++n) acc = fmaf(a.wv1[(size_t)n * a.ldv + 2 * kH + 1 + d], s.b[c * kH + n], acc);
      s.gM[tid] = acc;
    }
    __syncthreads();
    // gZc[a][c] = sum_d Zc[a][d] (gM[c][d] + gM[d][c])
    if (tid < 3 * C) {
      const int ax = tid / C, c = tid - ax * C;
      float acc = 0.f;
      for (int d = 0; d < C; ++d) acc = fmaf(s.Zc[ax * C + d], s.gM[c * C + d] + s.gM[d * C + c], acc);
      s.gZc[tid] = acc;
      a.gZ[(size_t)b * 3 * C + tid] += acc;
    }
    __syncthreads();
    if (tid < 3) {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc += s.gZc[tid * C + c];
      a.gxsum[(size_t)b * 3 + tid] = -acc * a.inv_nb[b];
    }
  }
  if (a.g_wv1 != nullptr) {
    wgrad_flush(wgV1s, a.g_wv1, a.ldv, kH, 1);
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int o = tid + q * kThreads;
      if (o < kH * C) {
        const int n = o / C, d = o - n * C;
        atomicAdd(a.g_wv1 + (size_t)n * a.ldv + 2 * kH + 1 + d, gV1m[q]);
      }
    }
  }
}

#define FEGNN_GRAPH_SMEM(kernel)                                                                              \
  do {                                                                                                       \
    static DevOnce done_;                                                                               \
    if (!done_.get()) {                                                                                            \
      cudaError_t e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGraphSmemBytes); \
      if (e_ != cudaSuccess) return e_;                                                                      \
      done_.set();                                                                                          \
    }                                                                                                        \
  } while (0)

cudaError_t launch_graph_pre_fwd(const GraphArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  FEGNN_GRAPH_SMEM(graph_pre_fwd_kernel);
  graph_pre_fwd_kernel<<<a.B, kThreads, kGraphSmemBytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_graph_post_fwd(const GraphArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  FEGNN_GRAPH_SMEM(graph_post_fwd_kernel);
  graph_post_fwd_kernel<<<a.B, kThreads, kGraphSmemBytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_graph_post_bwd(const GraphArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  FEGNN_GRAPH_SMEM(graph_post_bwd_kernel);
  graph_post_bwd_kernel<<<a.B < 32 ? a.B : 32, kThreads, kGraphSmemBytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_graph_pre_bwd(const GraphArgs& a, cudaStream_t st) {
  if (a.B == 0) return cudaSuccess;
  FEGNN_GRAPH_SMEM(graph_pre_bwd_kernel);
  graph_pre_bwd_kernel<<<a.B < 32 ? a.B : 32, kThreads, kGraphSmemBytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
