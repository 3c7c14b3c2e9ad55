// virtual_tc.cu -- dense N x C real<->virtual phase on the tcgen05 tensor cores (single-pass TF32 tiles).
//
// Same contract as virtual_fwd_kernel / virtual_bwd_kernel (virtual_kernels.cu; reference models/FastEGNN.py:111-119,
// :133-150 and their autograd).  Rows of a tile are (node, channel) pairs, channel fastest, TN = 128 / C whole nodes per
// tile; thread (quarter, lane, cg) owns row 32*quarter + lane (= its tensor-memory lane) and CPT = 64 / CG columns.
//
//   forward   G1  z2 = a1 V2^T                 a1 built K-major in shared memory (half-warp per row, coalesced gather)
//             GH  [zxv | zX] = u [Wxv ; WX]^T  ONE N = 128 GEMM, u read from tensor memory (written by its row owner)
//   backward  two kernels with two dependent GEMM stages each (everything else hangs off the critical path):
//     heads   GH as above from the saved u; A = [gzxv | gzX] in tensor memory;
//             gu = gu_ext + gUsum + [gzxv | gzX] [Wxv ; WX]   (K = 128);   [dWxv ; dWX | db..] += [gzxv | gzX]^T [u | 1..]
//             gu is written over the gu buffer (which then feeds the trunk kernel); coordinate outputs except the rho term.
//     trunk   G1 recomputed for silu'(z2); gz2 = gu silu'(z2); ga1 = gz2 V2; gz1 = ga1 silu'(z1);
//             [dc2.. | dV2] += gz2^T [1.. | a1];  dvr += gz1^T rho;  gAv, gG1, and the rho term of gx / gZ.
// Weight-gradient operands are row-major tiles read MN-major (SWIZZLE_128B_BASE32B), see edge_tc_bwd2.cu.
// attention=True layers use the fp32 FMA kernels.
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace fegnn {
namespace vtc {

using bwd2::desc_advance;
using bwd2::gemm_ts_kmajor;
using bwd2::gemm_ts_mn;
using bwd2::gemm_wgrad;
using bwd2::idesc_tf32;
using bwd2::make_desc_mn;
using bwd2::mma_ts;
using bwd2::mn_chunk_off;
using bwd2::mn_load_row;
using bwd2::mn_store_row;
using bwd2::silu_grad_tc;
using bwd2::tmem_ld;
using bwd2::tmem_st;
using bwd2::tmem_st_wait;

__device__ __forceinline__ float silu_tc(float z) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  return z * fmaf(0.5f, th, 0.5f);
}

// stage a reference [64][ld] weight K-major SWIZZLE_128B into rows [row0, row0+64) of a tile with R rows
template <int NT>
__device__ __forceinline__ void stage_w_kmajor(uint8_t* dst, const float* __restrict__ g, int ld, int row0, int R) {
  for (int i = threadIdx.x; i < kH * 16; i += NT) {
    const int n = i >> 4, c = i & 15;
    *reinterpret_cast<float4*>(dst + umma::tile_chunk_off(row0 + n, c, R)) = *reinterpret_cast<const float4*>(g + (size_t)n * ld + c * 4);
  }
}
// ... and row-major BASE32B (MN-major operand) into rows [row0, row0+64) of a tile with R rows
template <int NT>
__device__ __forceinline__ void stage_w_mn(uint8_t* dst, const float* __restrict__ g, int ld, int row0, int R) {
  for (int i = threadIdx.x; i < kH * 16; i += NT) {
    const int n = i >> 4, c = i & 15;
    *reinterpret_cast<float4*>(dst + mn_chunk_off(row0 + n, c, R)) = *reinterpret_cast<const float4*>(g + (size_t)n * ld + c * 4);
  }
}

// per-graph accumulators of a CTA that stays inside one graph (same policy as virtual_kernels.cu)
struct GraphAcc {
  float big[FEGNN_MAX_C * kH];     // Usum (fwd) / gG1 (bwd trunk)
  float small[3 * FEGNN_MAX_C];    // Dsum (fwd) / gZ (bwd)
  float x3[3];                     // xsum_new (fwd)
};
template <int NT>
__device__ __forceinline__ void acc_clear(GraphAcc* g) {
  for (int i = threadIdx.x; i < FEGNN_MAX_C * kH; i += NT) g->big[i] = 0.f;
  if (threadIdx.x < 3 * FEGNN_MAX_C) g->small[threadIdx.x] = 0.f;
  if (threadIdx.x < 3) g->x3[threadIdx.x] = 0.f;
}
// flush graph b (if any) and clear; all threads
template <int NT>
__device__ __forceinline__ void acc_flush(GraphAcc* g, int b, int C, float* dstBig, float* dstSmall, float* dstX) {
  __syncthreads();
  if (b >= 0) {
    if (dstBig != nullptr)
      for (int i = threadIdx.x; i < C * kH; i += NT) {
        atomicAdd(dstBig + (size_t)b * C * kH + i, g->big[i]);
        g->big[i] = 0.f;
      }
    if (dstSmall != nullptr && threadIdx.x < 3 * C) {
      atomicAdd(dstSmall + (size_t)b * 3 * C + threadIdx.x, g->small[threadIdx.x]);
      g->small[threadIdx.x] = 0.f;
    }
    if (dstX != nullptr && threadIdx.x < 3) {
      atomicAdd(dstX + (size_t)b * 3 + threadIdx.x, g->x3[threadIdx.x]);
      g->x3[threadIdx.x] = 0.f;
    }
  }
  __syncthreads();
}

// Add the rows of a BASE32B tile into the per-(graph, channel) [C][64] sums: thread (col, grp) walks the nodes of its
// group; single -> shared accumulator, otherwise one global atomic per key run.
template <int NT>
__device__ __forceinline__ void rows_to_graph(const uint8_t* T, const int* skey, GraphAcc* g, int C, int TN, bool single,
                                              float* __restrict__ dst) {
  constexpr int GR = NT / 64;
  const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int per = (TN + GR - 1) / GR;
  const int j0 = grp * per, j1 = min(TN, j0 + per);
  const uint32_t cbase = (col >> 5) * (kTM * 128) + ((col & 7) << 2);
  const int c8 = (col & 31) >> 3;
  for (int c = 0; c < C; ++c) {
    int cur = -1;
    float acc = 0.f;
    for (int jn = j0; jn < j1; ++jn) {
      const int r = jn * C + c;
      const int k = skey[r];
      if (!single && k != cur) {
        if (cur >= 0) atomicAdd(dst + (size_t)cur * kH + col, acc);
        cur = k;
        acc = 0.f;
      }
      if (k >= 0) acc += *reinterpret_cast<const float*>(T + cbase + r * 128 + ((c8 ^ (r & 3)) << 5));
    }
    if (single) {
      if (j1 > j0) atomicAdd(&g->big[c * kH + col], acc);
    } else if (cur >= 0) {
      atomicAdd(dst + (size_t)cur * kH + col, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------- forward
struct FwdVec {
  float vr[kH], c2[kH], bh[2 * kH], wh[2 * kH];     // bh = (bxv | bX), wh = (wxv | wX)
  int skey[kTM], snode[kTM], sb[kTM];
  float sD[kTM * 3], srho[kTM], spx[4 * kTM], spX[4 * kTM], ssxv[kTM], ssX[kTM];
  int b_first, b_last;
  GraphAcc acc;
  uint64_t bar[2];
  uint32_t tmem_slot;
};
struct FwdSmem {
  static constexpr int off_V2 = 0;                       // [64][64]  K-major
  static constexpr int off_WH = 16384;                   // [128][64] K-major: rows 0-63 Wxv, 64-127 WX
  static constexpr int off_T = off_WH + 32768;           // a1 K-major for G1, then u BASE32B for the Usum walk
  static constexpr int off_vec = off_T + 32768;
  static constexpr size_t bytes = off_vec + sizeof(FwdVec) + 1024;
};
constexpr uint32_t kF_ACC0 = 0, kF_ACCH = 64, kF_OPA = 192;    // 256 columns

template <int CG>
__global__ void __launch_bounds__(128 * CG, 2) virtual_fwd_tc_kernel(VirtArgs a) {
  constexpr int NT = 128 * CG, CPT = kH / CG, NW = NT / 32;
  using SM = FwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  FwdVec* v = reinterpret_cast<FwdVec*>(smem + SM::off_vec);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int quarter = warp & 3, cg = warp >> 2, row = quarter * 32 + lane, c0 = cg * CPT;
  const int C = a.C, TN = kTM / C;
  const bool use_tanh = a.flags & FEGNN_F_TANH, grav = a.flags & FEGNN_F_GRAVITY;

  stage_w_kmajor<NT>(smem + SM::off_V2, a.V2, kH, 0, kH);
  stage_w_kmajor<NT>(smem + SM::off_WH, a.Wxv, kH, 0, 2 * kH);
  stage_w_kmajor<NT>(smem + SM::off_WH, a.WX, kH, kH, 2 * kH);
  for (int i = t; i < kH; i += NT) {
    v->vr[i] = a.wv1[(size_t)i * a.ldv + 2 * kH];
    v->c2[i] = a.c2[i];
    v->bh[i] = a.bxv[i]; v->bh[kH + i] = a.bX[i];
    v->wh[i] = a.wxv[i]; v->wh[kH + i] = a.wX[i];
  }
  acc_clear<NT>(&v->acc);
  if (t == 0) {
    umma::mbar_init(&v->bar[0], 1);
    umma::mbar_init(&v->bar[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<256>(&v->tmem_slot);
  umma::fence_smem_to_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem = v->tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + c0;
  const uint32_t id_g1 = umma::make_idesc_tf32(128, 64), id_gh = umma::make_idesc_tf32(128, 128);
  const uint64_t dV2 = umma::make_desc(umma::smem_u32(smem + SM::off_V2));
  const uint64_t dWH = umma::make_desc(umma::smem_u32(smem + SM::off_WH));
  const uint64_t dT = umma::make_desc(umma::smem_u32(smem + SM::off_T));
  uint8_t* T = smem + SM::off_T;
  uint32_t phase = 0;
  int cur_b = -1;

  const int ntiles = (a.N + TN - 1) / TN;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    umma::fence_before();
    __syncthreads();
    // ---- geometry: one thread per (node, channel) row
    if (t < kTM) {
      const int jn = t / C, c = t - jn * C;
      const int i = tile * TN + jn;
      const bool valid = jn < TN && i < a.N;
      int key = -1, node = -1;
      float D0 = 0, D1 = 0, D2 = 0, rho = 0;
      if (valid) {
        const int b = a.batch[i];
        key = b * C + c;
        node = i;
        D0 = a.Z[((size_t)b * 3 + 0) * C + c] - a.x[(size_t)i * 3 + 0];
        D1 = a.Z[((size_t)b * 3 + 1) * C + c] - a.x[(size_t)i * 3 + 1];
        D2 = a.Z[((size_t)b * 3 + 2) * C + c] - a.x[(size_t)i * 3 + 2];
        rho = sqrtf(D0 * D0 + D1 * D1 + D2 * D2);
        if (c == 0) v->sb[jn] = b;
        if (t == 0) v->b_first = b;
        if (c == 0 && (jn == TN - 1 || i == a.N - 1)) v->b_last = b;
      }
      v->skey[t] = key;
      v->snode[t] = node;
      v->sD[t * 3 + 0] = D0; v->sD[t * 3 + 1] = D1; v->sD[t * 3 + 2] = D2;
      v->srho[t] = rho;
    }
    __syncthreads();
    const bool single = v->b_first == v->b_last;
    if (!single || v->b_first != cur_b) {
      const int nb = single ? v->b_first : -1;
      acc_flush<NT>(&v->acc, cur_b, C, a.Usum, a.Dsum, a.xsum_new);
      cur_b = nb;
    }
    // ---- assembly: a1 = silu(Av[node] + G1[key] + rho vr) -> T (K-major).  Half-warp per row, float4 per lane.
    {
      const int l16 = lane & 15, hsel = lane >> 4;
      const float4 vr = *reinterpret_cast<const float4*>(v->vr + 4 * l16);
      constexpr int RPW = kTM / NW;
#pragma unroll 1
      for (int i0 = 0; i0 < RPW; i0 += 8) {
        float4 p[4], g[4];
        int ki[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * RPW + i0 + 2 * j + hsel;
          ki[j] = v->skey[rr];
          p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ki[j] >= 0) {
            p[j] = *reinterpret_cast<const float4*>(a.Av + (size_t)v->snode[rr] * kH + 4 * l16);
            g[j] = *reinterpret_cast<const float4*>(a.G1 + (size_t)ki[j] * kH + 4 * l16);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * RPW + i0 + 2 * j + hsel;
          const float rho = v->srho[rr];
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ki[j] >= 0)
            o = make_float4(silu_tc(p[j].x + g[j].x + rho * vr.x), silu_tc(p[j].y + g[j].y + rho * vr.y),
                            silu_tc(p[j].z + g[j].z + rho * vr.z), silu_tc(p[j].w + g[j].w + rho * vr.w));
          *reinterpret_cast<float4*>(T + umma::tile_chunk_off(rr, l16, kTM)) = o;
        }
      }
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (umma::elect_one()) {
        umma::gemm_k64_desc(tmem + kF_ACC0, dT, kTM, dV2, kH, id_g1, false);       // G1: z2 = a1 V2^T
        umma::commit(&v->bar[0]);
      }
      __syncwarp();
    }
    umma::mbar_wait(&v->bar[0], phase);
    umma::fence_after();
    // ---- epilogue 1: u = silu(z2 + c2) -> HBM (saved for backward / phi_h), tensor memory (A of GH), T (Usum walk)
    {
      float u[CPT];
      tmem_ld<CPT>(tlane + kF_ACC0, u);
      const bool valid = v->skey[row] >= 0;
#pragma unroll
      for (int j = 0; j < CPT; ++j) u[j] = valid ? silu_tc(u[j] + v->c2[c0 + j]) : 0.f;
      tmem_st<CPT>(tlane + kF_OPA, u);
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.u + ((size_t)tile * TN * C + row) * kH + c0);
#pragma unroll
        for (int ch = 0; ch < CPT / 4; ++ch) dst[ch] = make_float4(u[ch * 4], u[ch * 4 + 1], u[ch * 4 + 2], u[ch * 4 + 3]);
      }
      mn_store_row<CPT>(T, row, cg, u);          // G1 has finished reading T (bar[0])
      tmem_st_wait();
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (umma::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)                                             // GH: [zxv | zX] = u [Wxv ; WX]^T
          mma_ts(tmem + kF_ACCH, tmem + kF_OPA + ks * 8, desc_advance(dWH, (ks >> 2) * (2 * kH * 128) + (ks & 3) * 32), id_gh, ks > 0);
        umma::commit(&v->bar[1]);
      }
      __syncwarp();
    }
    rows_to_graph<NT>(T, v->skey, &v->acc, C, TN, single, a.Usum);                 // overlaps GH
    umma::mbar_wait(&v->bar[1], phase);
    umma::fence_after();
    phase ^= 1;
    // ---- epilogue 2: the two coordinate heads  s = w . silu(z + b)
    {
      float z[CPT];
      tmem_ld<CPT>(tlane + kF_ACCH, z);
      float px = 0.f, pX = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) px = fmaf(silu_tc(z[j] + v->bh[c0 + j]), v->wh[c0 + j], px);
      tmem_ld<CPT>(tlane + kF_ACCH + kH, z);
#pragma unroll
      for (int j = 0; j < CPT; ++j) pX = fmaf(silu_tc(z[j] + v->bh[kH + c0 + j]), v->wh[kH + c0 + j], pX);
      v->spx[cg * kTM + row] = px;
      v->spX[cg * kTM + row] = pX;
    }
    __syncthreads();
    if (t < kTM) {
      float sxv = 0.f, sX = 0.f;
#pragma unroll
      for (int g = 0; g < CG; ++g) { sxv += v->spx[g * kTM + t]; sX += v->spX[g * kTM + t]; }
      if (use_tanh) { sxv = tanhf(sxv); sX = tanhf(sX); }
      v->ssxv[t] = sxv;
      const int key = v->skey[t];
      if (key >= 0) {                     // Dsum[b,:,c] += D * sX
        const int b = key / C, c = key - b * C;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float val = v->sD[t * 3 + k] * sX;
          if (single) atomicAdd(&v->acc.small[k * C + c], val);
          else atomicAdd(a.Dsum + ((size_t)b * 3 + k) * C + c, val);
        }
      }
    }
    __syncthreads();
    if (t < TN) {                         // x' (models/FastEGNN.py:133-142) and its per-graph sum
      const int i = tile * TN + t;
      if (i < a.N) {
        const int b = v->sb[t];
        const float di = a.dinv[i], svi = a.sv[i];
        const float sgi = grav ? a.sg[i] : 0.f;
        const float invC = 1.f / (float)C;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float vsum = 0.f;
          for (int c = 0; c < C; ++c) vsum += v->sD[(t * C + c) * 3 + k] * v->ssxv[t * C + c];
          float xn = a.x[(size_t)i * 3 + k] + a.tsum[(size_t)i * 3 + k] * di - vsum * invC + svi * a.v[(size_t)i * 3 + k];
          if (grav) xn += sgi * a.grav[k];
          a.x_new[(size_t)i * 3 + k] = xn;
          if (single) atomicAdd(&v->acc.x3[k], xn);
          else atomicAdd(a.xsum_new + (size_t)b * 3 + k, xn);
        }
      }
    }
  }
  acc_flush<NT>(&v->acc, cur_b, C, a.Usum, a.Dsum, a.xsum_new);
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

}  // namespace vtc

template <int CG>
cudaError_t launch_virtual_fwd_tc(const VirtArgs& a, int sms, cudaStream_t st) {
  static bool attr = false;
  const size_t bytes = vtc::FwdSmem::bytes;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(vtc::virtual_fwd_tc_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int TN = kTM / a.C;
  const int ntiles = (a.N + TN - 1) / TN;
  if (ntiles == 0) return cudaSuccess;
  const int grid = ntiles < 2 * sms ? ntiles : 2 * sms;
  vtc::virtual_fwd_tc_kernel<CG><<<grid, 128 * CG, bytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
