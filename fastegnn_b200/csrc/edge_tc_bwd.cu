// edge_tc_bwd.cu -- fused real-edge BACKWARD phase on tcgen05 tensor cores (single-pass TF32 tiles).
//
// Same contract as edge_bwd_kernel (edge_kernels.cu; autograd of models/FastEGNN.py:102-108,125-129,156):
// per tile of 128 CSR-sorted edges the forward is recomputed and six GEMMs run on the tensor cores,
// all operands K-major / SWIZZLE_128B in shared memory, all accumulators in TMEM:
//   G1  z2   = a1 W2^T                 (128x64, K=64)        TMEM cols   0.. 63
//   G2  z3   = m  W3^T                                        TMEM cols  64..127
//   D3  gmm  = g3 W3     (B = W3^T)                           TMEM cols 128..191
//   D2  gz1' = g2 W2     (B = W2^T)                           TMEM cols 192..255
//   W3g dW3 += g3^T m    (64x64, K=128 edges; A = g3^T, B = m^T)   cols 256..319, accumulated over all tiles
//   W2g dW2 += g2^T a1                                        cols 320..383, accumulated over all tiles
// The edge-transposed operands (a1^T, m^T, g3^T, g2^T) are written by the epilogue threads next to the
// row-major ones, so no MN-major descriptors are needed.  Thread (warp, lane) owns edge row
// (warp&3)*32+lane (= its TMEM lane) and the 32 columns of half (warp>>2); column sums (bias and
// vector gradients) use a 31-shuffle transpose-reduce per warp; row-segment sums (gP) walk the tile.
// silu'(z1) is kept as an fp16 tile from the assembly for the last epilogue (TF32-grade anyway).
// attention=True layers use the fp32 FMA kernel.
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace fegnn {

struct EdgeTcBwdVec {
  float wq[kH], Wa[kTcMaxFe * kH], b2[kH], b3[kH], w4[kH];
  int srow[kTM], scol[kTM];
  float sq[kTM], snrm[kTM], sd[kTM * 3], sgte[kTM * 3], sgs[kTM], sea[kTM * kTcMaxFe];
  float spart[kTM], ss[kTM], sgq[kTM];
  uint64_t bar[4];
  uint32_t tmem_slot;
};

struct EdgeTcBwdSmem {
  static constexpr int kW = kH * kH * 4;     // 16 KB
  static constexpr int kT = kTM * kH * 4;    // 32 KB
  static constexpr int off_W2 = 0, off_W3 = kW, off_W2T = 2 * kW, off_W3T = 3 * kW;
  static constexpr int off_X0 = 4 * kW;      // a1^T                       [64][128]
  static constexpr int off_X1 = off_X0 + kT; // a1 -> g3 -> g2 -> gz1      [128][64]
  static constexpr int off_X2 = off_X1 + kT; // m  -> g3^T -> g2^T
  static constexpr int off_X3 = off_X2 + kT; // m^T                        [64][128]
  static constexpr int off_D1 = off_X3 + kT; // silu'(z1) as fp16 [128][64], 16-byte chunks swizzled by (row & 7)
  static constexpr int off_vec = off_D1 + kTM * kH * 2;
  static constexpr size_t bytes = off_vec + sizeof(EdgeTcBwdVec) + 1024;
};

// TF32-grade silu and derivative: s = 0.5 + 0.5 tanh.approx(z/2); a = z s; d = s + a (1 - s)
__device__ __forceinline__ void tc_silu_grad(float z, float& a, float& d) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  const float s = fmaf(0.5f, th, 0.5f);
  a = z * s;
  d = fmaf(a, 1.f - s, s);
}

// Sum each of 32 per-lane values over the 32 lanes of the warp; lane j returns the total of value j.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool up = lane & 16;
    const float send = up ? v[j] : v[j + 16], keep = up ? v[j + 16] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool up = lane & 8;
    const float send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool up = lane & 4;
    const float send = up ? v[j] : v[j + 4], keep = up ? v[j + 4] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const bool up = lane & 2;
    const float send = up ? v[j] : v[j + 2], keep = up ? v[j + 2] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// K-major GEMM with nk8 K-steps of 8: D[M x N] (+)= A[a_rows x K] B[b_rows x K]^T
__device__ __forceinline__ void tc_gemm(uint32_t tmem_d, uint32_t a_saddr, int a_rows, uint32_t b_saddr, int b_rows,
                                        int nk8, uint32_t idesc, bool accumulate_first) {
  for (int ks = 0; ks < nk8; ++ks) {
    const uint32_t aoff = (ks >> 2) * (a_rows * 128) + (ks & 3) * 32;
    const uint32_t boff = (ks >> 2) * (b_rows * 128) + (ks & 3) * 32;
    umma::mma_tf32(tmem_d, umma::make_desc(a_saddr + aoff), umma::make_desc(b_saddr + boff), idesc,
                   (ks > 0 || accumulate_first) ? 1u : 0u);
  }
}

// stage W (reference [64][ld]) K-major and its transpose K-major
__device__ __forceinline__ void tc_stage_w_and_wt(uint8_t* w_dst, uint8_t* wt_dst, const float* __restrict__ g, int ld) {
#pragma unroll 4
  for (int i = threadIdx.x; i < kH * 16; i += blockDim.x) {
    const int n = i >> 4, c = i & 15;
    const float4 w = *reinterpret_cast<const float4*>(g + (size_t)n * ld + c * 4);
    *reinterpret_cast<float4*>(w_dst + umma::tile_chunk_off(n, c, kH)) = w;
    // transpose: element (n, k) -> row k, column n of the W^T tile
    *reinterpret_cast<float*>(wt_dst + umma::tile_off(c * 4 + 0, n, kH)) = w.x;
    *reinterpret_cast<float*>(wt_dst + umma::tile_off(c * 4 + 1, n, kH)) = w.y;
    *reinterpret_cast<float*>(wt_dst + umma::tile_off(c * 4 + 2, n, kH)) = w.z;
    *reinterpret_cast<float*>(wt_dst + umma::tile_off(c * 4 + 3, n, kH)) = w.w;
  }
}

// Write this thread's 32 values (row e, columns half*32 .. +31) row-major into N (a [128][64] tile) and
// edge-transposed into T (a [64][128] tile).
__device__ __forceinline__ void tc_store_both(uint8_t* N, uint8_t* T, int e, int half, const float (&v)[32]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<float4*>(N + umma::tile_chunk_off(e, half * 8 + ch, kTM)) =
        make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
  const uint32_t ebase = (uint32_t)((e >> 5) * (kH * 128) + ((e & 3) << 2) + half * 4096);
  const int ech = (e & 31) >> 2;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    *reinterpret_cast<float*>(T + ebase + (j >> 3) * 1024 + (j & 7) * 128 + ((ech ^ (j & 7)) << 4)) = v[j];
}

__global__ void __launch_bounds__(kTcThreads, 1) edge_bwd_tc_kernel(EdgeArgs a) {
  using SM = EdgeTcBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  EdgeTcBwdVec* v = reinterpret_cast<EdgeTcBwdVec*>(smem + SM::off_vec);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int quarter = warp & 3, half = warp >> 2, row = quarter * 32 + lane;
  const bool use_tanh = a.flags & FEGNN_F_TANH, norm = a.flags & FEGNN_F_NORMALIZE;

  tc_stage_w_and_wt(smem + SM::off_W2, smem + SM::off_W2T, a.W2, kH);
  tc_stage_w_and_wt(smem + SM::off_W3, smem + SM::off_W3T, a.W3, kH);
  for (int i = t; i < kH; i += kTcThreads) {
    v->wq[i] = a.w1[(size_t)i * a.ld1 + 2 * kH];
    for (int f = 0; f < a.Fe; ++f) v->Wa[f * kH + i] = a.w1[(size_t)i * a.ld1 + 2 * kH + 1 + f];
    v->b2[i] = a.b2[i];
    v->b3[i] = a.b3[i];
    v->w4[i] = a.w4[i];
  }
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) umma::mbar_init(&v->bar[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(&v->tmem_slot);
  umma::fence_smem_to_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem = v->tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + half * 32;   // this thread's lane, its 32 columns
  const uint32_t idesc128 = umma::make_idesc_tf32(128, 64), idesc64 = umma::make_idesc_tf32(64, 64);
  const uint32_t sW2 = umma::smem_u32(smem + SM::off_W2), sW3 = umma::smem_u32(smem + SM::off_W3);
  const uint32_t sW2T = umma::smem_u32(smem + SM::off_W2T), sW3T = umma::smem_u32(smem + SM::off_W3T);
  const uint32_t sX0 = umma::smem_u32(smem + SM::off_X0), sX1 = umma::smem_u32(smem + SM::off_X1);
  const uint32_t sX2 = umma::smem_u32(smem + SM::off_X2), sX3 = umma::smem_u32(smem + SM::off_X3);
  uint8_t* X0 = smem + SM::off_X0;
  uint8_t* X1 = smem + SM::off_X1;
  uint8_t* X2 = smem + SM::off_X2;
  uint8_t* X3 = smem + SM::off_X3;
  uint8_t* D1 = smem + SM::off_D1;
  uint32_t phase = 0;
  bool first_tile = true;
  // per-thread column accumulators (column half*32+lane, rows of this warp's quarter), flushed at the end
  float c_w4 = 0.f, c_b3 = 0.f, c_b2 = 0.f, c_wq = 0.f;
  float c_Wa[kTcMaxFe];
#pragma unroll
  for (int f = 0; f < kTcMaxFe; ++f) c_Wa[f] = 0.f;

  const int ntiles = (a.E + kTM - 1) / kTM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    umma::fence_before();
    __syncthreads();
    // ---- geometry: one thread per edge
    if (t < kTM) {
      const int e = tile * kTM + t;
      int r = -1, c = 0;
      float d0 = 0, d1 = 0, d2 = 0, q = 0, nrm = 1.f, g0 = 0, g1 = 0, g2 = 0;
      if (e < a.E) {
        r = a.row[e];
        c = a.col[e];
        d0 = a.x[(size_t)r * 3 + 0] - a.x[(size_t)c * 3 + 0];
        d1 = a.x[(size_t)r * 3 + 1] - a.x[(size_t)c * 3 + 1];
        d2 = a.x[(size_t)r * 3 + 2] - a.x[(size_t)c * 3 + 2];
        q = d0 * d0 + d1 * d1 + d2 * d2;
        g0 = a.gt[(size_t)r * 3 + 0]; g1 = a.gt[(size_t)r * 3 + 1]; g2 = a.gt[(size_t)r * 3 + 2];
        for (int f = 0; f < a.Fe; ++f) v->sea[t * kTcMaxFe + f] = a.ea[(size_t)e * a.Fe + f];
      } else {
        for (int f = 0; f < a.Fe; ++f) v->sea[t * kTcMaxFe + f] = 0.f;
      }
      v->sd[t * 3 + 0] = d0; v->sd[t * 3 + 1] = d1; v->sd[t * 3 + 2] = d2;
      if (norm) {
        nrm = sqrtf(q) + a.eps;
        const float inv = 1.f / nrm;
        d0 *= inv; d1 *= inv; d2 *= inv;
      }
      v->srow[t] = r;
      v->scol[t] = c;
      v->sq[t] = q;
      v->snrm[t] = nrm;
      v->sgte[t * 3 + 0] = g0; v->sgte[t * 3 + 1] = g1; v->sgte[t * 3 + 2] = g2;
      v->sgs[t] = d0 * g0 + d1 * g1 + d2 * g2;
    }
    __syncthreads();
    // ---- assembly: a1 = silu(z1) -> X1 (row-major) and X0 (edge-transposed).  Half-warp per row, float4 per lane.
    {
      const int l16 = lane & 15, hsel = lane >> 4;
      const float4 wq = *reinterpret_cast<const float4*>(v->wq + 4 * l16);
#pragma unroll 1
      for (int i0 = 0; i0 < 16; i0 += 8) {
        float4 p[4], qv[4];
        int ri[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * 16 + i0 + 2 * j + hsel;
          ri[j] = v->srow[rr];
          p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ri[j] >= 0) {
            p[j] = *reinterpret_cast<const float4*>(a.P + (size_t)ri[j] * kH + 4 * l16);
            qv[j] = *reinterpret_cast<const float4*>(a.Q + (size_t)v->scol[rr] * kH + 4 * l16);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * 16 + i0 + 2 * j + hsel;
          const float qi = v->sq[rr];
          float z0 = p[j].x + qv[j].x + qi * wq.x, z1 = p[j].y + qv[j].y + qi * wq.y,
                z2 = p[j].z + qv[j].z + qi * wq.z, z3 = p[j].w + qv[j].w + qi * wq.w;
#pragma unroll
          for (int f = 0; f < kTcMaxFe; ++f) {
            if (f < a.Fe) {
              const float ef = v->sea[rr * kTcMaxFe + f];
              const float4 wf = *reinterpret_cast<const float4*>(v->Wa + f * kH + 4 * l16);
              z0 = fmaf(ef, wf.x, z0); z1 = fmaf(ef, wf.y, z1); z2 = fmaf(ef, wf.z, z2); z3 = fmaf(ef, wf.w, z3);
            }
          }
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f), od = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ri[j] >= 0) {
            tc_silu_grad(z0, o.x, od.x); tc_silu_grad(z1, o.y, od.y);
            tc_silu_grad(z2, o.z, od.z); tc_silu_grad(z3, o.w, od.w);
          }
          *reinterpret_cast<float4*>(X1 + umma::tile_chunk_off(rr, l16, kTM)) = o;
          // silu'(z1) kept as fp16 for the last epilogue: row rr, 8-byte slot l16 of the swizzled 128-byte row
          __half2 h01 = __floats2half2_rn(od.x, od.y), h23 = __floats2half2_rn(od.z, od.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01);
          pk.y = *reinterpret_cast<uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(D1 + rr * 128 + ((((l16 >> 1) ^ (rr & 7)) << 4) | ((l16 & 1) << 3))) = pk;
        }
      }
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (t == 0) {
      umma::fence_after();
      tc_gemm(tmem + 0, sX1, kTM, sW2, kH, 8, idesc128, false);             // G1
      umma::commit(&v->bar[0]);
    }
    // a1^T -> X0 while G1 runs: each thread re-reads its own row segment (read-only on X1) and writes it
    // edge-transposed (conflict-free: lanes are consecutive edges)
    {
      const uint32_t ebase = (uint32_t)((row >> 5) * (kH * 128) + ((row & 3) << 2) + half * 4096);
      const int ech = (row & 31) >> 2;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 q4 = *reinterpret_cast<const float4*>(X1 + umma::tile_chunk_off(row, half * 8 + ch, kTM));
        const float vv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = ch * 4 + k;
          *reinterpret_cast<float*>(X0 + ebase + (j >> 3) * 1024 + (j & 7) * 128 + ((ech ^ (j & 7)) << 4)) = vv[k];
        }
      }
    }
    umma::mbar_wait(&v->bar[0], phase);
    umma::fence_after();
    // ---- epilogue 1: m = silu(z2 + b2), keep silu'(z2) ; m -> X2, m^T -> X3
    float d2[32];
    {
      float m[32];
      umma::tmem_ld32(tlane, m);
#pragma unroll
      for (int j = 0; j < 32; ++j) tc_silu_grad(m[j] + v->b2[half * 32 + j], m[j], d2[j]);
      tc_store_both(X2, X3, row, half, m);
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (t == 0) {
      umma::fence_after();
      tc_gemm(tmem + 64, sX2, kTM, sW3, kH, 8, idesc128, false);            // G2
      umma::commit(&v->bar[1]);
    }
    umma::mbar_wait(&v->bar[1], phase);
    umma::fence_after();
    // ---- epilogue 2: a3, silu'(z3), s = w4.a3 ; g3 = gs w4 silu'(z3) -> X1, g3^T -> X2
    {
      float a3[32], d3[32];
      umma::tmem_ld32(tlane + 64, a3);
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        tc_silu_grad(a3[j] + v->b3[half * 32 + j], a3[j], d3[j]);
        part = fmaf(a3[j], v->w4[half * 32 + j], part);
      }
      if (half == 1) v->spart[row] = part;
      __syncthreads();
      float s = half == 0 ? part + v->spart[row] : 0.f;
      __syncthreads();
      if (half == 0) v->spart[row] = s;
      __syncthreads();
      s = v->spart[row];
      float gs = v->sgs[row];
      if (use_tanh) {
        s = tanhf(s);
        gs *= (1.f - s * s);
      }
      if (half == 0) v->ss[row] = s;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float g = gs * v->w4[half * 32 + j] * d3[j];
        a3[j] *= gs;            // gs * a3 : summand of dw4
        d3[j] = g;              // g3
      }
      tc_store_both(X1, X2, row, half, d3);
      c_b3 += warp_transpose_reduce(d3, lane);
      c_w4 += warp_transpose_reduce(a3, lane);
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (t == 0) {
      umma::fence_after();
      tc_gemm(tmem + 256, sX2, kH, sX3, kH, 16, idesc64, !first_tile);      // W3g: dW3 += g3^T m
      tc_gemm(tmem + 128, sX1, kTM, sW3T, kH, 8, idesc128, false);          // D3 : g3 W3
      umma::commit(&v->bar[2]);
    }
    umma::mbar_wait(&v->bar[2], phase);
    umma::fence_after();
    // ---- epilogue 3: g2 = (gm[row] + g3 W3) * silu'(z2) -> X1, g2^T -> X2
    {
      float g2v[32];
      umma::tmem_ld32(tlane + 128, g2v);
      const int r = v->srow[row];
      if (r >= 0 && a.gm != nullptr) {
        const float4* gmr = reinterpret_cast<const float4*>(a.gm + (size_t)r * kH + half * 32);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const float4 g = gmr[ch];
          g2v[ch * 4] += g.x; g2v[ch * 4 + 1] += g.y; g2v[ch * 4 + 2] += g.z; g2v[ch * 4 + 3] += g.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) g2v[j] = r >= 0 ? g2v[j] * d2[j] : 0.f;
      tc_store_both(X1, X2, row, half, g2v);
      c_b2 += warp_transpose_reduce(g2v, lane);
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (t == 0) {
      umma::fence_after();
      tc_gemm(tmem + 320, sX2, kH, sX0, kH, 16, idesc64, !first_tile);      // W2g: dW2 += g2^T a1
      tc_gemm(tmem + 192, sX1, kTM, sW2T, kH, 8, idesc128, false);          // D2 : g2 W2
      umma::commit(&v->bar[3]);
    }
    umma::mbar_wait(&v->bar[3], phase);
    umma::fence_after();
    phase ^= 1;
    first_tile = false;
    // ---- epilogue 4: gz1 = (g2 W2) * silu'(z1), z1 re-gathered ; gz1 -> X1 ; gQ scatter ; column sums
    {
      float g1v[32];
      umma::tmem_ld32(tlane + 192, g1v);
      const int r = v->srow[row];
      float gq = 0.f;
      if (r >= 0) {
        const int c = v->scol[row];
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {          // 16-byte chunks (8 halves) of this thread's 64-byte d1 segment
          const uint4 pk = *reinterpret_cast<const uint4*>(D1 + row * 128 + (((half * 4 + c16) ^ (row & 7)) << 4));
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 dd = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            const int j = c16 * 8 + k * 2;
            g1v[j] *= dd.x;
            g1v[j + 1] *= dd.y;
            gq = fmaf(g1v[j], v->wq[half * 32 + j], gq);
            gq = fmaf(g1v[j + 1], v->wq[half * 32 + j + 1], gq);
          }
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          atomicAdd(reinterpret_cast<float4*>(a.gQ + (size_t)c * kH + half * 32 + ch * 4),
                    make_float4(g1v[ch * 4], g1v[ch * 4 + 1], g1v[ch * 4 + 2], g1v[ch * 4 + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) g1v[j] = 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<float4*>(X1 + umma::tile_chunk_off(row, half * 8 + ch, kTM)) =
            make_float4(g1v[ch * 4], g1v[ch * 4 + 1], g1v[ch * 4 + 2], g1v[ch * 4 + 3]);
      if (half == 1) v->spart[row] = gq;
      // column sums: dwq += gz1 * q_e ; dWa_f += gz1 * ea_f   (transpose-reduce destroys its argument)
      {
        const float q = v->sq[row];
        float tmp[32];
#pragma unroll
        for (int f = 0; f < kTcMaxFe; ++f) {
          if (f < a.Fe) {
            const float ef = v->sea[row * kTcMaxFe + f];
#pragma unroll
            for (int j = 0; j < 32; ++j) tmp[j] = g1v[j] * ef;
            c_Wa[f] += warp_transpose_reduce(tmp, lane);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) tmp[j] = g1v[j] * q;
        c_wq += warp_transpose_reduce(tmp, lane);
      }
      __syncthreads();
      if (half == 0) v->sgq[row] = gq + v->spart[row];
    }
    __syncthreads();
    // ---- gP: row-segment sums of gz1 (column walk over X1) ; gx: both ends of every edge
    {
      const int col = t & 63, grp = t >> 6;
      const uint32_t cbase = (col >> 5) * (kTM * 128) + ((col & 3) << 2);
      const int cch = (col & 31) >> 2;
      int cur = -1;
      float acc = 0.f;
#pragma unroll 1
      for (int g8 = 0; g8 < 4; ++g8) {
        const uint8_t* gbase = X1 + cbase + (grp * 4 + g8) * 1024;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = v->srow[grp * 32 + g8 * 8 + j];
          const float mv = *reinterpret_cast<const float*>(gbase + j * 128 + ((cch ^ j) << 4));
          if (k != cur) {
            if (cur >= 0) atomicAdd(a.gP + (size_t)cur * kH + col, acc);
            cur = k;
            acc = 0.f;
          }
          acc += k >= 0 ? mv : 0.f;
        }
      }
      if (cur >= 0) atomicAdd(a.gP + (size_t)cur * kH + col, acc);
    }
    if (t < kTM) {
      const int r = v->srow[t], c = v->scol[t];
      const float s = v->ss[t], gq2 = 2.f * v->sgq[t];
      const float inv = norm ? 1.f / v->snrm[t] : 1.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float gd = r == c ? 0.f : s * v->sgte[t * 3 + k] * inv + gq2 * v->sd[t * 3 + k];
        if (r >= 0 && r != c) atomicAdd(a.gx + (size_t)c * 3 + k, -gd);
        bool tail;
        float tot = warp_segsum(r >= 0 ? gd : 0.f, r, lane, tail);
        if (tail && r >= 0) atomicAdd(a.gx + (size_t)r * 3 + k, tot);
      }
    }
  }
  // ---- flush: weight-gradient tiles from TMEM (M = 64 layout: row m in lane (m/16)*32 + m%16), column sums
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  if (!first_tile) {
    float w[32];
    umma::tmem_ld32(tlane + 256, w);
    if (lane < 16 && a.g_W3 != nullptr) {
      const int n = quarter * 16 + lane;
      float* dst = a.g_W3 + (size_t)n * kH + half * 32;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, w[j]);
      }
    }
    umma::tmem_ld32(tlane + 320, w);
    if (lane < 16 && a.g_W2 != nullptr) {
      const int n = quarter * 16 + lane;
      float* dst = a.g_W2 + (size_t)n * kH + half * 32;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, w[j]);
      }
    }
    const int n = half * 32 + lane;
    if (a.g_w4 != nullptr) atomicAdd(a.g_w4 + n, c_w4);
    if (a.g_b3 != nullptr) atomicAdd(a.g_b3 + n, c_b3);
    if (a.g_b2 != nullptr) atomicAdd(a.g_b2 + n, c_b2);
    if (a.g_w1 != nullptr) {
      atomicAdd(a.g_w1 + (size_t)n * a.ld1 + 2 * kH, c_wq);
#pragma unroll
      for (int f = 0; f < kTcMaxFe; ++f)
        if (f < a.Fe) atomicAdd(a.g_w1 + (size_t)n * a.ld1 + 2 * kH + 1 + f, c_Wa[f]);
    }
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

cudaError_t launch_edge_bwd_tc(const EdgeArgs& a, int sms, cudaStream_t st) {
  static DevOnce attr;
  const size_t bytes = EdgeTcBwdSmem::bytes;
  if (!attr.get()) {
    cudaError_t e = cudaFuncSetAttribute(edge_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    attr.set();
  }
  const int ntiles = (a.E + kTM - 1) / kTM;
  if (ntiles == 0) return cudaSuccess;
  const int grid = ntiles < sms ? ntiles : sms;
  edge_bwd_tc_kernel<<<grid, kTcThreads, bytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
