// node_tc.cu -- per-node dense phases on the tcgen05 tensor cores (single-pass TF32 tiles).
//
// node_pre forward (models/FastEGNN.py:104,115,139,142,162, the h-dependent half of every first Linear):
//   one 128-node tile of h (K-major SWIZZLE_128B in shared memory, coalesced loads) times up to three stacked
//   64x64 weight blocks = ONE GEMM with N = 64 * nb, accumulators in tensor memory.  Work item = (tile, group):
//   group 0 = {P, Q, Av}, group 1 = {Uh (unless last layer), phi_v head, phi_g head (gravity)}; 2 CTAs / SM.
// The fp32 FMA kernels (node_kernels.cu) remain mode 0.
#include "common.cuh"
#include "umma.cuh"

namespace fegnn {
namespace ntc {

using bwd2::desc_advance;
using bwd2::tmem_ld;

__device__ __forceinline__ float silu_tc(float z) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  return z * fmaf(0.5f, th, 0.5f);
}

// stage a [64][64] block (element (n,k) = g[n*ld + off + k]) K-major SWIZZLE_128B into rows [row0, row0+64) of an R-row tile
template <int NT>
__device__ __forceinline__ void stage_block_kmajor(uint8_t* dst, const float* __restrict__ g, int ld, int off, int row0, int R) {
  if (((ld | off) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int i = threadIdx.x; i < kH * 16; i += NT) {
      const int n = i >> 4, c = i & 15;
      *reinterpret_cast<float4*>(dst + umma::tile_chunk_off(row0 + n, c, R)) =
          *reinterpret_cast<const float4*>(g + (size_t)n * ld + off + c * 4);
    }
  } else {
    float w[kH * kH / NT];
#pragma unroll
    for (int j = 0; j < kH * kH / NT; ++j) {
      const int i = threadIdx.x + j * NT;
      w[j] = g[(size_t)(i >> 6) * ld + off + (i & 63)];
    }
#pragma unroll
    for (int j = 0; j < kH * kH / NT; ++j) {
      const int i = threadIdx.x + j * NT;
      *reinterpret_cast<float*>(dst + umma::tile_off(row0 + (i >> 6), i & 63, R)) = w[j];
    }
  }
}

struct PreVec {
  float bias[3 * kH];      // first-layer bias of each block (0 where the reference has none)
  float w2[3 * kH];        // output weights of head blocks
  float sp[2 * kTM];       // cross-column-group partial sums of a head
  uint64_t bar;
  uint32_t tmem_slot;
};
struct PreSmem {
  static constexpr int off_W = 0;                    // up to [192][64] K-major
  static constexpr int off_A = 3 * 16384;            // h tile [128][64] K-major
  static constexpr int off_vec = off_A + 32768;
  static constexpr size_t bytes = off_vec + sizeof(PreVec) + 1024;
};

__global__ void __launch_bounds__(256, 2) node_pre_fwd_tc_kernel(NodePreArgs a) {
  constexpr int NT = 256, CPT = 32;
  using SM = PreSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  PreVec* v = reinterpret_cast<PreVec*>(smem + SM::off_vec);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int quarter = warp & 3, cg = warp >> 2, row = quarter * 32 + lane, c0 = cg * CPT;
  const bool grav = a.flags & FEGNN_F_GRAVITY, last = a.flags & FEGNN_F_LAST;
  const int group = blockIdx.x & 1, cta = blockIdx.x >> 1, nctas = gridDim.x >> 1;
  // block ids: 0 P, 1 Q, 2 Av, 3 Uh, 4 phi_v head, 5 phi_g head
  int blk[3], nb;
  if (group == 0) { blk[0] = 0; blk[1] = 1; blk[2] = 2; nb = 3; }
  else {
    nb = 0;
    if (!last) blk[nb++] = 3;
    blk[nb++] = 4;
    if (grav) blk[nb++] = 5;
    for (int j = nb; j < 3; ++j) blk[j] = -1;
  }
  const int R = nb * kH;
  for (int j = 0; j < nb; ++j) {
    const int b = blk[j];
    const float* wsrc = b <= 1 ? a.edge_w0 : b == 2 ? a.edgev_w0 : b == 3 ? a.node_w0 : b == 4 ? a.vel_w0 : a.grav_w0;
    const int ld = b <= 1 ? a.ld1 : b == 2 ? a.ldv : b == 3 ? a.ldn : kH;
    stage_block_kmajor<NT>(smem + SM::off_W, wsrc, ld, b == 1 ? kH : 0, j * kH, R);
    const float* bias = b == 0 ? a.edge_b0 : b == 2 ? a.edgev_b0 : b == 3 ? a.node_b0 : b == 4 ? a.vel_b0 : b == 5 ? a.grav_b0 : nullptr;
    const float* w2 = b == 4 ? a.vel_w2 : b == 5 ? a.grav_w2 : nullptr;
    for (int i = t; i < kH; i += NT) {
      v->bias[j * kH + i] = bias != nullptr ? bias[i] : 0.f;
      v->w2[j * kH + i] = w2 != nullptr ? w2[i] : 0.f;
    }
  }
  if (t == 0) {
    umma::mbar_init(&v->bar, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<256>(&v->tmem_slot);
  umma::fence_smem_to_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem = v->tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + c0;
  const uint32_t idesc = umma::make_idesc_tf32(128, R);
  const uint64_t dW = umma::make_desc(umma::smem_u32(smem + SM::off_W));
  const uint64_t dA = umma::make_desc(umma::smem_u32(smem + SM::off_A));
  uint8_t* A = smem + SM::off_A;
  uint32_t phase = 0;

  const int ntiles = (a.N + kTM - 1) / kTM;
  for (int tile = cta; tile < ntiles; tile += nctas) {
    const int i0 = tile * kTM, nvalid = min(kTM, a.N - i0);
    umma::fence_before();
    __syncthreads();
    // ---- h tile -> shared memory (half-warp per row, one 16-byte chunk per lane)
    {
      const int l16 = lane & 15, hsel = lane >> 4;
#pragma unroll 4
      for (int i = 0; i < kTM / 8 / 2; ++i) {
        const int rr = warp * (kTM / 8) + 2 * i + hsel;
        float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < nvalid) hv = *reinterpret_cast<const float4*>(a.h + (size_t)(i0 + rr) * kH + 4 * l16);
        *reinterpret_cast<float4*>(A + umma::tile_chunk_off(rr, l16, kTM)) = hv;
      }
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (umma::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma::mma_tf32(tmem, desc_advance(dA, (ks >> 2) * (kTM * 128) + (ks & 3) * 32),
                         desc_advance(dW, (ks >> 2) * (R * 128) + (ks & 3) * 32), idesc, ks > 0);
        umma::commit(&v->bar);
      }
      __syncwarp();
    }
    umma::mbar_wait(&v->bar, phase);
    umma::fence_after();
    phase ^= 1;
    // ---- epilogue: bias + store (P, Q, Av, Uh) or head output w2 . silu(z + b0) + b2 (sv, sg)
    for (int j = 0; j < nb; ++j) {
      const int b = blk[j];
      float z[CPT];
      tmem_ld<CPT>(tlane + j * kH, z);
      if (b < 4) {
        float* out = b == 0 ? a.P : b == 1 ? a.Q : b == 2 ? a.Av : a.Uh;
        if (row < nvalid) {
          float4* dst = reinterpret_cast<float4*>(out + (size_t)(i0 + row) * kH + c0);
#pragma unroll
          for (int ch = 0; ch < CPT / 4; ++ch)
            dst[ch] = make_float4(z[ch * 4] + v->bias[j * kH + c0 + ch * 4], z[ch * 4 + 1] + v->bias[j * kH + c0 + ch * 4 + 1],
                                  z[ch * 4 + 2] + v->bias[j * kH + c0 + ch * 4 + 2], z[ch * 4 + 3] + v->bias[j * kH + c0 + ch * 4 + 3]);
        }
      } else {
        float s = 0.f;
#pragma unroll
        for (int jj = 0; jj < CPT; ++jj) s = fmaf(silu_tc(z[jj] + v->bias[j * kH + c0 + jj]), v->w2[j * kH + c0 + jj], s);
        v->sp[cg * kTM + row] = s;
        __syncthreads();
        if (cg == 0 && row < nvalid) {
          const float b2 = b == 4 ? a.vel_b2[0] : a.grav_b2[0];
          (b == 4 ? a.sv : a.sg)[i0 + row] = s + v->sp[kTM + row] + b2;
        }
        __syncthreads();
      }
    }
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

}  // namespace ntc

cudaError_t launch_node_pre_fwd_tc(const NodePreArgs& a, int sms, cudaStream_t st) {
  static DevOnce attr;
  if (!attr.get()) {
    cudaError_t e = cudaFuncSetAttribute(ntc::node_pre_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ntc::PreSmem::bytes);
    if (e != cudaSuccess) return e;
    attr.set();
  }
  const int ntiles = (a.N + kTM - 1) / kTM;
  if (ntiles == 0) return cudaSuccess;
  int per = ntiles < sms ? ntiles : sms;          // CTAs per group; 2 groups -> up to 2 CTAs / SM
  ntc::node_pre_fwd_tc_kernel<<<2 * per, 256, ntc::PreSmem::bytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
