// edge_tc_bwd2.cu -- fused real-edge BACKWARD phase, second tcgen05 formulation ("TS + MN-major").
//
// Same contract as edge_bwd_kernel / edge_bwd_tc_kernel (autograd of models/FastEGNN.py:102-108,125-129,156).
// Differences from edge_tc_bwd.cu, all validated with tools/umma_probe.cu on hardware:
//   * the A operand of every [128 x 64] GEMM lives in TENSOR MEMORY (tcgen05.mma with [a_tmem]): the epilogue thread
//     that owns an edge row writes its row with one tcgen05.st -- no K-major activation tiles in shared memory;
//   * the weight-gradient GEMMs (dW = g^T a, K = the 128 edges of the tile) read row-major tiles as MN-major
//     operands (layout SWIZZLE_128B_BASE32B: 128-byte rows, 32-byte chunks XOR (row & 3)) -- no transposed copies;
//   * g W products use the row-major weight tile as an MN-major B operand -- no W^T tiles;
//   * bias-type column sums (db3, db2, dwq, dWa) are one more small GEMM each against an auxiliary
//     [128 x 32] tile whose columns are (1, q_e, ea_e0 ...), instead of shuffle reductions;
//   * silu'(z2) waits in tensor memory instead of registers.
// CG = column groups per edge row: 2 -> 256 threads x 32 columns, 4 -> 512 threads x 16 columns.
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace fegnn {
namespace bwd2 {

// ---- SWIZZLE_128B_BASE32B tiles: [R rows][64 cols] fp32 = two 32-column blocks of R x 128 bytes
__device__ __forceinline__ uint32_t mn_off(int row, int col, int R) {
  return (uint32_t)((col >> 5) * (R * 128) + row * 128 + (((((col & 31) >> 3)) ^ (row & 3)) << 5) + ((col & 7) << 2));
}
// 16-byte chunk c16 (columns 4 c16 .. +3)
__device__ __forceinline__ uint32_t mn_chunk_off(int row, int c16, int R) {
  return (uint32_t)((c16 >> 3) * (R * 128) + row * 128 + (((((c16 & 7) >> 1)) ^ (row & 3)) << 5) + ((c16 & 1) << 4));
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t block_stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((block_stride_bytes >> 4) & 0x3FFF) << 16;   // LBO: next 32-element block along M/N
  d |= (uint64_t)(512 >> 4) << 32;                             // SBO: next 4-row atom along K
  d |= 1ull << 46;
  d |= 1ull << 61;                                             // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// Descriptors are built ONCE per kernel; inside a GEMM only the 14-bit start-address field moves (bytes >> 4, no carry
// out of the field: shared memory is < 256 KB), so an MMA costs one 32-bit add per operand in the issuing thread.
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[128 x 64] = A(tmem, [128 x 64]) * W^T, W K-major SWIZZLE_128B [64 x 64]  (dW = descriptor of the W tile)
__device__ __forceinline__ void gemm_ts_kmajor(uint32_t tmem_d, uint32_t tmem_a, uint64_t dW, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    mma_ts(tmem_d, tmem_a + ks * 8, desc_advance(dW, (ks >> 2) * (kH * 128) + (ks & 3) * 32), idesc, ks > 0);
}
// D[128 x 64] = A(tmem) * W, W row-major BASE32B [64 n][64 k] read as MN-major B (N = k, K = n)
__device__ __forceinline__ void gemm_ts_mn(uint32_t tmem_d, uint32_t tmem_a, uint64_t dW, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) mma_ts(tmem_d, tmem_a + ks * 8, desc_advance(dW, ks * 1024), idesc, ks > 0);
}
// D[64 x N] (+)= G^T B over the 128 rows of the tile; G [128 x 64], B [128 x N] row-major BASE32B (32-column blocks
// kTM*128 bytes apart)
__device__ __forceinline__ void gemm_wgrad(uint32_t tmem_d, uint64_t dG, uint64_t dB, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int ks = 0; ks < 16; ++ks)
    umma::mma_tf32(tmem_d, desc_advance(dG, ks * 1024), desc_advance(dB, ks * 1024), idesc, (ks > 0 || accumulate) ? 1u : 0u);
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) { umma::tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_st<32>(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
         "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
         "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
         "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
         "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
         "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
         "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
         "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
         "r"(__float_as_uint(v[31]))
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
         "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
         "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
         "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
         "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// This thread's CPT columns (group cg) of row `row`  <->  a BASE32B tile.  Rows r and r+4 map a chunk to the same
// banks, so threads with (row & 4) walk each chunk pair in the opposite order: every STS.128 / LDS.128 is conflict-free.
template <int CPT>
__device__ __forceinline__ void mn_store_row(uint8_t* tile, int row, int cg, const float (&v)[CPT]) {
  const bool flip = row & 4;
#pragma unroll
  for (int p = 0; p < CPT / 8; ++p) {
    const float4 c0 = make_float4(v[8 * p], v[8 * p + 1], v[8 * p + 2], v[8 * p + 3]);
    const float4 c1 = make_float4(v[8 * p + 4], v[8 * p + 5], v[8 * p + 6], v[8 * p + 7]);
    const int k0 = cg * (CPT / 4) + 2 * p + (flip ? 1 : 0), k1 = k0 ^ 1;
    *reinterpret_cast<float4*>(tile + mn_chunk_off(row, k0, kTM)) = flip ? c1 : c0;
    *reinterpret_cast<float4*>(tile + mn_chunk_off(row, k1, kTM)) = flip ? c0 : c1;
  }
}
template <int CPT>
__device__ __forceinline__ void mn_load_row(const uint8_t* tile, int row, int cg, float (&v)[CPT]) {
  const bool flip = row & 4;
#pragma unroll
  for (int p = 0; p < CPT / 8; ++p) {
    const int k0 = cg * (CPT / 4) + 2 * p + (flip ? 1 : 0), k1 = k0 ^ 1;
    const float4 a = *reinterpret_cast<const float4*>(tile + mn_chunk_off(row, k0, kTM));
    const float4 b = *reinterpret_cast<const float4*>(tile + mn_chunk_off(row, k1, kTM));
    const float4 c0 = flip ? b : a, c1 = flip ? a : b;
    v[8 * p] = c0.x; v[8 * p + 1] = c0.y; v[8 * p + 2] = c0.z; v[8 * p + 3] = c0.w;
    v[8 * p + 4] = c1.x; v[8 * p + 5] = c1.y; v[8 * p + 6] = c1.z; v[8 * p + 7] = c1.w;
  }
}

// stage W (reference [64][ld]) twice: K-major SWIZZLE_128B (y = a W^T) and row-major BASE32B (g W as MN-major B)
template <int NT>
__device__ __forceinline__ void stage_w_both(uint8_t* wk, uint8_t* wm, const float* __restrict__ g, int ld) {
  for (int i = threadIdx.x; i < kH * 16; i += NT) {
    const int n = i >> 4, c = i & 15;
    const float4 w = *reinterpret_cast<const float4*>(g + (size_t)n * ld + c * 4);
    *reinterpret_cast<float4*>(wk + umma::tile_chunk_off(n, c, kH)) = w;
    *reinterpret_cast<float4*>(wm + mn_chunk_off(n, c, kH)) = w;
  }
}

struct Vec {
  float wq[kH], Wa[kTcMaxFe * kH], b2[kH], b3[kH], w4[kH];
  float cw4[kH];                              // dw4 column sums (shared-memory atomics at kernel end)
  int srow[kTM], scol[kTM];
  float sq[kTM], snrm[kTM], sd[kTM * 3], sgte[kTM * 3], sgs[kTM], sea[kTM * kTcMaxFe];
  float spart[4 * kTM], sgqp[4 * kTM], ss[kTM];
  uint64_t bar[6];
  uint32_t tmem_slot;
};

struct Smem {
  static constexpr int kW = kH * kH * 4;     // 16 KB
  static constexpr int kT = kTM * kH * 4;    // 32 KB
  static constexpr int off_W2k = 0, off_W3k = kW, off_W2m = 2 * kW, off_W3m = 3 * kW;
  // TM | AUX | TA are adjacent 32-column blocks: dW3 and db3 are ONE GEMM against B = [TM | AUX] (N = 96), dW2 and db2
  // one against B = [AUX | TA]
  static constexpr int off_TM = 4 * kW;              // m, later gz1      (MN-major B of dW3; column walk for gP)
  static constexpr int off_AUX = off_TM + kT;        // [128][32]: 1, q, ea0..ea3, 0...
  static constexpr int off_TA = off_AUX + kTM * 128; // a1                (MN-major B of dW2)
  static constexpr int off_TG = off_TA + kT;         // g3, later g2      (MN-major A of dW3 / dW2)
  static constexpr int off_D1 = off_TG + kT;           // silu'(z1) fp16 [128][64], 16-byte chunks ^ (row & 7)
  static constexpr int off_vec = off_D1 + kTM * kH * 2;
  static constexpr size_t bytes = off_vec + sizeof(Vec) + 1024;
};

// tensor-memory columns
// kR3: [dW3 (64) | db3 sums (32)], kR2: [db2 sums (32) | dW2 (64)]
constexpr uint32_t kACC0 = 0, kACC1 = 64, kR3 = 128, kR2 = 224, kOPA = 320, kD2T = 384, kDXZ = 448;
constexpr uint32_t kDW3 = kR3, kDX3 = kR3 + 64, kDX2 = kR2, kDW2 = kR2 + 32;

__device__ __forceinline__ void silu_grad_tc(float z, float& a, float& d) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  const float s = fmaf(0.5f, th, 0.5f);
  a = z * s;
  d = fmaf(a, 1.f - s, s);
}

template <int CG>
__global__ void __launch_bounds__(128 * CG, 1) edge_bwd_tc2_kernel(EdgeArgs a) {
  constexpr int NT = 128 * CG, CPT = kH / CG, NW = NT / 32;
  using SM = Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  Vec* v = reinterpret_cast<Vec*>(smem + SM::off_vec);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int quarter = warp & 3, cg = warp >> 2, row = quarter * 32 + lane, c0 = cg * CPT;
  const bool use_tanh = a.flags & FEGNN_F_TANH, norm = a.flags & FEGNN_F_NORMALIZE;

  stage_w_both<NT>(smem + SM::off_W2k, smem + SM::off_W2m, a.W2, kH);
  stage_w_both<NT>(smem + SM::off_W3k, smem + SM::off_W3m, a.W3, kH);
  for (int i = t; i < kH; i += NT) {
    v->wq[i] = a.w1[(size_t)i * a.ld1 + 2 * kH];
    for (int f = 0; f < a.Fe; ++f) v->Wa[f * kH + i] = a.w1[(size_t)i * a.ld1 + 2 * kH + 1 + f];
    v->b2[i] = a.b2[i];
    v->b3[i] = a.b3[i];
    v->w4[i] = a.w4[i];
    v->cw4[i] = 0.f;
  }
  for (int i = t; i < kTM * 128 / 16; i += NT) reinterpret_cast<float4*>(smem + SM::off_AUX)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) umma::mbar_init(&v->bar[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<512>(&v->tmem_slot);
  umma::fence_smem_to_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem = v->tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + c0;   // this thread's lane, first of its CPT columns
  const uint32_t id_ts_k = idesc_tf32(128, 64, 0, 0), id_ts_mn = idesc_tf32(128, 64, 0, 1);
  const uint32_t id_wg96 = idesc_tf32(64, 96, 1, 1), id_aux = idesc_tf32(64, 32, 1, 1);
  const uint64_t dW2k = umma::make_desc(umma::smem_u32(smem + SM::off_W2k)), dW3k = umma::make_desc(umma::smem_u32(smem + SM::off_W3k));
  const uint64_t dW2m = make_desc_mn(umma::smem_u32(smem + SM::off_W2m), kH * 128);
  const uint64_t dW3m = make_desc_mn(umma::smem_u32(smem + SM::off_W3m), kH * 128);
  const uint64_t dTM = make_desc_mn(umma::smem_u32(smem + SM::off_TM), kTM * 128);     // also [TM | AUX] with N = 96
  const uint64_t dAUX = make_desc_mn(umma::smem_u32(smem + SM::off_AUX), kTM * 128);   // also [AUX | TA] with N = 96
  const uint64_t dTG = make_desc_mn(umma::smem_u32(smem + SM::off_TG), kTM * 128);
  uint8_t* TA = smem + SM::off_TA;
  uint8_t* TM = smem + SM::off_TM;
  uint8_t* TG = smem + SM::off_TG;
  uint8_t* AUX = smem + SM::off_AUX;
  uint8_t* D1 = smem + SM::off_D1;
  uint32_t phase = 0;
  bool first_tile = true;
  float pw4[CPT];          // per-row partial sums of dw4 over this thread's tiles
#pragma unroll
  for (int j = 0; j < CPT; ++j) pw4[j] = 0.f;

  const int ntiles = (a.E + kTM - 1) / kTM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (!first_tile) {                         // the previous tile's last GEMM still reads AUX and TM
      umma::mbar_wait(&v->bar[4], phase ^ 1);
      umma::fence_after();
    }
    umma::fence_before();
    __syncthreads();
    // ---- geometry: one thread per edge
    if (t < kTM) {
      const int e = tile * kTM + t;
      int r = -1, c = 0;
      float d0 = 0, d1 = 0, d2 = 0, q = 0, nrm = 1.f, g0 = 0, g1 = 0, g2 = 0;
      float ea[kTcMaxFe];
#pragma unroll
      for (int f = 0; f < kTcMaxFe; ++f) ea[f] = 0.f;
      if (e < a.E) {
        r = a.row[e];
        c = a.col[e];
        d0 = a.x[(size_t)r * 3 + 0] - a.x[(size_t)c * 3 + 0];
        d1 = a.x[(size_t)r * 3 + 1] - a.x[(size_t)c * 3 + 1];
        d2 = a.x[(size_t)r * 3 + 2] - a.x[(size_t)c * 3 + 2];
        q = d0 * d0 + d1 * d1 + d2 * d2;
        g0 = a.gt[(size_t)r * 3 + 0]; g1 = a.gt[(size_t)r * 3 + 1]; g2 = a.gt[(size_t)r * 3 + 2];
#pragma unroll
        for (int f = 0; f < kTcMaxFe; ++f)
          if (f < a.Fe) ea[f] = a.ea[(size_t)e * a.Fe + f];
      }
#pragma unroll
      for (int f = 0; f < kTcMaxFe; ++f) v->sea[t * kTcMaxFe + f] = ea[f];
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
      // aux columns (1, q, ea0, ea1 | ea2, ea3, 0, 0): logical 32-byte chunk 0 of the row
      uint8_t* arow = AUX + t * 128 + ((t & 3) << 5);
      *reinterpret_cast<float4*>(arow) = make_float4(r >= 0 ? 1.f : 0.f, q, ea[0], ea[1]);
      *reinterpret_cast<float4*>(arow + 16) = make_float4(ea[2], ea[3], 0.f, 0.f);
    }
    __syncthreads();
    // ---- assembly: a1 = silu(z1) -> TA (row-major BASE32B), silu'(z1) -> D1 (fp16).  Half-warp per row, float4 per lane.
    {
      const int l16 = lane & 15, hsel = lane >> 4;
      const float4 wq = *reinterpret_cast<const float4*>(v->wq + 4 * l16);
      constexpr int RPW = kTM / NW;            // rows per warp: 16 (8 warps) or 8 (16 warps)
#pragma unroll 1
      for (int i0 = 0; i0 < RPW; i0 += 8) {
        float4 p[4], qv[4];
        int ri[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * RPW + i0 + 2 * j + hsel;
          ri[j] = v->srow[rr];
          p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ri[j] >= 0 && !(a.exp & 4u)) {
            p[j] = *reinterpret_cast<const float4*>(a.P + (size_t)ri[j] * kH + 4 * l16);
            qv[j] = *reinterpret_cast<const float4*>(a.Q + (size_t)v->scol[rr] * kH + 4 * l16);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = warp * RPW + i0 + 2 * j + hsel;
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
            silu_grad_tc(z0, o.x, od.x); silu_grad_tc(z1, o.y, od.y);
            silu_grad_tc(z2, o.z, od.z); silu_grad_tc(z3, o.w, od.w);
          }
          *reinterpret_cast<float4*>(TA + mn_chunk_off(rr, l16, kTM)) = o;
          __half2 h01 = __floats2half2_rn(od.x, od.y), h23 = __floats2half2_rn(od.z, od.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01);
          pk.y = *reinterpret_cast<uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(D1 + rr * 128 + ((((l16 >> 1) ^ (rr & 7)) << 4) | ((l16 & 1) << 3))) = pk;
        }
      }
    }
    __syncthreads();
    // ---- row owners move a1 into tensor memory (A operand of G1)
    {
      float a1[CPT];
      mn_load_row<CPT>(TA, row, cg, a1);
      tmem_st<CPT>(tlane + kOPA, a1);
      tmem_st_wait();
    }
    umma::fence_smem_to_async();       // TA / AUX: generic-proxy writes -> visible to the tensor core (dW2 / aux GEMMs)
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (elect_one()) {
        gemm_ts_kmajor(tmem + kACC0, tmem + kOPA, dW2k, id_ts_k);           // G1: z2 = a1 W2^T
        umma::commit(&v->bar[0]);
      }
      __syncwarp();
    }
    umma::mbar_wait(&v->bar[0], phase);
    umma::fence_after();
    // ---- epilogue 1: m = silu(z2 + b2) -> OPA (tensor memory) and TM ; silu'(z2) -> D2T
    {
      float m[CPT], d2[CPT];
      tmem_ld<CPT>(tlane + kACC0, m);
#pragma unroll
      for (int j = 0; j < CPT; ++j) silu_grad_tc(m[j] + v->b2[c0 + j], m[j], d2[j]);
      tmem_st<CPT>(tlane + kOPA, m);
      tmem_st<CPT>(tlane + kD2T, d2);
      mn_store_row<CPT>(TM, row, cg, m);
      tmem_st_wait();
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (elect_one()) {
        gemm_ts_kmajor(tmem + kACC1, tmem + kOPA, dW3k, id_ts_k);           // G2: z3 = m W3^T
        umma::commit(&v->bar[1]);
      }
      __syncwarp();
    }
    umma::mbar_wait(&v->bar[1], phase);
    umma::fence_after();
    // ---- epilogue 2: a3, silu'(z3), s = w4 . a3 ; g3 = gs w4 silu'(z3) -> OPA and TG
    {
      float a3[CPT], d3[CPT];
      tmem_ld<CPT>(tlane + kACC1, a3);
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        silu_grad_tc(a3[j] + v->b3[c0 + j], a3[j], d3[j]);
        part = fmaf(a3[j], v->w4[c0 + j], part);
      }
      v->spart[cg * kTM + row] = part;
      __syncthreads();
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < CG; ++g) s += v->spart[g * kTM + row];
      float gs = v->sgs[row];
      if (use_tanh) {
        s = tanhf(s);
        gs *= (1.f - s * s);
      }
      if (cg == 0) v->ss[row] = s;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        pw4[j] = fmaf(gs, a3[j], pw4[j]);
        d3[j] = gs * v->w4[c0 + j] * d3[j];     // g3
      }
      tmem_st<CPT>(tlane + kOPA, d3);
      mn_store_row<CPT>(TG, row, cg, d3);
      tmem_st_wait();
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (elect_one()) {
        gemm_ts_mn(tmem + kACC0, tmem + kOPA, dW3m, id_ts_mn);              // D3 : g3 W3  (epilogue 3 waits for this only)
        umma::commit(&v->bar[2]);
        gemm_wgrad(tmem + kR3, dTG, dTM, id_wg96, !first_tile);             // [dW3 | db3..] += g3^T [m | 1, q, ea]
        umma::commit(&v->bar[5]);                                           // TG / TM may be overwritten after this
      }
      __syncwarp();
    }
    umma::mbar_wait(&v->bar[2], phase);
    umma::fence_after();
    // ---- epilogue 3: g2 = (gm[row] + g3 W3) * silu'(z2) -> OPA and TG
    {
      float g2v[CPT], d2[CPT];
      tmem_ld<CPT>(tlane + kACC0, g2v);
      tmem_ld<CPT>(tlane + kD2T, d2);
      const int r = v->srow[row];
      if (r >= 0 && a.gm != nullptr) {
        const float4* gmr = reinterpret_cast<const float4*>(a.gm + (size_t)r * kH + c0);
#pragma unroll
        for (int ch = 0; ch < CPT / 4; ++ch) {
          const float4 g = gmr[ch];
          g2v[ch * 4] += g.x; g2v[ch * 4 + 1] += g.y; g2v[ch * 4 + 2] += g.z; g2v[ch * 4 + 3] += g.w;
        }
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) g2v[j] = r >= 0 ? g2v[j] * d2[j] : 0.f;
      tmem_st<CPT>(tlane + kOPA, g2v);
      umma::mbar_wait(&v->bar[5], phase);        // dW3 GEMM has finished reading TG (g3) and TM (m)
      mn_store_row<CPT>(TG, row, cg, g2v);
      tmem_st_wait();
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (elect_one()) {
        gemm_ts_mn(tmem + kACC1, tmem + kOPA, dW2m, id_ts_mn);              // D2 : g2 W2
        umma::commit(&v->bar[3]);
        gemm_wgrad(tmem + kR2, dTG, dAUX, id_wg96, !first_tile);            // [db2.. | dW2] += g2^T [1, q, ea | a1]
      }                                                                     // (completion: bar[4] below)
      __syncwarp();
    }
    umma::mbar_wait(&v->bar[3], phase);
    umma::fence_after();
    // ---- epilogue 4: gz1 = (g2 W2) * silu'(z1) -> TM ; gQ scatter ; gq = gz1 . wq
    {
      float g1v[CPT];
      tmem_ld<CPT>(tlane + kACC1, g1v);
      const int r = v->srow[row];
      float gq = 0.f;
      if (r >= 0) {
        const int c = v->scol[row];
#pragma unroll
        for (int c16 = 0; c16 < CPT / 8; ++c16) {     // 16-byte chunks (8 halves) of this thread's d1 segment
          const uint4 pk = *reinterpret_cast<const uint4*>(D1 + row * 128 + (((cg * (CPT / 8) + c16) ^ (row & 7)) << 4));
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 dd = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            const int j = c16 * 8 + k * 2;
            g1v[j] *= dd.x;
            g1v[j + 1] *= dd.y;
            gq = fmaf(g1v[j], v->wq[c0 + j], gq);
            gq = fmaf(g1v[j + 1], v->wq[c0 + j + 1], gq);
          }
        }
        if (!(a.exp & 1u)) {
#pragma unroll
          for (int ch = 0; ch < CPT / 4; ++ch)
            atomicAdd(reinterpret_cast<float4*>(a.gQ + (size_t)c * kH + c0 + ch * 4),
                      make_float4(g1v[ch * 4], g1v[ch * 4 + 1], g1v[ch * 4 + 2], g1v[ch * 4 + 3]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) g1v[j] = 0.f;
      }
      mn_store_row<CPT>(TM, row, cg, g1v);
      v->sgqp[cg * kTM + row] = gq;
    }
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::fence_after();
      if (elect_one()) {
        gemm_wgrad(tmem + kDXZ, dTM, dAUX, id_aux, !first_tile);            // (dwq, dWa) += gz1^T (q, ea)  (columns 1..)
        umma::commit(&v->bar[4]);        // waited at the top of the next tile / before the final flush
      }
      __syncwarp();
    }
    phase ^= 1;
    first_tile = false;
    // ---- gP: row-segment sums of gz1 over TM ; gx: both ends of every edge.  Thread (c4, grp) owns the 16-byte column
    //      quad c4 of RPG4 consecutive rows: one LDS.128 per row, one red.global.add.v4.f32 per run of equal row ids.
    if (a.exp & 2u) continue;
    {
      constexpr int GR4 = NT / 16, RPG4 = kTM / GR4;    // 32 groups x 4 rows (512 threads) / 16 groups x 8 rows (256)
      const int c4 = t & 15, grp = t >> 4;
      int cur = -1;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < RPG4; ++i) {
        const int rr = grp * RPG4 + i;
        const int k = v->srow[rr];
        const float4 mv = *reinterpret_cast<const float4*>(TM + mn_chunk_off(rr, c4, kTM));
        if (k != cur) {
          if (cur >= 0) atomicAdd(reinterpret_cast<float4*>(a.gP + (size_t)cur * kH + c4 * 4), acc);
          cur = k;
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (k >= 0) { acc.x += mv.x; acc.y += mv.y; acc.z += mv.z; acc.w += mv.w; }
      }
      if (cur >= 0) atomicAdd(reinterpret_cast<float4*>(a.gP + (size_t)cur * kH + c4 * 4), acc);
    }
    if (t < kTM) {
      const int r = v->srow[t], c = v->scol[t];
      float gq = 0.f;
#pragma unroll
      for (int g = 0; g < CG; ++g) gq += v->sgqp[g * kTM + t];
      const float s = v->ss[t], gq2 = 2.f * gq;
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
  // ---- flush: wait for the last aux GEMM, then read the weight-gradient tiles (M = 64 layout: row n in lane (n/16)*32 + n%16)
  if (!first_tile) umma::mbar_wait(&v->bar[4], phase ^ 1);
  umma::fence_after();
  // dw4: per-row partials -> shared-memory column sums
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    float s = pw4[j];
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (lane == 0) atomicAdd(&v->cw4[c0 + j], s);
  }
  __syncthreads();
  if (!first_tile) {
    const int n = quarter * 16 + lane;       // valid for lane < 16
    float w[CPT];
    tmem_ld<CPT>(tlane + kDW3, w);
    if (lane < 16 && a.g_W3 != nullptr) {
      float* dst = a.g_W3 + (size_t)n * kH + c0;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4) atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) atomicAdd(dst + j, w[j]);
      }
    }
    tmem_ld<CPT>(tlane + kDW2, w);
    if (lane < 16 && a.g_W2 != nullptr) {
      float* dst = a.g_W2 + (size_t)n * kH + c0;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4) atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) atomicAdd(dst + j, w[j]);
      }
    }
    // aux sums: column 0 of DX3 / DX2, columns 1.. of DXZ  (read by the column-group-0 warps)
    if (cg == 0) {
      float x3[16], x2[16], xz[16];
      const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16);
      tmem_ld<16>(tl + kDX3, x3);
      tmem_ld<16>(tl + kDX2, x2);
      tmem_ld<16>(tl + kDXZ, xz);
      if (lane < 16) {
        if (a.g_b3 != nullptr) atomicAdd(a.g_b3 + n, x3[0]);
        if (a.g_b2 != nullptr) atomicAdd(a.g_b2 + n, x2[0]);
        if (a.g_w1 != nullptr) {
          atomicAdd(a.g_w1 + (size_t)n * a.ld1 + 2 * kH, xz[1]);
#pragma unroll
          for (int f = 0; f < kTcMaxFe; ++f)
            if (f < a.Fe) atomicAdd(a.g_w1 + (size_t)n * a.ld1 + 2 * kH + 1 + f, xz[2 + f]);
        }
      }
    }
    if (t < kH && a.g_w4 != nullptr) atomicAdd(a.g_w4 + t, v->cw4[t]);
  }
  umma::fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

}  // namespace bwd2

template <int CG>
cudaError_t launch_edge_bwd_tc2(const EdgeArgs& a, int sms, cudaStream_t st) {
  static DevOnce attr;
  const size_t bytes = bwd2::Smem::bytes;
  if (!attr.get()) {
    cudaError_t e = cudaFuncSetAttribute(bwd2::edge_bwd_tc2_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    attr.set();
  }
  const int ntiles = (a.E + kTM - 1) / kTM;
  if (ntiles == 0) return cudaSuccess;
  const int grid = ntiles < sms ? ntiles : sms;
  bwd2::edge_bwd_tc2_kernel<CG><<<grid, 128 * CG, bytes, st>>>(a); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
