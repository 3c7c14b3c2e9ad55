// umma.cuh -- tcgen05 / TMEM / mbarrier wrappers (inline PTX, sm_100a) and the shared-memory
// operand layout used by the tensor-core kernels.
//
// Operand layout ("K-major, SWIZZLE_128B", validated on hardware by tools/umma_probe.cu):
//   a tile of R rows x 64 fp32 columns is stored as two K-blocks of 32 columns; a K-block is
//   R rows of 128 bytes; groups of 8 rows form 1024-byte atoms; inside a row the eight 16-byte
//   chunks are XOR-swizzled with (row & 7).  Tile bases must be 1024-byte aligned.
//   Descriptor: start>>4 | LBO(16)>>4 <<16 | SBO(1024)>>4 <<32 | version 1 <<46 | SWIZZLE_128B(2) <<61;
//   one tcgen05.mma kind::tf32 consumes K = 8 columns = 32 bytes: advance the start address by 32 B
//   inside a K-block, by R*128 B to the next K-block.
// Accumulator layout (cta_group::1): M = 128 -> row m in TMEM lane m, column n in TMEM column n;
//   M = 64 -> row m in lane (m/16)*32 + m%16.  Warp w of a CTA may read lanes 32*(w%4) .. +31.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fegnn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of (row, col) inside a [R x 64] fp32 operand tile
__device__ __forceinline__ uint32_t tile_off(int row, int col, int R) {
  return (uint32_t)((col >> 5) * (R * 128) + (row >> 3) * 1024 + (row & 7) * 128 +
                    ((((col & 31) >> 2) ^ (row & 7)) << 4) + ((col & 3) << 2));
}
// byte offset of 16-byte chunk c (columns 4c..4c+3) of a row
__device__ __forceinline__ uint32_t tile_chunk_off(int row, int c, int R) {
  return (uint32_t)((c >> 3) * (R * 128) + (row >> 3) * 1024 + (row & 7) * 128 + (((c & 7) ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // LBO = 16 B (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO = 1024 B between 8-row atoms
  d |= 1ull << 46;                        // descriptor version
  d |= 2ull << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// D[128 x N] (+)= A[128 x 64] * B[N x 64]^T  : 8 MMAs of K = 8.  a_rows / b_rows = rows of the operand tiles.
__device__ __forceinline__ void gemm_k64(uint32_t tmem_d, uint32_t a_saddr, int a_rows, uint32_t b_saddr, int b_rows,
                                         uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t aoff = (ks >> 2) * (a_rows * 128) + (ks & 3) * 32;
    const uint32_t boff = (ks >> 2) * (b_rows * 128) + (ks & 3) * 32;
    mma_tf32(tmem_d, make_desc(a_saddr + aoff), make_desc(b_saddr + boff), idesc, (ks > 0 || accumulate_first) ? 1u : 0u);
  }
}

// Same GEMM from a descriptor built once per kernel: per MMA only the 14-bit start-address field moves (bytes >> 4;
// no carry out of the field, shared memory is < 256 KB) -- one add per operand in the issuing thread.
__device__ __forceinline__ void gemm_k64_desc(uint32_t tmem_d, uint64_t da, int a_rows, uint64_t db, int b_rows,
                                              uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t aoff = (ks >> 2) * (a_rows * 128) + (ks & 3) * 32;
    const uint32_t boff = (ks >> 2) * (b_rows * 128) + (ks & 3) * 32;
    mma_tf32(tmem_d, da + (uint64_t)(aoff >> 4), db + (uint64_t)(boff >> 4), idesc, (ks > 0 || accumulate_first) ? 1u : 0u);
  }
}
// one lane of a converged warp (the form under which the compiler keeps tcgen05.mma operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// round-to-nearest tf32 (kept in a 32-bit container)
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace umma
}  // namespace fegnn
