// umma_probe.cu -- standalone probe of tcgen05.mma kind::tf32 operand layouts on sm_100a.
// Development tool (not part of the product): checks shared-memory descriptor / instruction
// descriptor hypotheses against a CPU reference before they are used in the kernels.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
// run  : ./umma_probe            (runs a fixed list of cases, prints max |err| per case)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

struct Case {
  int M, N, K;          // MMA tile: D[M][N] = sum_k A[m][k] B[n][k]
  int a_mn, b_mn;       // 0 = K-major operand, 1 = MN-major
  uint32_t a_lbo, a_sbo, a_kstep;   // bytes; a_kstep = start-address advance per K=8 step
  uint32_t b_lbo, b_sbo, b_kstep;
  uint32_t a_kblk, b_kblk;          // extra advance every 4 k-steps for K-major (next 128B K-block); 0 = fold into kstep
  uint32_t a_lt = 2, b_lt = 2;
  int a_tmem = 0;                   // 1 = A operand from TMEM (lane = row m, 32-bit column = k), written with tcgen05.st      // descriptor layout type: 2 = SWIZZLE_128B (16-byte chunks ^ row&7), 1 = SWIZZLE_128B_BASE32B (32-byte chunks ^ row&3)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;        // descriptor version (sm_100)
  d |= (uint64_t)lt << 61;  // layout type
  return d;
}

// A region: 64 KB, B region: 64 KB, both 1024-aligned
__global__ void __launch_bounds__(128, 1) probe_kernel(Case c, const float* __restrict__ a_img, const float* __restrict__ b_img,
                                                       float* __restrict__ d_out /*[128 lanes][N]*/) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* As = (float*)smem;
  float* Bs = (float*)(smem + 65536);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) { As[i] = a_img[i]; Bs[i] = b_img[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");     // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (c.a_tmem) {
    // thread = row m: A[m][0..K) -> TMEM lane m, columns 128 .. 128+K  (a_img is plain row-major [128][64] here)
    for (int c0 = 0; c0 < c.K; c0 += 32) {
      uint32_t r[32];
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(a_img[tid * 64 + c0 + j]);
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                   "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                     ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    for (int ks = 0; ks < c.K / 8; ++ks) {
      uint32_t aoff = c.a_kblk ? (ks / 4) * c.a_kblk + (ks % 4) * c.a_kstep : ks * c.a_kstep;
      uint32_t boff = c.b_kblk ? (ks / 4) * c.b_kblk + (ks % 4) * c.b_kstep : ks * c.b_kstep;
      uint64_t da = make_desc(a0 + aoff, c.a_lbo, c.a_sbo, c.a_lt);
      uint64_t db = make_desc(b0 + boff, c.b_lbo, c.b_sbo, c.b_lt);
      uint32_t acc = ks > 0;
      if (c.a_tmem) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(tmem), "r"(tmem + 128 + ks * 8), "l"(db), "r"(idesc), "r"(acc));
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)));
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // every thread reads its TMEM lane, all N columns (in chunks of 32)
  for (int c0 = 0; c0 < c.N; c0 += 32) {
    uint32_t r[32];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) d_out[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem));
}

// physical float index of logical (row, col) in a tile made of 128-byte-row column blocks:
// rows of 32 floats, 8-row atoms of 1024 B, 16-byte chunks XOR-swizzled by (row & 7); column block cb
// (32 columns each) starts at cb * blk_bytes.
static size_t phys(int row, int col, size_t blk_bytes, int lt = 2) {
  int cb = col / 32, cc = col % 32;
  if (lt == 1) {   // 128-byte rows, 32-byte chunks XOR-swizzled by (row & 3)
    size_t byte = (size_t)cb * blk_bytes + (size_t)row * 128 + (size_t)(((cc / 8) ^ (row % 4)) * 32) + (size_t)(cc % 8) * 4;
    return byte / 4;
  }
  size_t byte = (size_t)cb * blk_bytes + (size_t)(row / 8) * 1024 + (size_t)(row % 8) * 128 +
                (size_t)(((cc / 4) ^ (row % 8)) * 16) + (size_t)(cc % 4) * 4;
  return byte / 4;
}

static float rnd_tf32_exact(uint32_t& s) {   // small dyadic values: exactly representable in tf32, products exact in fp32
  s = s * 1664525u + 1013904223u;
  return (float)((int)((s >> 20) % 33) - 16) / 8.0f;
}

int main() {
  std::vector<Case> cases;
  std::vector<const char*> names;
  auto add = [&](const char* n, Case c) { cases.push_back(c); names.push_back(n); };
  // T1: A K-major [128 x 64], B K-major [64 x 64]; K-block stride: A 128 rows * 128 B = 16384, B 64 rows * 128 B = 8192
  add("kmajor A,B  M128 N64 K64 (lbo16 sbo1024)", Case{128, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 16384, 8192});
  add("kmajor A,B  M128 N64 K64 (lbo0 sbo1024)", Case{128, 64, 64, 0, 0, 0, 1024, 32, 0, 1024, 32, 16384, 8192});
  // T2: dgrad-like: A K-major [128 x 64(n)], B MN-major: W stored [K=n rows][N=k cols]; column blocks of 32 at 8192 (64 rows*128B)
  add("B mn-major hypA (lbo=S_mn sbo=S_k)", Case{128, 64, 64, 0, 1, 16, 1024, 32, 8192, 1024, 1024, 16384, 0});
  add("B mn-major hypB (lbo=S_k sbo=S_mn)", Case{128, 64, 64, 0, 1, 16, 1024, 32, 1024, 8192, 1024, 16384, 0});
  // T3: wgrad-like: A MN-major [K=128 e][M=64 n] (col blocks at 16384), B MN-major [K=128 e][N=64 k]; M=64
  add("A,B mn-major M64 K128 hypA", Case{64, 64, 128, 1, 1, 16384, 1024, 1024, 16384, 1024, 1024, 0, 0});
  add("A,B mn-major M64 K128 hypB", Case{64, 64, 128, 1, 1, 1024, 16384, 1024, 1024, 16384, 1024, 0, 0});
  // T4: same with M=128: A MN-major [K=128 e][M=128] needs 4 column blocks (A region = 64 KB: 4 x 16 KB)
  add("A,B mn-major M128 K128 hypA", Case{128, 64, 128, 1, 1, 16384, 1024, 1024, 16384, 1024, 1024, 0, 0});
  add("A,B mn-major M128 K128 hypB", Case{128, 64, 128, 1, 1, 1024, 16384, 1024, 1024, 16384, 1024, 0, 0});
  // T6: MN-major tf32 operands need SWIZZLE_128B_BASE32B (layout type 1): 128-byte rows along MN, 4-row K atoms of 512 B
  add("B mn32 dgrad  lbo=blk sbo=512", Case{128, 64, 64, 0, 1, 16, 1024, 32, 8192, 512, 1024, 16384, 0, 2, 1});
  add("B mn32 dgrad  lbo=512 sbo=blk", Case{128, 64, 64, 0, 1, 16, 1024, 32, 512, 8192, 1024, 16384, 0, 2, 1});
  add("A,B mn32 M64 K128 lbo=blk sbo=512", Case{64, 64, 128, 1, 1, 16384, 512, 1024, 16384, 512, 1024, 0, 0, 1, 1});
  add("A,B mn32 M64 K128 lbo=512 sbo=blk", Case{64, 64, 128, 1, 1, 512, 16384, 1024, 512, 16384, 1024, 0, 0, 1, 1});
  add("A,B mn32 M128 K128 lbo=blk sbo=512", Case{128, 64, 128, 1, 1, 16384, 512, 1024, 16384, 512, 1024, 0, 0, 1, 1});
  // T7: A operand from TMEM (TS form), B K-major from smem / B MN-major BASE32B from smem
  add("A tmem, B kmajor   M128 N64 K64", Case{128, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 16384, 8192, 2, 2, 1});
  add("A tmem, B mn32     M128 N64 K64", Case{128, 64, 64, 0, 1, 16, 1024, 32, 8192, 512, 1024, 16384, 0, 2, 1, 1});
  // T5: K-major A with M=64 (TMEM lane layout probe)
  add("kmajor A,B  M64 N64 K64", Case{64, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 8192, 8192});

  float *a_d, *b_d, *d_d;
  CK(cudaMalloc(&a_d, 65536)); CK(cudaMalloc(&b_d, 65536)); CK(cudaMalloc(&d_d, 128 * 64 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024));
  for (size_t ci = 0; ci < cases.size(); ++ci) {
    Case c = cases[ci];
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), Ai(16384, 0.f), Bi(16384, 0.f), D((size_t)c.M * c.N, 0.f);
    uint32_t s = 1234u + (uint32_t)ci;
    for (auto& v : A) v = rnd_tf32_exact(s);
    for (auto& v : B) v = rnd_tf32_exact(s);
    // fill images
    for (int m = 0; m < c.M; ++m)
      for (int k = 0; k < c.K; ++k) {
        size_t p = c.a_tmem ? (size_t)m * 64 + k : c.a_mn ? phys(k, m, (size_t)c.K * 128, c.a_lt) : phys(m, k, (size_t)c.M * 128, c.a_lt);
        Ai[p] = A[(size_t)m * c.K + k];
      }
    for (int n = 0; n < c.N; ++n)
      for (int k = 0; k < c.K; ++k) {
        size_t p = c.b_mn ? phys(k, n, (size_t)c.K * 128, c.b_lt) : phys(n, k, (size_t)c.N * 128, c.b_lt);
        Bi[p] = B[(size_t)n * c.K + k];
      }
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float acc = 0.f;
        for (int k = 0; k < c.K; ++k) acc += A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
        D[(size_t)m * c.N + n] = acc;
      }
    CK(cudaMemcpy(a_d, Ai.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b_d, Bi.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_d, 0xff, 128 * 64 * 4));
    probe_kernel<<<1, 128, 2 * 65536 + 1024>>>(c, a_d, b_d, d_d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", names[ci], cudaGetErrorString(e)); return 1; }
    std::vector<float> out((size_t)128 * c.N);
    CK(cudaMemcpy(out.data(), d_d, out.size() * 4, cudaMemcpyDeviceToHost));
    // hypothesis 1: row m <-> lane m.  hypothesis 2 (M=64): row m <-> lane (m/16)*32 + m%16
    double e1 = 0, e2 = 0;
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float ref = D[(size_t)m * c.N + n];
        float g1 = out[(size_t)m * c.N + n];
        int l2 = c.M == 64 ? (m / 16) * 32 + m % 16 : m;
        float g2 = out[(size_t)l2 * c.N + n];
        double d1 = fabs((double)g1 - ref), d2 = fabs((double)g2 - ref);
        if (!(d1 == d1)) d1 = 1e30;
        if (!(d2 == d2)) d2 = 1e30;
        if (d1 > e1) e1 = d1;
        if (d2 > e2) e2 = d2;
      }
    printf("[%-45s] max|err| lane=row: %.3g   lane=(m/16)*32+m%%16: %.3g   (ref[0][0..2]=%g %g %g got %g %g %g)\n", names[ci], e1, e2,
           D[0], D[1], D[2], out[0], out[1], out[2]);
  }
  return 0;
}
