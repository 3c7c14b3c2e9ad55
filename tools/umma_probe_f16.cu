// umma_probe_f16.cu -- standalone probe of tcgen05.mma kind::f16 operand layouts on sm_100a (development tool, not part
// of the product; companion of umma_probe.cu, which settled the kind::tf32 layouts the kernels use today).
//
// Questions it answers in ONE run, for the two-tiles-in-flight edge backward planned in DESIGN.md section 6 item 2
// (fp16 operand tiles, fp32 accumulation):
//   F1  K-major A and B from shared memory, SWIZZLE_128B, K = 16 per instruction                      (sanity)
//   F2  B MN-major (a row-major [n][k] fp16 weight tile read as B with N = k, K = n): which of LBO / SBO is the 64-element
//       MN block stride and which the 8-row K atom stride
//   F3  A and B MN-major with M = 64, K = 128 (the weight-gradient GEMM dW = g^T a over the 128 edges of a tile)
//   F5  A from TENSOR MEMORY with two fp16 per 32-bit column (TS form), B K-major
//   F6  A from tensor memory, B MN-major
//   F7  M = 64 accumulators at lane offset 0 AND 16 of the same columns (two weight-gradient tiles sharing 64 columns)
// Every case prints max |err| against a CPU reference under both candidate accumulator lane maps.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe_f16 tools/umma_probe_f16.cu
// run  : ./umma_probe_f16
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

struct Case {
  int M, N, K;                      // D[M][N] = sum_k A[m][k] B[n][k]
  int a_mn, b_mn;                   // 0 = K-major, 1 = MN-major
  uint32_t a_lbo, a_sbo, a_kstep;   // bytes; kstep = start-address advance per K = 16 instruction
  uint32_t b_lbo, b_sbo, b_kstep;
  int a_tmem;                       // 1 = A operand from tensor memory: lane = row m, 32-bit column c = (A[m][2c], A[m][2c+1])
  int d_lane;                       // lane offset of the accumulator address (0, or 16 for the second M = 64 tile)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;                  // descriptor version (sm_100)
  d |= 2ull << 61;                  // SWIZZLE_128B
  return d;
}

// A region 32 KB, B region 32 KB (fp16 images, 1024-aligned); a_pack = A as packed pairs [128][32] u32 for the TS form
__global__ void __launch_bounds__(128, 1) probe_kernel(Case c, const __half* __restrict__ a_img, const __half* __restrict__ b_img,
                                                       const uint32_t* __restrict__ a_pack, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __half* As = (__half*)smem;
  __half* Bs = (__half*)(smem + 32768);
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
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // clear the accumulator columns of every lane (F7 reads lanes no MMA wrote)
  {
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr + c0), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (c.a_tmem) {
    // thread = row m: packed pairs of A[m][.] -> lane m, columns 128 .. 128 + K/2
    for (int c0 = 0; c0 < c.K / 2; c0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = a_pack[tid * 64 + c0 + j];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    // kind::f16: c_format F32 (1 << 4), a_format = b_format = F16 (0)
    uint32_t idesc = (1u << 4) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                     ((uint32_t)(c.M >> 4) << 24);
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    const uint32_t d_addr = tmem + ((uint32_t)c.d_lane << 16);
    for (int ks = 0; ks < c.K / 16; ++ks) {
      uint64_t da = make_desc(a0 + ks * c.a_kstep, c.a_lbo, c.a_sbo);
      uint64_t db = make_desc(b0 + ks * c.b_kstep, c.b_lbo, c.b_sbo);
      uint32_t acc = ks > 0;
      if (c.a_tmem) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     :: "r"(d_addr), "r"(tmem + 128 + ks * 8), "l"(db), "r"(idesc), "r"(acc));
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d_addr), "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)));
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) d_out[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem));
}

// physical half index of logical (row, col) in a tile of 128-byte rows (64 halves), 8-row atoms of 1024 B, 16-byte chunks
// XOR-swizzled by (row & 7); column block cb (64 columns each) starts at cb * blk_bytes.
static size_t phys(int row, int col, size_t blk_bytes) {
  int cb = col / 64, cc = col % 64;
  size_t byte = (size_t)cb * blk_bytes + (size_t)(row / 8) * 1024 + (size_t)(row % 8) * 128 +
                (size_t)(((cc / 8) ^ (row % 8)) * 16) + (size_t)(cc % 8) * 2;
  return byte / 2;
}

static float rnd_exact(uint32_t& s) {      // small dyadic values: exact in fp16, products and sums exact in fp32
  s = s * 1664525u + 1013904223u;
  return (float)((int)((s >> 20) % 33) - 16) / 8.0f;
}

int main() {
  std::vector<Case> cases;
  std::vector<const char*> names;
  auto add = [&](const char* n, Case c) { cases.push_back(c); names.push_back(n); };
  // F1: A K-major [128 x 64] (one 64-half K block), B K-major [64 x 64]; K = 16 per MMA = 32 B inside the 128 B row
  add("F1 kmajor A,B       M128 N64 K64", Case{128, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 0, 0});
  add("F1 kmajor A,B       M64  N64 K64", Case{64, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 0, 0});
  // F2: B MN-major: W row-major [K = n rows][N = k cols], 64 halves per 128 B row, 8-row atoms; one MMA = 16 K rows = 2048 B
  add("F2 B mn  lbo=blk sbo=1024", Case{128, 64, 64, 0, 1, 16, 1024, 32, 8192, 1024, 2048, 0, 0});
  add("F2 B mn  lbo=1024 sbo=blk", Case{128, 64, 64, 0, 1, 16, 1024, 32, 1024, 8192, 2048, 0, 0});
  // F3: A MN-major [K = 128 e][M = 64], B MN-major [K = 128 e][N = 64]
  add("F3 A,B mn M64 K128 lbo=blk sbo=1024", Case{64, 64, 128, 1, 1, 16384, 1024, 2048, 16384, 1024, 2048, 0, 0});
  add("F3 A,B mn M64 K128 lbo=1024 sbo=blk", Case{64, 64, 128, 1, 1, 1024, 16384, 2048, 1024, 16384, 2048, 0, 0});
  // F5 / F6: A from tensor memory (packed pairs), B K-major / MN-major
  add("F5 A tmem, B kmajor M128 N64 K64", Case{128, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 1, 0});
  add("F6 A tmem, B mn (lbo=blk sbo=1024)", Case{128, 64, 64, 0, 1, 16, 1024, 32, 8192, 1024, 2048, 1, 0});
  add("F6 A tmem, B mn (lbo=1024 sbo=blk)", Case{128, 64, 64, 0, 1, 16, 1024, 32, 1024, 8192, 2048, 1, 0});
  // F7: M = 64 accumulator with the address at lane 16 (does it land in lanes 16-31 / 48-63 / ... of the same columns?)
  add("F7 kmajor M64 D at lane 16", Case{64, 64, 64, 0, 0, 16, 1024, 32, 16, 1024, 32, 0, 16});
  add("F7 A,B mn M64 K128 D at lane 16 (lbo=blk)", Case{64, 64, 128, 1, 1, 16384, 1024, 2048, 16384, 1024, 2048, 0, 16});
  // F8 (added after the first run, not yet executed): N = 72 = one 64-element MN block + 8 columns of the next block
  // (the [dW | bias sums] form of the weight-gradient GEMMs), and N = 8 alone
  add("F8 A,B mn M64 N72 K128 (B two blocks)", Case{64, 72, 128, 1, 1, 16384, 1024, 2048, 16384, 1024, 2048, 0, 0});
  add("F8 A,B mn M64 N72 K128 D at lane 16", Case{64, 72, 128, 1, 1, 16384, 1024, 2048, 16384, 1024, 2048, 0, 16});
  add("F8 A,B mn M64 N8  K128", Case{64, 8, 128, 1, 1, 16384, 1024, 2048, 16384, 1024, 2048, 0, 0});

  __half *a_d, *b_d;
  uint32_t* p_d;
  float* d_d;
  CK(cudaMalloc(&a_d, 32768)); CK(cudaMalloc(&b_d, 32768)); CK(cudaMalloc(&p_d, 128 * 64 * 4)); CK(cudaMalloc(&d_d, 128 * 128 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768 + 1024));
  for (size_t ci = 0; ci < cases.size(); ++ci) {
    Case c = cases[ci];
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), D((size_t)c.M * c.N, 0.f);
    std::vector<__half> Ai(16384, __float2half(0.f)), Bi(16384, __float2half(0.f));
    std::vector<uint32_t> Ap(128 * 64, 0u);
    uint32_t s = 4321u + (uint32_t)ci;
    for (auto& v : A) v = rnd_exact(s);
    for (auto& v : B) v = rnd_exact(s);
    for (int m = 0; m < c.M; ++m)
      for (int k = 0; k < c.K; ++k) {
        size_t p = c.a_mn ? phys(k, m, (size_t)c.K * 128) : phys(m, k, (size_t)c.M * 128);
        Ai[p] = __float2half(A[(size_t)m * c.K + k]);
        __half hv = __float2half(A[(size_t)m * c.K + k]);
        uint16_t bits;
        memcpy(&bits, &hv, 2);
        Ap[(size_t)m * 64 + k / 2] |= (uint32_t)bits << (16 * (k & 1));
      }
    for (int n = 0; n < c.N; ++n)
      for (int k = 0; k < c.K; ++k) {
        size_t p = c.b_mn ? phys(k, n, (size_t)c.K * 128) : phys(n, k, (size_t)c.N * 128);
        Bi[p] = __float2half(B[(size_t)n * c.K + k]);
      }
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float acc = 0.f;
        for (int k = 0; k < c.K; ++k) acc += A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
        D[(size_t)m * c.N + n] = acc;
      }
    CK(cudaMemcpy(a_d, Ai.data(), 32768, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b_d, Bi.data(), 32768, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p_d, Ap.data(), 128 * 64 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_d, 0xff, 128 * 128 * 4));
    probe_kernel<<<1, 128, 2 * 32768 + 1024>>>(c, a_d, b_d, p_d, d_d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", names[ci], cudaGetErrorString(e)); return 1; }
    std::vector<float> out((size_t)128 * c.N);
    CK(cudaMemcpy(out.data(), d_d, out.size() * 4, cudaMemcpyDeviceToHost));
    // lane maps: (1) lane = row; (2) M = 64: lane = (m/16)*32 + m%16 + d_lane
    double e1 = 0, e2 = 0, stray = 0;
    std::vector<char> used(128, 0);
    for (int m = 0; m < c.M; ++m) {
      int l2 = c.M == 64 ? (m / 16) * 32 + m % 16 + c.d_lane : m;
      used[l2] = 1;
      for (int n = 0; n < c.N; ++n) {
        float ref = D[(size_t)m * c.N + n];
        double d1 = fabs((double)out[(size_t)m * c.N + n] - ref), d2 = fabs((double)out[(size_t)l2 * c.N + n] - ref);
        if (!(d1 == d1)) d1 = 1e30;
        if (!(d2 == d2)) d2 = 1e30;
        if (d1 > e1) e1 = d1;
        if (d2 > e2) e2 = d2;
      }
    }
    for (int l = 0; l < 128; ++l)          // lanes the second map leaves alone must still hold the zeros written before the MMA
      if (!used[l])
        for (int n = 0; n < c.N; ++n) stray = fmax(stray, fabs((double)out[(size_t)l * c.N + n]));
    printf("[%-44s] max|err| lane=row: %.3g   lane=(m/16)*32+m%%16+off: %.3g   other lanes max|v|: %.3g\n", names[ci], e1, e2, stray);
  }
  return 0;
}
