// peak_probe.cu -- micro-benchmarks for the roofs the path's kernels are quoted against:
//   kind 0  tcgen05.mma kind::tf32, cta_group::1, M=128 N=256 K=8 per instruction, operands in shared memory
//   kind 1  tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate), M=128 N=256 K=16 per instruction
//   kind 2  fp32 FMA pipe (FFMA): 8 independent chains per thread
//   kind 3  MUFU pipe: tanh.approx.f32, 8 independent chains per thread (the SiLU of the tensor-core kernels)
// One launch does `iters` instructions per issuing thread; bench.py times it with CUDA events and divides.  These are
// the SAME instruction forms the edge / virtual kernels use (one CTA per SM, one elected issuing thread, SWIZZLE_128B
// K-major operand tiles), so "fraction of this probe" answers "how far is the kernel from what this formulation could
// reach", while MEASURED_PEAKS.json's cuBLAS bf16 figure stays the cross-kernel denominator.
#include "common.cuh"
#include "umma.cuh"

namespace fegnn {
namespace probe {

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) { return umma::make_desc(saddr); }
__device__ __forceinline__ uint32_t idesc(int f16, int M, int N) {
  // kind::tf32: a/b format 2 (tf32) ; kind::f16: a/b format 0 (fp16); D format 1 (fp32) in both
  return (1u << 4) | (f16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int F16>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  if (F16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// shared memory: A tile [128 rows][128 B] = 16 KB, B tile [256 rows][128 B] = 32 KB (one 128-byte K block each:
// 32 tf32 or 64 fp16 columns -> 4 K-steps of 32 bytes per block)
constexpr int kProbeSmem = 16384 + 32768 + 1024 + 64;

template <int F16>
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(int iters, float* __restrict__ sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x;
  // non-trivial operand values (all-zero operands draw less power than real data)
  for (int i = t; i < (16384 + 32768) / 4; i += 128) {
    const uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    if (F16) {
      const __half2 v = __floats2half2_rn(((h & 255) - 128) * (1.f / 256.f), (((h >> 8) & 255) - 128) * (1.f / 256.f));
      reinterpret_cast<uint32_t*>(smem)[i] = *reinterpret_cast<const uint32_t*>(&v);
    } else {
      reinterpret_cast<float*>(smem)[i] = umma::to_tf32(((int)(h & 1023) - 512) * (1.f / 1024.f));
    }
  }
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_fence_init();
  }
  if (t < 32) umma::tmem_alloc<512>(slot);
  umma::fence_smem_to_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem = *slot;
  if (t < 32) {
    if (umma::elect_one()) {
      const uint64_t dA = desc_sw128(umma::smem_u32(smem)), dB = desc_sw128(umma::smem_u32(smem + 16384));
      const uint32_t id = idesc(F16, 128, 256);
#pragma unroll 1
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {      // two accumulators (columns 0-255 / 256-511), 4 K-steps of the 128-byte block
          const uint32_t koff = (uint32_t)((j & 3) * 32) >> 4;
          mma<F16>(tmem + (j >> 2) * 256, dA + koff, dB + koff, id, (i | (j & 3)) ? 1u : 0u);
        }
      }
      umma::commit(bar);
    }
    __syncwarp();
  }
  umma::mbar_wait(bar, 0);
  umma::fence_after();
  float v[32];
  umma::tmem_ld32(tmem + ((uint32_t)((t >> 5) * 32) << 16), v);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  if (s == 123.456f) sink[0] = s;            // keeps the accumulator read alive
  umma::fence_before();
  __syncthreads();
  if (t < 32) umma::tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(256) ffma_probe_kernel(int iters, float* __restrict__ sink) {
  float a[8];
  const float x = 1.0001f + threadIdx.x * 1e-7f, y = 0.9999f;
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = k * 0.1f + blockIdx.x * 1e-3f;
#pragma unroll 1
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) mufu_probe_kernel(int iters, float* __restrict__ sink) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.05f * k + threadIdx.x * 1e-4f;
#pragma unroll 1
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[k]));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace probe

// returns the number of "operations" one launch performs through *ops (flops for kinds 0-2, MUFU results for kind 3)
cudaError_t launch_peak_probe(int kind, int iters, float* sink, double* ops, int sms, cudaStream_t st) {
  iters = (iters + 7) & ~7;
  if (kind == 0 || kind == 1) {
    static DevOnce a0, a1;
    DevOnce& once = kind == 0 ? a0 : a1;
    if (!once.get()) {
      cudaError_t e = kind == 0
          ? cudaFuncSetAttribute(probe::tc_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, probe::kProbeSmem)
          : cudaFuncSetAttribute(probe::tc_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, probe::kProbeSmem);
      if (e != cudaSuccess) return e;
      once.set();
    }
    if (kind == 0) probe::tc_probe_kernel<0><<<sms, 128, probe::kProbeSmem, st>>>(iters, sink);
    else probe::tc_probe_kernel<1><<<sms, 128, probe::kProbeSmem, st>>>(iters, sink);
    ++g_launches;
    *ops = (double)sms * iters * 2.0 * 128 * 256 * (kind == 0 ? 8 : 16);
  } else if (kind == 2) {
    probe::ffma_probe_kernel<<<sms * 8, 256, 0, st>>>(iters, sink); ++g_launches;
    *ops = (double)sms * 8 * 256 * (double)iters * 8 * 2.0;
  } else {
    probe::mufu_probe_kernel<<<sms * 8, 256, 0, st>>>(iters, sink); ++g_launches;
    *ops = (double)sms * 8 * 256 * (double)iters * 8;
  }
  return cudaGetLastError();
}

}  // namespace fegnn
