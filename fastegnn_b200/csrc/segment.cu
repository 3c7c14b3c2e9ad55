// segment.cu -- the module-level helpers of models/FastEGNN.py:279-294 (unsorted_segment_sum / unsorted_segment_mean) as
// stand-alone kernels.  The layer itself never calls them (its row-segment sums live inside the fused edge kernels); they
// are exported because the reference file exports them.  HBM-bound: one pass over data [E,K] with red.global.add into
// out [S,K] (L2-resident for the sizes the reference uses), segment ids are the reference's int64.
#include "common.cuh"

namespace fegnn {

__global__ void segment_sum_kernel(long long total, int K, int S, const float* __restrict__ data,
                                   const long long* __restrict__ ids, float* __restrict__ out, float* __restrict__ cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long e = i / K;
    const int k = (int)(i - e * K);
    const long long s = ids[e];
    if (s < 0 || s >= S) continue;                      // torch would raise; out-of-range ids are ignored here
    atomicAdd(out + s * K + k, data[i]);
    if (cnt != nullptr && k == 0) atomicAdd(cnt + s, 1.f);
  }
}
__global__ void segment_div_kernel(long long total, int K, float* __restrict__ out, const float* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = out[i] / fmaxf(cnt[i / K], 1.f);          // count.clamp(min=1), :294
}
// adjoint: gdata[e,k] = g[ids[e],k] (/ max(cnt,1) for the mean)
__global__ void segment_gather_kernel(long long total, int K, int S, const float* __restrict__ g,
                                      const long long* __restrict__ ids, const float* __restrict__ cnt,
                                      float* __restrict__ gdata) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long e = i / K;
    const int k = (int)(i - e * K);
    const long long s = ids[e];
    float v = 0.f;
    if (s >= 0 && s < S) v = cnt != nullptr ? g[s * K + k] / fmaxf(cnt[s], 1.f) : g[s * K + k];
    gdata[i] = v;
  }
}

inline int seg_grid(long long total, int sms) {
  long long b = (total + 255) / 256;
  const long long cap = 8LL * sms;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_segment_reduce(long long E, int K, int S, const float* data, const long long* ids, bool mean,
                                  float* out, float* cnt, int sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)S * K, st);
  if (e != cudaSuccess) return e;
  if (mean) {
    e = cudaMemsetAsync(cnt, 0, sizeof(float) * (size_t)S, st);
    if (e != cudaSuccess) return e;
  }
  const long long total = E * K;
  if (total > 0) {
    segment_sum_kernel<<<seg_grid(total, sms), 256, 0, st>>>(total, K, S, data, ids, out, mean ? cnt : nullptr); ++g_launches;
  }
  if (mean && (long long)S * K > 0) {
    const long long n = (long long)S * K;
    segment_div_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, K, out, cnt); ++g_launches;
  }
  return cudaGetLastError();
}
cudaError_t launch_segment_gather(long long E, int K, int S, const float* g, const long long* ids, const float* cnt,
                                  float* gdata, int sms, cudaStream_t st) {
  const long long total = E * K;
  if (total == 0) return cudaSuccess;
  segment_gather_kernel<<<seg_grid(total, sms), 256, 0, st>>>(total, K, S, g, ids, cnt, gdata); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
