// adam.cu -- one-launch Adam over the flat parameter / gradient buffers of a FastEGNN model.
//
// Replaces the optimizer step of utils/train.py:168-170 (torch.optim.Adam over 119-135 small tensors, 19 multi-tensor
// launches per step) for models whose parameters live in one flat buffer (fastegnn_b200.optim.FusedAdam).  Arithmetic
// is torch.optim.Adam's (amsgrad=False, maximize=False, L2 weight decay added to the gradient):
//   g' = g + wd p ;  m = b1 m + (1-b1) g' ;  v = b2 v + (1-b2) g'^2 ;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// live[i] == 0 marks elements of parameters that received no gradient (torch skips those parameters entirely).
// The step counter lives on the device (float), so the launch pair is CUDA-graph capturable.
#include "common.cuh"

namespace fegnn {

__global__ void __launch_bounds__(256) adam_kernel(long long n4, float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v,
                                                   const uchar4* __restrict__ live, const float* __restrict__ step,
                                                   float lr, double b1d, double b2d, float eps, float wd) {
  // scalar factors in double, as torch evaluates them on the host (1 - beta^t cancels badly in fp32 for small t)
  __shared__ float sc[2];
  if (threadIdx.x == 0) {
    const double t = (double)*step + 1.0;
    sc[0] = (float)((double)lr / (1.0 - pow(b1d, t)));
    sc[1] = (float)sqrt(1.0 - pow(b2d, t));
  }
  __syncthreads();
  const float step_size = sc[0], c2s = sc[1];
  const float b1 = (float)b1d, b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uchar4 lv = live[i];
    if (!(lv.x | lv.y | lv.z | lv.w)) continue;
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
#define FEGNN_ADAM1(c, l)                                        \
    if (l) {                                                     \
      const float gr = fmaf(wd, pp.c, gg.c);                     \
      mm.c = fmaf(b1, mm.c, omb1 * gr);                          \
      vv.c = fmaf(b2, vv.c, omb2 * gr * gr);                     \
      pp.c -= step_size * (mm.c / (sqrtf(vv.c) / c2s + eps));    \
    }
    FEGNN_ADAM1(x, lv.x) FEGNN_ADAM1(y, lv.y) FEGNN_ADAM1(z, lv.z) FEGNN_ADAM1(w, lv.w)
#undef FEGNN_ADAM1
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}
__global__ void adam_step_inc_kernel(float* step) { *step += 1.f; }

cudaError_t launch_adam(long long n, float* p, const float* g, float* m, float* v, const unsigned char* live, float* step,
                        float lr, double b1, double b2, float eps, float wd, int sms, cudaStream_t st) {
  const long long n4 = n / 4;
  if (n4 > 0) {
    long long blocks = (n4 + 255) / 256;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(n4, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                                  reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                                  reinterpret_cast<const uchar4*>(live), step, lr, b1, b2, eps, wd);
    ++g_launches;
  }
  adam_step_inc_kernel<<<1, 1, 0, st>>>(step); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
