// mmd.cu -- MMD regulariser between sampled real coordinates and virtual coordinates.
//
// Replaces utils/train.py:17-20 (Laplacian kernel on the plain L2 distance) and the
// per-graph Python loop / boolean masking of :111-165.  The random sample is an input
// (global node indices), so torch.randperm stays the caller's RNG.
//   loss = 1/(B C^2) sum_b sum_{c,c'} k(Z_bc, Z_bc') - 2/(B ns C) sum_b sum_{s,c} k(x_s, Z_bc)
//   k(p,q) = exp(-|p-q| / (2 sigma^2)),  d|p-q|/dp = 0 at p == q (torch.cdist convention).
#include "common.cuh"

namespace fegnn {

__global__ void __launch_bounds__(128) mmd_fwd_kernel(int B, int C, int ns, float inv2s2, float svv, float srv, const float* __restrict__ x,
                                                      const float* __restrict__ Z, const int* __restrict__ idx,
                                                      float* __restrict__ loss) {
  const int b = blockIdx.x;
  const float cvv = svv / ((float)B * C * C), crv = ns > 0 ? srv * 2.f / ((float)B * ns * C) : 0.f;
  float acc = 0.f;
  for (int p = threadIdx.x; p < C * C + ns * C; p += blockDim.x) {
    float px, py, pz, w;
    int c;
    if (p < C * C) {
      const int c0 = p / C;
      c = p - c0 * C;
      px = Z[((size_t)b * 3 + 0) * C + c0]; py = Z[((size_t)b * 3 + 1) * C + c0]; pz = Z[((size_t)b * 3 + 2) * C + c0];
      w = cvv;
    } else {
      const int q = p - C * C, s = q / C;
      c = q - s * C;
      const int i = idx[(size_t)b * ns + s];
      px = x[(size_t)i * 3 + 0]; py = x[(size_t)i * 3 + 1]; pz = x[(size_t)i * 3 + 2];
      w = -crv;
    }
    const float dx = px - Z[((size_t)b * 3 + 0) * C + c], dy = py - Z[((size_t)b * 3 + 1) * C + c],
                dz = pz - Z[((size_t)b * 3 + 2) * C + c];
    acc += w * __expf(-sqrtf(dx * dx + dy * dy + dz * dz) * inv2s2);
  }
  acc = warpsum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc);
}

__global__ void __launch_bounds__(128) mmd_bwd_kernel(int B, int C, int ns, float inv2s2, float svv, float srv, const float* __restrict__ x,
                                                      const float* __restrict__ Z, const int* __restrict__ idx,
                                                      const float* __restrict__ gloss, float* __restrict__ gx,
                                                      float* __restrict__ gZ) {
  __shared__ float gz[3 * FEGNN_MAX_C];
  const int b = blockIdx.x;
  const float gl = gloss[0];
  const float cvv = svv * gl / ((float)B * C * C), crv = ns > 0 ? srv * 2.f * gl / ((float)B * ns * C) : 0.f;
  if (threadIdx.x < 3 * C) gz[threadIdx.x] = 0.f;
  __syncthreads();
  for (int p = threadIdx.x; p < C * C + ns * C; p += blockDim.x) {
    float px, py, pz, w;
    int c, c0 = -1, i = -1;
    if (p < C * C) {
      c0 = p / C;
      c = p - c0 * C;
      px = Z[((size_t)b * 3 + 0) * C + c0]; py = Z[((size_t)b * 3 + 1) * C + c0]; pz = Z[((size_t)b * 3 + 2) * C + c0];
      w = cvv;
    } else {
      const int q = p - C * C, s = q / C;
      c = q - s * C;
      i = idx[(size_t)b * ns + s];
      px = x[(size_t)i * 3 + 0]; py = x[(size_t)i * 3 + 1]; pz = x[(size_t)i * 3 + 2];
      w = -crv;
    }
    const float dx = px - Z[((size_t)b * 3 + 0) * C + c], dy = py - Z[((size_t)b * 3 + 1) * C + c],
                dz = pz - Z[((size_t)b * 3 + 2) * C + c];
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    if (dist > 0.f) {
      // d/dp of w*exp(-dist*inv2s2) = -w*inv2s2*k * (p-q)/dist ; d/dq is the negative
      const float f = -w * inv2s2 * __expf(-dist * inv2s2) / dist;
      const float fx = f * dx, fy = f * dy, fz = f * dz;
      atomicAdd(&gz[0 * C + c], -fx); atomicAdd(&gz[1 * C + c], -fy); atomicAdd(&gz[2 * C + c], -fz);
      if (c0 >= 0) {
        atomicAdd(&gz[0 * C + c0], fx); atomicAdd(&gz[1 * C + c0], fy); atomicAdd(&gz[2 * C + c0], fz);
      } else {
        atomicAdd(gx + (size_t)i * 3 + 0, fx); atomicAdd(gx + (size_t)i * 3 + 1, fy);
        atomicAdd(gx + (size_t)i * 3 + 2, fz);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * C) gZ[(size_t)b * 3 * C + threadIdx.x] = gz[threadIdx.x];
}

// ------------------------------------------------------------------ the step's loss in one kernel per direction
// utils/train.py:104,163: loss = MSE(x, target) + weight * (l_vv - l_rv).  As torch ops around mmd_loss this was ten small
// kernels between the last forward kernel and the first backward kernel (reduce, mul, add, their backward, the fills), none
// overlapping another.  Here: CTAs [0, B) evaluate the MMD term of one graph each, the remaining CTAs sum (x - target)^2.
//   out[0] = total, out[1] = the MSE term alone (what the reference logs at :107), both zeroed by the launcher.
__global__ void __launch_bounds__(128) mse_mmd_fwd_kernel(int N, int B, int C, int ns, float inv2s2, float weight, float svv,
                                                          float srv, float inv_count, const float* __restrict__ x,
                                                          const float* __restrict__ target, const float* __restrict__ Z,
                                                          const int* __restrict__ idx, float* __restrict__ out_total,
                                                          float* __restrict__ out_mse) {
  if ((int)blockIdx.x < B) {
    const int b = blockIdx.x;
    const float cvv = svv / ((float)B * C * C), crv = ns > 0 ? srv * 2.f / ((float)B * ns * C) : 0.f;
    float acc = 0.f;
    for (int p = threadIdx.x; p < C * C + ns * C; p += blockDim.x) {
      float px, py, pz, w;
      int c;
      if (p < C * C) {
        const int c0 = p / C;
        c = p - c0 * C;
        px = Z[((size_t)b * 3 + 0) * C + c0]; py = Z[((size_t)b * 3 + 1) * C + c0]; pz = Z[((size_t)b * 3 + 2) * C + c0];
        w = cvv;
      } else {
        const int q = p - C * C, s = q / C;
        c = q - s * C;
        const int i = idx[(size_t)b * ns + s];
        px = x[(size_t)i * 3 + 0]; py = x[(size_t)i * 3 + 1]; pz = x[(size_t)i * 3 + 2];
        w = -crv;
      }
      const float dx = px - Z[((size_t)b * 3 + 0) * C + c], dy = py - Z[((size_t)b * 3 + 1) * C + c],
                  dz = pz - Z[((size_t)b * 3 + 2) * C + c];
      acc += w * __expf(-sqrtf(dx * dx + dy * dy + dz * dz) * inv2s2);
    }
    acc = warpsum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out_total, weight * acc);
    return;
  }
  const size_t n = (size_t)N * 3, stride = (size_t)(gridDim.x - B) * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)(blockIdx.x - B) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = x[i] - target[i];
    acc = fmaf(d, d, acc);
  }
  acc = warpsum(acc) * inv_count;
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out_total, acc);
    atomicAdd(out_mse, acc);
  }
}

// gx = (g_total + g_mse) * 2 inv_count (x - target) + g_total * weight * d(MMD)/dx ; gZ = g_total * weight * d(MMD)/dZ.
// Everything that lands in gx is an atomicAdd onto the launcher's zero fill (the MMD term touches the sampled nodes only).
__global__ void __launch_bounds__(128) mse_mmd_bwd_kernel(int N, int B, int C, int ns, float inv2s2, float weight, float svv,
                                                          float srv, float inv_count, const float* __restrict__ x,
                                                          const float* __restrict__ target, const float* __restrict__ Z,
                                                          const int* __restrict__ idx, const float* __restrict__ g_total,
                                                          const float* __restrict__ g_mse, float* __restrict__ gx,
                                                          float* __restrict__ gZ) {
  __shared__ float gz[3 * FEGNN_MAX_C];
  const float gt = g_total != nullptr ? g_total[0] : 0.f, gm = g_mse != nullptr ? g_mse[0] : 0.f;
  if ((int)blockIdx.x < B) {
    const int b = blockIdx.x;
    const float gl = gt * weight;
    const float cvv = svv * gl / ((float)B * C * C), crv = ns > 0 ? srv * 2.f * gl / ((float)B * ns * C) : 0.f;
    if (threadIdx.x < 3 * C) gz[threadIdx.x] = 0.f;
    __syncthreads();
    for (int p = threadIdx.x; p < C * C + ns * C; p += blockDim.x) {
      float px, py, pz, w;
      int c, c0 = -1, i = -1;
      if (p < C * C) {
        c0 = p / C;
        c = p - c0 * C;
        px = Z[((size_t)b * 3 + 0) * C + c0]; py = Z[((size_t)b * 3 + 1) * C + c0]; pz = Z[((size_t)b * 3 + 2) * C + c0];
        w = cvv;
      } else {
        const int q = p - C * C, s = q / C;
        c = q - s * C;
        i = idx[(size_t)b * ns + s];
        px = x[(size_t)i * 3 + 0]; py = x[(size_t)i * 3 + 1]; pz = x[(size_t)i * 3 + 2];
        w = -crv;
      }
      const float dx = px - Z[((size_t)b * 3 + 0) * C + c], dy = py - Z[((size_t)b * 3 + 1) * C + c],
                  dz = pz - Z[((size_t)b * 3 + 2) * C + c];
      const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
      if (dist > 0.f) {
        const float f = -w * inv2s2 * __expf(-dist * inv2s2) / dist;
        const float fx = f * dx, fy = f * dy, fz = f * dz;
        atomicAdd(&gz[0 * C + c], -fx); atomicAdd(&gz[1 * C + c], -fy); atomicAdd(&gz[2 * C + c], -fz);
        if (c0 >= 0) {
          atomicAdd(&gz[0 * C + c0], fx); atomicAdd(&gz[1 * C + c0], fy); atomicAdd(&gz[2 * C + c0], fz);
        } else {
          atomicAdd(gx + (size_t)i * 3 + 0, fx); atomicAdd(gx + (size_t)i * 3 + 1, fy);
          atomicAdd(gx + (size_t)i * 3 + 2, fz);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 3 * C) gZ[(size_t)b * 3 * C + threadIdx.x] = gz[threadIdx.x];
    return;
  }
  const size_t n = (size_t)N * 3, stride = (size_t)(gridDim.x - B) * blockDim.x;
  const float k = (gt + gm) * 2.f * inv_count;
  for (size_t i = (size_t)(blockIdx.x - B) * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(gx + i, k * (x[i] - target[i]));
}

static inline int mse_ctas(int N) {
  const long long n = (long long)N * 3;
  long long c = (n + 128 * 8 - 1) / (128 * 8);
  return (int)(c < 1 ? 1 : (c > 1184 ? 1184 : c));      // at most 8 x 148 CTAs
}
cudaError_t launch_mse_mmd_fwd(int N, int B, int C, int ns, float sigma, float weight, float svv, float srv, float inv_count,
                               const float* x, const float* target, const float* Z, const int* idx, float* out_total,
                               float* out_mse, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out_total, 0, sizeof(float), st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(out_mse, 0, sizeof(float), st);
  if (e != cudaSuccess) return e;
  mse_mmd_fwd_kernel<<<B + mse_ctas(N), 128, 0, st>>>(N, B, C, ns, 1.f / (2.f * sigma * sigma), weight, svv, srv, inv_count, x,
                                                       target, Z, idx, out_total, out_mse); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_mse_mmd_bwd(int N, int B, int C, int ns, float sigma, float weight, float svv, float srv, float inv_count,
                               const float* x, const float* target, const float* Z, const int* idx, const float* g_total,
                               const float* g_mse, float* gx, float* gZ, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * 3 * (size_t)N, st);
  if (e != cudaSuccess) return e;
  mse_mmd_bwd_kernel<<<B + mse_ctas(N), 128, 0, st>>>(N, B, C, ns, 1.f / (2.f * sigma * sigma), weight, svv, srv, inv_count, x,
                                                       target, Z, idx, g_total, g_mse, gx, gZ); ++g_launches;
  return cudaGetLastError();
}

cudaError_t launch_mmd_fwd(int B, int C, int ns, float sigma, float svv, float srv, const float* x, const float* Z, const int* idx,
                           float* loss, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess || B == 0) return e;
  mmd_fwd_kernel<<<B, 128, 0, st>>>(B, C, ns, 1.f / (2.f * sigma * sigma), svv, srv, x, Z, idx, loss); ++g_launches;
  return cudaGetLastError();
}
cudaError_t launch_mmd_bwd(int N, int B, int C, int ns, float sigma, float svv, float srv, const float* x, const float* Z, const int* idx,
                           const float* gloss, float* gx, float* gZ, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * 3 * (size_t)N, st);
  if (e != cudaSuccess || B == 0) return e;
  mmd_bwd_kernel<<<B, 128, 0, st>>>(B, C, ns, 1.f / (2.f * sigma * sigma), svv, srv, x, Z, idx, gloss, gx, gZ); ++g_launches;
  return cudaGetLastError();
}

}  // namespace fegnn
