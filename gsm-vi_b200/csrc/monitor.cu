// KLMonitor device reductions: sum_b log q(x_b) for q = N(mu, L L^T).
//
// Replaces numpyro.distributions.MultivariateNormal(mu, cov).log_prob inside reverse_kl / forward_kl
// (gsmvi/monitors.py:10-22, 107-113):  log q(x) = -1/2 |L^{-1}(x - mu)|^2 - sum_i log L_ii - D/2 log(2 pi).
//  * samples drawn from q as x = mu + L z  =>  L^{-1}(x - mu) = z: only |z|^2 and the log-determinant are needed
//    (reverse KL; HBM-bound reduction over Z and diag(L));
//  * arbitrary x (forward KL on reference samples): one CTA per sample does the forward substitution y = L^{-1}(x-mu).
#include "dev_once.cuh"
#include "monitor.cuh"

#include <math.h>

namespace gsmvi {

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < (blockDim.x >> 5)) t = sh[threadIdx.x];
  if (warp == 0) {
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// out[0] += -1/2 sum over this CTA's rows of |z_b|^2 ; CTA 0 also adds  -N (sum_i log L_ii + D/2 log 2pi)
__global__ void logq_from_z_kernel(const float* __restrict__ Z, long long ldz, int N, int D, const float* __restrict__ L,
                                   long long ldl, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long b = blockIdx.x; b < N; b += gridDim.x)
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      const double z = Z[b * ldz + j];
      acc += z * z;
    }
  double tot = -0.5 * block_sum(acc, sh);
  if (blockIdx.x == 0) {
    double ld = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) ld += log(static_cast<double>(L[static_cast<long long>(i) * ldl + i]));
    ld = block_sum(ld, sh);
    if (threadIdx.x == 0) tot -= static_cast<double>(N) * (ld + 0.5 * D * 1.8378770664093453);  // log(2 pi)
  }
  if (threadIdx.x == 0) atomicAdd(out, tot);
}

// One CTA per sample: y = L^{-1}(x - mu) by forward substitution (y kept in shared memory), out[0] += log q(x).
__global__ void logq_from_x_kernel(const float* __restrict__ X, long long ldx, int D, const float* __restrict__ mu,
                                   const float* __restrict__ L, long long ldl, double* __restrict__ out) {
  extern __shared__ float y[];  // D floats
  __shared__ double sh[32];
  __shared__ float yi;
  const float* x = X + static_cast<long long>(blockIdx.x) * ldx;
  double maha = 0.0, logdet = 0.0;
  for (int i = 0; i < D; ++i) {
    const float* Li = L + static_cast<long long>(i) * ldl;
    double acc = 0.0;
    for (int k = threadIdx.x; k < i; k += blockDim.x) acc += static_cast<double>(Li[k]) * y[k];
    const double s = block_sum(acc, sh);
    if (threadIdx.x == 0) {
      const double d = Li[i];
      const double v = (static_cast<double>(x[i]) - mu[i] - s) / d;
      yi = static_cast<float>(v);
      y[i] = yi;
      maha += v * v;
      logdet += log(d);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, -0.5 * maha - logdet - 0.5 * D * 1.8378770664093453);
}

int gauss_logq_from_z(cudaStream_t st, const float* Z, long long ldz, int N, int D, const float* L, long long ldl,
                      double* out) {
  if (!Z || !L || !out || N <= 0 || D <= 0) return GSMVI_EINVAL;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = N < 592 ? N : 592;
  logq_from_z_kernel<<<grid, 256, 0, st>>>(Z, ldz, N, D, L, ldl, out);
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int gauss_logq_from_x(cudaStream_t st, const float* X, long long ldx, int N, int D, const float* mu, const float* L,
                      long long ldl, double* out) {
  if (!X || !mu || !L || !out || N <= 0 || D <= 0 || D > 48 * 1024) return GSMVI_EINVAL;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  static PerDeviceOnce attr_set;
  if (!attr_set.get()) {
    e = cudaFuncSetAttribute(logq_from_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 * 4);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  logq_from_x_kernel<<<N, 256, D * sizeof(float), st>>>(X, ldx, D, mu, L, ldl, out);
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
