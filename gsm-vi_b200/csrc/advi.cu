// ADVI on the GSM / BaM kernels: full-rank Gaussian q = N(mu, L L^T) fitted by stochastic
// gradient ascent of the ELBO with reparameterised samples - gsmvi/advi.py:31-112 (neg_elbo advi.py:31-45, the Adam step of
// optax.adam advi.py:69-74).  With x_b = mu + L z_b:
//     -ELBO = -( sum_b log p(x_b) - sum_b log q(x_b) ),   sum_b log q(x_b) = -1/2 sum_b |z_b|^2 - B sum_i log L_ii - const
//     d(-ELBO)/d mu   = -sum_b g_b                         (g_b = grad log p(x_b), the same scores GSM / BaM consume)
//     d(-ELBO)/d L_ij = -(G^T Z)_ij - B delta_ij / L_ii    (i >= j: the parameters are the lower triangle, advi.py:21-29)
// jax.grad of the reference's loss is exactly this (checked against finite differences in tests/test_host_cpu.py).
// The D x D product G^T Z runs on the tensor-core GEMM (tc_gemm.cuh, lower tiles only); this file fuses the rest: the
// gradient assembly and the Adam update of (mu, L) in one launch.
#include "advi.cuh"

#include <math.h>

namespace gsmvi {

// Adam (optax.adam defaults: b1 = 0.9, b2 = 0.999, eps = 1e-8, eps_root = 0) on the lower triangle of L and on mu.
//   g_L[i][j] = -GtZ[i][j] - B / L_ii (i == j);  g_mu[j] = -gsum[j]
//   m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;  p <- p - lr (m / c1) / (sqrt(v / c2) + eps),  c_k = 1 - b_k^t
__global__ void __launch_bounds__(256) advi_adam_kernel(float* __restrict__ Lm, long long ldl, const float* __restrict__ GtZ,
                                                        long long ldg, float* __restrict__ mL, float* __restrict__ vL,
                                                        float* __restrict__ mu, const float* __restrict__ gsum,
                                                        float* __restrict__ m_mu, float* __restrict__ v_mu, int D, int B,
                                                        float lr, float b1, float b2, float eps, float c1, float c2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j >= D) return;
  if (i == D) {  // the extra grid row updates mu
    const float g = -gsum[j];
    const float m = b1 * m_mu[j] + (1.0f - b1) * g;
    const float v = b2 * v_mu[j] + (1.0f - b2) * g * g;
    m_mu[j] = m;
    v_mu[j] = v;
    mu[j] -= lr * (m / c1) / (sqrtf(v / c2) + eps);
    return;
  }
  if (j > i) return;
  const long long o = i * ldl + j;
  float g = -GtZ[i * ldg + j];
  if (i == j) g -= static_cast<float>(B) / Lm[o];
  const float m = b1 * mL[o] + (1.0f - b1) * g;
  const float v = b2 * vL[o] + (1.0f - b2) * g * g;
  mL[o] = m;
  vL[o] = v;
  Lm[o] -= lr * (m / c1) / (sqrtf(v / c2) + eps);
}

// gsum[j] = sum_b G[b][j]  (fp32 accumulation in fp64 per CTA column strip, then one atomicAdd per strip of 32 rows)
__global__ void __launch_bounds__(256) advi_colsum_kernel(const float* __restrict__ G, long long ldg, int B, int D,
                                                          float* __restrict__ gsum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * 32;
  if (j >= D) return;
  double acc = 0.0;
  const int r1 = min(r0 + 32, B);
  for (long long b = r0; b < r1; ++b) acc += G[b * ldg + j];
  atomicAdd(gsum + j, static_cast<float>(acc));
}

int advi_step(cudaStream_t st, float* L, long long ldl, float* mu, const float* G, long long ldg, const float* Z, long long ldz,
              float* GtZ, long long ldgz, float* gsum, float* mL, float* vL, float* m_mu, float* v_mu, int B, int D, float lr,
              float b1, float b2, float eps, int t, int npass) {
  if (!L || !mu || !G || !Z || !GtZ || !gsum || !mL || !vL || !m_mu || !v_mu || B <= 0 || D <= 0 || t < 1) return GSMVI_EINVAL;
  cudaError_t e = cudaMemsetAsync(gsum, 0, static_cast<size_t>(D) * sizeof(float), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  advi_colsum_kernel<<<dim3((D + 255) / 256, (B + 31) / 32), 256, 0, st>>>(G, ldg, B, D, gsum);
  // G^T Z: rows of G and Z are the contraction index, so both operands are MN-major views; lower tiles only
  GemmOpts o;
  o.npass = npass;
  o.a_mn = o.b_mn = true;
  o.tri = true;
  MatView vg{G, B, D, ldg}, vz{Z, B, D, ldz};
  int rc = launch_gemm_tf32(st, D, D, B, vg, vz, GtZ, ldgz, o);
  if (rc != GSMVI_OK) return rc;
  const float c1 = 1.0f - powf(b1, static_cast<float>(t)), c2 = 1.0f - powf(b2, static_cast<float>(t));
  advi_adam_kernel<<<dim3((D + 255) / 256, D + 1), 256, 0, st>>>(L, ldl, GtZ, ldgz, mL, vL, mu, gsum, m_mu, v_mu, D, B, lr, b1,
                                                               b2, eps, c1, c2);
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
