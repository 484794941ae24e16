// fp64 GSM path for small dimensions (D <= 64): BASELINE.json configs[0], the reference's own CPU-runnable case
// (examples/example_gsm_numpy.py:38-46: D = 10, batch 2, 500 iterations of gsmvi/gsm_numpy.py in numpy fp64).
//
// At these sizes one iteration is a few thousand flops: nothing is gained from tensor cores, and the example target
// (L L^T + 1e-3 I, condition number ~3e4) turns fp32 rounding of the precision matrix alone into a 2e-4 shift of the fit.
// So the whole loop body of gsmvi/gsm.py:107-129 - sample x = mu + L z (gsm.py:117-119), score (built-in dense-Gaussian
// target -(x - m) P, example_gsm_numpy.py:24-29, or the caller's), gsm_update (gsm.py:31-58 in the GEMM restatement of
// gsm_kernels.cu), Cholesky goodness check (gsm.py:136-150) and the accept / revert (gsm.py:125-129) - runs in fp64
// inside ONE CTA, with the commit predicated on the device: the host never reads a flag between iterations, and with
// the built-in target any number of iterations run in a single launch.
#include "gsm_small64.cuh"

#include <math.h>
#include <stdint.h>

namespace gsmvi {

constexpr int S64_THREADS = 256;
constexpr int S64_MAXD = 64;

struct Small64Args {
  double* mu;          // [D]      state
  double* Sigma;       // [D, D]   state (dense, ld = D)
  double* L;           // [D, D]   Cholesky factor of Sigma (state; upper triangle zero)
  const float* ztape;  // [iters, B, D] standard-normal draws of these iterations, or null (Philox)
  unsigned long long seed, iter0;
  double* X;           // [B, D] samples   (written by SAMPLE / FULL, read by UPDATE)
  const double* G;     // [B, D] scores    (UPDATE: the caller's)
  const double* P;     // [D, D] target precision, c [D] = P m   (FULL)
  const double* c;
  int B, D, iters, mode;
  int* status;         // [0] += rejected updates, [1] = 1 if the initial covariance was not PD (INIT), [2] = last update ok
  double* ws;          // scratch: Z [B,D] | Gs [B,D] | W [B,D] | U [B,D] | E [B,D] | Sn [D,D] | Ln [D,D] | mun [D] | ab [2B]
};

__device__ __forceinline__ void philox4_64(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// In-place lower Cholesky of the n x n matrix a (dense, ld = n; only the lower triangle is read), upper triangle zeroed.
// Right-looking, one column per step; returns true when a pivot is non-positive or non-finite (np.linalg.cholesky
// raising LinAlgError / the NaN test of gsm.py:144-147).
__device__ bool chol64_inplace(double* a, int n, int* bad_smem) {
  const int tid = threadIdx.x;
  if (tid == 0) *bad_smem = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    const double p = a[j * n + j];
    if (!(p > 0.0) || isinf(p)) {
      if (tid == 0) *bad_smem = 1;
    }
    const double r = sqrt(p);
    __syncthreads();  // every thread has read the pivot before it is overwritten
    for (int i = j + tid; i < n; i += S64_THREADS) a[i * n + j] = (i == j) ? r : a[i * n + j] / r;
    __syncthreads();
    // trailing update of the lower triangle: a[i][k] -= l[i][j] l[k][j], j < k <= i
    const int m = n - j - 1;
    for (int idx = tid; idx < m * m; idx += S64_THREADS) {
      const int i = j + 1 + idx / m, k = j + 1 + idx % m;
      if (k <= i) a[i * n + k] -= a[i * n + j] * a[k * n + j];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += S64_THREADS)
    if (idx % n > idx / n) a[idx] = 0.0;
  __syncthreads();
  const bool bad = *bad_smem != 0;
  __syncthreads();
  return bad;
}

__global__ void __launch_bounds__(S64_THREADS, 1) gsm_small64_kernel(Small64Args a) {
  __shared__ int bad_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = a.B, D = a.D;
  const long long BD = static_cast<long long>(B) * D;
  double* Z = a.ws;
  double* Gs = Z + BD;
  double* W = Gs + BD;
  double* U = W + BD;
  double* E = U + BD;
  double* Sn = E + BD;
  double* Ln = Sn + D * D;
  double* mun = Ln + D * D;
  double* ab = mun + D;

  if (a.mode == GSMVI_SMALL64_INIT) {
    for (int idx = tid; idx < D * D; idx += S64_THREADS) a.L[idx] = a.Sigma[idx];
    __syncthreads();
    const bool bad = chol64_inplace(a.L, D, &bad_smem);
    if (tid == 0) a.status[1] = bad ? 1 : 0;
    return;
  }

  for (int it = 0; it < a.iters; ++it) {
    if (a.mode != GSMVI_SMALL64_UPDATE) {
      // ---- draws: the caller's tape, or Philox4x32-10 with the counter layout of philox_normal_kernel (gsm_kernels.cu)
      if (a.ztape) {
        const float* zt = a.ztape + static_cast<long long>(it) * BD;
        for (long long idx = tid; idx < BD; idx += S64_THREADS) Z[idx] = static_cast<double>(zt[idx]);
      } else {
        const int gpr = (D + 3) / 4;
        for (long long gid = tid; gid < static_cast<long long>(B) * gpr; gid += S64_THREADS) {
          const unsigned long long off = a.iter0 + it;
          uint32_t c[4] = {static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), static_cast<uint32_t>(off),
                           static_cast<uint32_t>(off >> 32)};
          philox4_64(c, static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));
          const float u0 = (static_cast<float>(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float u1 = (static_cast<float>(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float u2 = (static_cast<float>(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float u3 = (static_cast<float>(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
          const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
          float s0, c0, s1, c1;
          sincospif(2.0f * u1, &s0, &c0);
          sincospif(2.0f * u3, &s1, &c1);
          const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
          const int b = static_cast<int>(gid / gpr), j = static_cast<int>(gid % gpr) * 4;
          for (int t = 0; t < 4 && j + t < D; ++t) Z[static_cast<long long>(b) * D + j + t] = static_cast<double>(z[t]);
        }
      }
      __syncthreads();
      // ---- x_b = mu + L z_b   (gsm.py:117-119 with the Cholesky factor)
      for (long long idx = tid; idx < BD; idx += S64_THREADS) {
        const int b = static_cast<int>(idx / D), i = static_cast<int>(idx % D);
        double acc = a.mu[i];
        for (int k = 0; k <= i; ++k) acc += Z[static_cast<long long>(b) * D + k] * a.L[i * D + k];
        a.X[idx] = acc;
      }
      __syncthreads();
      if (a.mode == GSMVI_SMALL64_SAMPLE) return;
      // ---- g_b = -(x_b - m) P = -x_b P + c   (example_gsm_numpy.py:24-29)
      for (long long idx = tid; idx < BD; idx += S64_THREADS) {
        const int b = static_cast<int>(idx / D), j = static_cast<int>(idx % D);
        double acc = 0.0;
        for (int k = 0; k < D; ++k) acc += a.X[static_cast<long long>(b) * D + k] * a.P[k * D + j];
        Gs[idx] = a.c[j] - acc;
      }
      __syncthreads();
    }
    const double* G = (a.mode == GSMVI_SMALL64_UPDATE) ? a.G : Gs;
    // ---- w_b = Sigma g_b   (gsm.py:11)
    for (long long idx = tid; idx < BD; idx += S64_THREADS) {
      const int b = static_cast<int>(idx / D), j = static_cast<int>(idx % D);
      double acc = 0.0;
      for (int k = 0; k < D; ++k) acc += G[static_cast<long long>(b) * D + k] * a.Sigma[k * D + j];
      W[idx] = acc;
    }
    __syncthreads();
    // ---- per-sample scalars (gsm.py:12-21), a warp per sample
    for (int b = warp; b < B; b += S64_THREADS / 32) {
      double vSv = 0.0, mu_v = 0.0;
      for (int j = lane; j < D; j += 32) {
        const double g = G[static_cast<long long>(b) * D + j];
        vSv += W[static_cast<long long>(b) * D + j] * g;
        mu_v += (a.mu[j] - a.X[static_cast<long long>(b) * D + j]) * g;
      }
      for (int o = 16; o > 0; o >>= 1) {
        vSv += __shfl_xor_sync(0xffffffffu, vSv, o);
        mu_v += __shfl_xor_sync(0xffffffffu, mu_v, o);
      }
      if (lane == 0) {
        const double rho = 0.5 * sqrt(1.0 + 4.0 * (vSv + mu_v * mu_v)) - 0.5;
        const double alpha = 1.0 / (1.0 + rho);
        ab[b] = alpha;
        ab[B + b] = -alpha * (1.0 + (vSv - mu_v) / (1.0 + rho + mu_v));
      }
    }
    __syncthreads();
    // ---- u = alpha w + beta d (= mu_update), e = d + u   (gsm.py:21-22)
    for (long long idx = tid; idx < BD; idx += S64_THREADS) {
      const int b = static_cast<int>(idx / D), j = static_cast<int>(idx % D);
      const double d = a.mu[j] - a.X[idx];
      const double u = ab[b] * W[idx] + ab[B + b] * d;
      U[idx] = u;
      E[idx] = d + u;
    }
    __syncthreads();
    // ---- mu_new = mu + mean_b u ; Sigma_new = Sigma + mean_b (d d^T - e e^T)   (gsm.py:25-27, 53-56), lower + mirror
    for (int j = tid; j < D; j += S64_THREADS) {
      double acc = 0.0;
      for (int b = 0; b < B; ++b) acc += U[static_cast<long long>(b) * D + j];
      mun[j] = a.mu[j] + acc / B;
    }
    for (int idx = tid; idx < D * D; idx += S64_THREADS) {
      const int i = idx / D, j = idx % D;
      if (j > i) continue;
      double acc = 0.0;
      for (int b = 0; b < B; ++b) {
        const long long r = static_cast<long long>(b) * D;
        const double di = a.mu[i] - a.X[r + i], dj = a.mu[j] - a.X[r + j];
        acc += di * dj - E[r + i] * E[r + j];
      }
      const double v = a.Sigma[idx] + acc / B;
      Sn[idx] = v;
      Sn[j * D + i] = v;
      Ln[idx] = v;
    }
    __syncthreads();
    // ---- goodness check = Cholesky of the proposal (gsm.py:136-150); accept / revert on the device (gsm.py:125-129)
    const bool bad = chol64_inplace(Ln, D, &bad_smem);
    if (!bad) {
      for (int idx = tid; idx < D * D; idx += S64_THREADS) {
        a.Sigma[idx] = Sn[idx];
        a.L[idx] = Ln[idx];
      }
      for (int j = tid; j < D; j += S64_THREADS) a.mu[j] = mun[j];
    }
    if (tid == 0) {
      if (bad) a.status[0] += 1;
      a.status[2] = bad ? 0 : 1;
    }
    __syncthreads();
  }
}

long long gsm_small64_workspace_bytes(int B, int D) {
  if (B <= 0 || D <= 0 || D > S64_MAXD) return -1;
  return static_cast<long long>(5LL * B * D + 2LL * D * D + D + 2LL * B) * sizeof(double);
}

int gsm_small64(cudaStream_t st, int mode, double* mu, double* Sigma, double* L, const float* ztape, unsigned long long seed,
                unsigned long long iter0, double* X, const double* G, const double* P, const double* c, int B, int D,
                int iters, int* status, void* workspace) {
  if (!mu || !Sigma || !L || !status || !workspace || B <= 0 || D <= 0 || D > S64_MAXD || iters < 0) return GSMVI_EINVAL;
  if (mode < GSMVI_SMALL64_INIT || mode > GSMVI_SMALL64_UPDATE) return GSMVI_EINVAL;
  if (mode == GSMVI_SMALL64_FULL && (!P || !c || !X)) return GSMVI_EINVAL;
  if (mode == GSMVI_SMALL64_SAMPLE && !X) return GSMVI_EINVAL;
  if (mode == GSMVI_SMALL64_UPDATE && (!X || !G)) return GSMVI_EINVAL;
  if (mode == GSMVI_SMALL64_SAMPLE || mode == GSMVI_SMALL64_UPDATE) iters = 1;
  Small64Args a{mu, Sigma, L, ztape, seed, iter0, X, G, P, c, B, D, iters, mode, status, static_cast<double*>(workspace)};
  gsm_small64_kernel<<<1, S64_THREADS, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
