// Batched small-matrix path: an ensemble of independent GSM fits, one persistent CTA per fit (BASELINE.json configs[4]:
// 1024 fits, D = 64, batch 32; SURVEY.md section 8e "replicas only", no communication).
//
// Each fit is the complete loop of gsmvi/gsm.py:107-129 - sample, dense-Gaussian score (examples/
// example_gsm_numpy.py:24-29), gsm_update (gsm.py:31-58), Cholesky goodness check (gsm.py:136-150), accept/revert - run
// for all iterations inside one kernel with (mu, Sigma, L, P) resident in shared memory (~96 KB per fit, two fits
// per SM).  D <= 64 and B <= 32 are far below a tensor-core tile, so everything is exact fp32 FMA arithmetic with
// conflict-free float4 shared-memory tiles; the same Philox stream, row formulas and low-cancellation covariance
// update as the large path (gsm_kernels.cu).
#include "dev_once.cuh"
#include "gsm_ensemble.cuh"

#include <math.h>

#include "chol_block.cuh"

namespace gsmvi {

constexpr int ED = 64;        // max dimension
constexpr int EB = 32;        // max batch
constexpr int ELD = ED + 4;   // shared leading dimension: rows 16-byte aligned, consecutive rows 4 banks apart
constexpr int ETHREADS = 256;

__device__ __forceinline__ void philox4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
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

// out[b][n] = alpha * sum_k A[b][k] * Bm[n][k] + bias[n]   (A: [EB][ELD], Bm: [ED][ELD], both row-major over k).
// Thread (r0 = tid / 16, c = tid % 16) owns rows {r0, r0 + 16} and columns {c, c + 16, c + 32, c + 48}.
__device__ __forceinline__ void mm_bd(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ out,
                                      float alpha, const float* __restrict__ bias, int kmax) {
  const int r0 = threadIdx.x >> 4, c = threadIdx.x & 15;
  float acc[2][4] = {};
  for (int k = 0; k < kmax; k += 4) {
    float4 a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const float4*>(A + (r0 + 16 * i) * ELD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bm + (c + 16 * j) * ELD + k);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i].x * b[j].x + a[i].y * b[j].y + a[i].z * b[j].z + a[i].w * b[j].w;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(r0 + 16 * i) * ELD + c + 16 * j] = alpha * acc[i][j] + (bias ? bias[c + 16 * j] : 0.0f);
}

// In-place Cholesky of the 64x64 matrix in `s` (lower triangle valid; upper triangle is zeroed); returns true if bad.
__device__ __forceinline__ bool chol64(float* s, float* dinv, int* bad_smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) *bad_smem = 0;
  __syncthreads();
  for (int p = 0; p < 2; ++p) {
    const int c0 = 32 * p;
    if (warp == 0) {
      float row[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) row[k] = s[(c0 + lane) * ELD + c0 + k];
      int isbad = 0;
      chol32_step<0>(row, lane, dinv + c0, isbad);
#pragma unroll
      for (int k = 0; k < 32; ++k) s[(c0 + lane) * ELD + c0 + k] = (k <= lane) ? row[k] : 0.0f;
      if (isbad && lane == 0) *bad_smem = 1;
    }
    __syncthreads();
    if (p == 0) {
      if (tid >= 32 && tid < 64) {  // rows 32..63: x L11^T = a
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 t = *reinterpret_cast<const float4*>(s + tid * ELD + k);
          v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a0 = v[j], a1 = 0.0f;
          const float* lrow = s + j * ELD;
#pragma unroll
          for (int k = 0; k + 1 < j; k += 2) {
            a0 -= v[k] * lrow[k];
            a1 -= v[k + 1] * lrow[k + 1];
          }
          if (j & 1) a0 -= v[j - 1] * lrow[j - 1];
          v[j] = (a0 + a1) * dinv[j];
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(s + tid * ELD + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
      __syncthreads();
      // trailing 32x32 block: S[i][k] -= sum_c P[i][c] P[k][c]; thread owns (i = 32 + tid/8, k = 32 + 4 (tid%8) ..+3)
      {
        const int i = 32 + (tid >> 3), k0 = 32 + 4 * (tid & 7);
        float acc[4] = {};
        for (int cc = 0; cc < 32; cc += 4) {
          const float4 a = *reinterpret_cast<const float4*>(s + i * ELD + cc);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(s + (k0 + q) * ELD + cc);
            acc[q] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
          }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (k0 + q <= i) s[i * ELD + k0 + q] -= acc[q];
      }
      __syncthreads();
    }
  }
  // zero the upper-right 32x32 block (rows 0..31, cols 32..63)
  for (int idx = tid; idx < 32 * 32; idx += ETHREADS) s[(idx >> 5) * ELD + 32 + (idx & 31)] = 0.0f;
  __syncthreads();
  return *bad_smem != 0;
}

__global__ void __launch_bounds__(ETHREADS, 2)
gsm_ensemble_kernel(const float* __restrict__ Pg, const float* __restrict__ cg, float* __restrict__ mug,
                    float* __restrict__ Sg, int F, int D, int B, int niter, unsigned long long seed, int first_fit,
                    const float* __restrict__ ztape, long long z_fit_stride, int* __restrict__ reverts) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                      // current Sigma
  float* Sn = S + ED * ELD;           // proposal
  float* L = Sn + ED * ELD;           // factor of S (work area for the proposal's factor)
  float* P = L + ED * ELD;            // target precision
  float* X = P + ED * ELD;            // [EB][ELD]  samples -> D = mu - x
  float* G = X + EB * ELD;            //            scores  -> E
  float* W = G + EB * ELD;            //            Z, then W = G Sigma -> U
  float* mu = W + EB * ELD;           // [ED]
  float* mun = mu + ED;
  float* cv = mun + ED;
  float* dinv = cv + ED;
  float* al = dinv + ED;              // [EB] alpha, beta
  float* be = al + EB;
  __shared__ int bad;
  const int f = blockIdx.x;
  if (f >= F) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- load state (padded with identity / zeros up to 64 x 64, so the padded problem is block diagonal)
  for (int idx = tid; idx < ED * ED; idx += ETHREADS) {
    const int i = idx >> 6, j = idx & 63;
    const bool in = i < D && j < D;
    S[i * ELD + j] = in ? Sg[(static_cast<long long>(f) * D + i) * D + j] : ((i == j) ? 1.0f : 0.0f);
    P[i * ELD + j] = in ? Pg[(static_cast<long long>(f) * D + i) * D + j] : 0.0f;
  }
  for (int i = tid; i < ED; i += ETHREADS) {
    mu[i] = i < D ? mug[static_cast<long long>(f) * D + i] : 0.0f;
    cv[i] = i < D ? cg[static_cast<long long>(f) * D + i] : 0.0f;
  }
  for (int idx = tid; idx < EB * ELD; idx += ETHREADS) X[idx] = G[idx] = W[idx] = 0.0f;
  __syncthreads();
  for (int idx = tid; idx < ED * ED; idx += ETHREADS) L[(idx >> 6) * ELD + (idx & 63)] = ((idx & 63) <= (idx >> 6)) ? S[(idx >> 6) * ELD + (idx & 63)] : 0.0f;
  __syncthreads();
  int nrev = 0;
  if (chol64(L, dinv, &bad)) {  // initial covariance not PD: report and leave the state untouched
    if (tid == 0) reverts[f] = -1;
    return;
  }
  const float invB = 1.0f / static_cast<float>(B);

  for (int it = 0; it <= niter; ++it) {  // gsm.py:107: niter + 1 updates
    // ---- (1) z ~ N(0, I) into W (rows >= B and columns >= D stay zero)
    if (ztape) {
      const float* zt = ztape + f * z_fit_stride + static_cast<long long>(it) * B * D;
      for (int idx = tid; idx < EB * ED; idx += ETHREADS) {
        const int b = idx >> 6, j = idx & 63;
        W[b * ELD + j] = (b < B && j < D) ? zt[b * D + j] : 0.0f;
      }
    } else {
      for (int g4 = tid; g4 < EB * ED / 4; g4 += ETHREADS) {
        const int b = g4 >> 4, j = (g4 & 15) * 4;
        uint32_t c[4] = {static_cast<uint32_t>(g4), static_cast<uint32_t>(first_fit + f), static_cast<uint32_t>(it), 0x454e53u};  // global fit index
        philox4(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        const float u0 = (static_cast<float>(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u1 = (static_cast<float>(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = (static_cast<float>(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u3 = (static_cast<float>(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
        float s0, c0, s1, c1;
        sincospif(2.0f * u1, &s0, &c0);
        sincospif(2.0f * u3, &s1, &c1);
        const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
        for (int t = 0; t < 4; ++t) W[b * ELD + j + t] = (b < B && j + t < D) ? z[t] : 0.0f;
      }
    }
    __syncthreads();
    // ---- (2) X = mu + Z L^T   (gsm.py:119)
    mm_bd(W, L, X, 1.0f, mu, ED);
    __syncthreads();
    // ---- (3) G = -X P + c   (score, gsm.py:121; P symmetric)
    mm_bd(X, P, G, -1.0f, cv, ED);
    __syncthreads();
    // ---- (4) W = G Sigma   (Sigma symmetric)
    mm_bd(G, S, W, 1.0f, nullptr, ED);
    __syncthreads();
    // ---- (5) per-sample scalars (gsm.py:12-21): warp handles rows warp, warp + 8, ...
    for (int b = warp; b < EB; b += ETHREADS / 32) {
      float vSv = 0.0f, mu_v = 0.0f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        const float g = G[b * ELD + j];
        vSv += W[b * ELD + j] * g;
        mu_v += (mu[j] - X[b * ELD + j]) * g;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vSv += __shfl_xor_sync(0xffffffffu, vSv, o);
        mu_v += __shfl_xor_sync(0xffffffffu, mu_v, o);
      }
      if (lane == 0) {
        const float rho = 0.5f * sqrtf(1.0f + 4.0f * (vSv + mu_v * mu_v)) - 0.5f;
        const float a = 1.0f / (1.0f + rho);
        al[b] = (b < B) ? a : 0.0f;
        be[b] = (b < B) ? -a * (1.0f + (vSv - mu_v) / (1.0f + rho + mu_v)) : 0.0f;
      }
    }
    __syncthreads();
    // ---- (6) D = mu - x (in X), U = alpha w + beta d (in W), E = d + u (in G); padded rows are zeroed
    for (int idx = tid; idx < EB * ED; idx += ETHREADS) {
      const int b = idx >> 6, j = idx & 63;
      const float d = (b < B && j < D) ? mu[j] - X[b * ELD + j] : 0.0f;
      const float u = al[b] * W[b * ELD + j] + be[b] * d;
      X[b * ELD + j] = d;
      W[b * ELD + j] = (b < B) ? u : 0.0f;
      G[b * ELD + j] = (b < B) ? d + u : 0.0f;
    }
    __syncthreads();
    if (tid < ED) {  // mu_new = mu + mean_b u   (gsm.py:53,55)
      float acc = 0.0f;
      for (int b = 0; b < EB; ++b) acc += W[b * ELD + tid];
      mun[tid] = mu[tid] + acc * invB;
    }
    // ---- (7) Sn = S - (E^T U + U^T D) / B   (gsm.py:25-27,54 in low-cancellation form); thread owns a 4x4 tile
    {
      const int ti = tid >> 4, tj = tid & 15;  // rows ti + 16 a, cols tj + 16 c
      float acc[4][4] = {};
      for (int b = 0; b < EB; ++b) {
        float e[4], ui[4], uj[4], dj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          e[a] = G[b * ELD + ti + 16 * a];
          ui[a] = W[b * ELD + ti + 16 * a];
          uj[a] = W[b * ELD + tj + 16 * a];
          dj[a] = X[b * ELD + tj + 16 * a];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] += e[a] * uj[c] + ui[a] * dj[c];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = ti + 16 * a, j = tj + 16 * c;
          Sn[i * ELD + j] = S[i * ELD + j] - acc[a][c] * invB;
        }
    }
    __syncthreads();
    // symmetrise exactly (the two triangles are computed by different threads) and stage the lower triangle in L
    for (int idx = tid; idx < ED * ED; idx += ETHREADS) {
      const int i = idx >> 6, j = idx & 63;
      if (j < i) {
        const float v = Sn[i * ELD + j];
        Sn[j * ELD + i] = v;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < ED * ED; idx += ETHREADS) {
      const int i = idx >> 6, j = idx & 63;
      L[i * ELD + j] = (j <= i) ? Sn[i * ELD + j] : 0.0f;
    }
    __syncthreads();
    // ---- (8) goodness check = Cholesky of the proposal (gsm.py:125, 136-150); accept or revert
    const bool isbad = chol64(L, dinv, &bad);
    if (!isbad) {
      float* t = S; S = Sn; Sn = t;
      t = mu; mu = mun; mun = t;
    } else {
      ++nrev;
      for (int idx = tid; idx < ED * ED; idx += ETHREADS) {
        const int i = idx >> 6, j = idx & 63;
        L[i * ELD + j] = (j <= i) ? S[i * ELD + j] : 0.0f;
      }
      __syncthreads();
      chol64(L, dinv, &bad);  // restore the factor of the kept covariance
    }
    __syncthreads();
  }
  // ---- store
  for (int idx = tid; idx < D * D; idx += ETHREADS) {
    const int i = idx / D, j = idx % D;
    Sg[(static_cast<long long>(f) * D + i) * D + j] = S[i * ELD + j];
  }
  for (int i = tid; i < D; i += ETHREADS) mug[static_cast<long long>(f) * D + i] = mu[i];
  if (tid == 0) reverts[f] = nrev;
}

int gsm_ensemble_fit(cudaStream_t st, const float* P, const float* c, float* mu, float* Sigma, int F, int D, int B,
                     int niter, unsigned long long seed, const float* ztape, int* reverts, int first_fit) {
  if (!P || !c || !mu || !Sigma || !reverts || F <= 0 || D <= 0 || D > ED || B <= 0 || B > EB || niter < 0) return GSMVI_EINVAL;
  const int smem = (4 * ED * ELD + 3 * EB * ELD + 4 * ED + 2 * EB) * sizeof(float);
  static PerDeviceOnce attr_set;
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(gsm_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  gsm_ensemble_kernel<<<F, ETHREADS, smem, st>>>(P, c, mu, Sigma, F, D, B, niter, seed, first_fit, ztape,
                                                 static_cast<long long>(niter + 1) * B * D, reverts);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
