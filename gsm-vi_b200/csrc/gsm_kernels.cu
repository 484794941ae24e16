// Bandwidth-bound kernels of the GSM iteration and its orchestration.
//
// gsm_update (gsmvi/gsm.py:31-58) restated for the device (derivation: SURVEY.md section 9, oracle/gsmvi_oracle.py
// gsm_update):   per sample b, d = mu0 - x, w = Sigma0 g:
//     vSv = <w,g>, mu_v = <d,g>, rho = 0.5 sqrt(1 + 4 (vSv + mu_v^2)) - 0.5           (gsm.py:12-14)
//     alpha = 1/(1+rho),  beta = -alpha (1 + (vSv - mu_v)/(1 + rho + mu_v))            (gsm.py:18-21, g^T eps0 = vSv - mu_v)
//     u = alpha w + beta d  (= mu_update),  e = d + u  (= mu - x)                     (gsm.py:21-22)
//     mu = mu0 + mean_b u,   Sigma = Sigma0 + (D^T D - E^T E)/B                       (gsm.py:25-27, 53-56)
// The difference of the two Gram matrices cancels to ~1% of either near convergence, so it is formed as
//     D^T D - E^T E = -(E^T U + U^T D)            (E = D + U; exact identity, every term already small)
// which one GEMM computes from T = [E; U; D]: A = T[0:2B] and B = T[B:3B] (both MN-major), K = 2B.
// Three launches: W = G Sigma0 (tensor-core GEMM), the row pass below (HBM-bound, 24 B D bytes), and that GEMM
// with the "+ Sigma0" and -1/B fused in its epilogue (lower tiles only, mirrored stores).
#include "gsm_kernels.cuh"
#include "h3_gemm.cuh"
#include "comm.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace gsmvi {

constexpr int RP_ROWS = 16;     // sample rows per CTA
constexpr int RP_THREADS = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// hi = tf32_rn(v), lo = tf32_rn(v - hi): the covariance GEMM consumes T pre-split (no in-kernel conversion).
__device__ __forceinline__ void store_split1(float* hi, float* lo, float v) {
  const float h = ptx::to_tf32(v);
  *hi = h;
  *lo = ptx::to_tf32(v - h);
}
__device__ __forceinline__ void store_split4(float* hi, float* lo, const float4 v) {
  float4 h, l;
  h.x = ptx::to_tf32(v.x); l.x = ptx::to_tf32(v.x - h.x);
  h.y = ptx::to_tf32(v.y); l.y = ptx::to_tf32(v.y - h.y);
  h.z = ptx::to_tf32(v.z); l.z = ptx::to_tf32(v.z - h.z);
  h.w = ptx::to_tf32(v.w); l.w = ptx::to_tf32(v.w - h.w);
  *reinterpret_cast<float4*>(hi) = h;
  *reinterpret_cast<float4*>(lo) = l;
}

__device__ __forceinline__ unsigned max4_bits(const float4 v) {
  return max(max(__float_as_uint(fabsf(v.x)), __float_as_uint(fabsf(v.y))),
             max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w))));
}

// hi = tf32_rn(a), lo = tf32_rn(a - hi): the round-to-nearest 3xTF32 split of a reused GEMM operand, done once
__global__ void tf32_split_kernel(const float* __restrict__ A, long long lda, float* __restrict__ Hi, float* __restrict__ Lo,
                                  long long ldo, int rows, int cols) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long i = blockIdx.y;
  if (j >= cols) return;
  if (j + 3 < cols && ((lda | ldo) & 3) == 0 &&
      (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Hi) | reinterpret_cast<uintptr_t>(Lo)) & 15) == 0)) {
    store_split4(Hi + i * ldo + j, Lo + i * ldo + j, *reinterpret_cast<const float4*>(A + i * lda + j));
  } else {
    for (int t = 0; t < 4 && j + t < cols; ++t) store_split1(Hi + i * ldo + j + t, Lo + i * ldo + j + t, A[i * lda + j + t]);
  }
}

// One CTA handles RP_ROWS consecutive samples.  Pass 1 (a warp per row): the two dot products and the per-sample
// scalars.  Pass 2 (a thread per 4 columns): rows e, u, d of T = [E; U; D] and the column sums of u.
template <int MODE>
__global__ void __launch_bounds__(RP_THREADS) gsm_rowpass_kernel(const float* __restrict__ X, long long ldx,
                                                                 const float* __restrict__ G, long long ldg,
                                                                 const float* __restrict__ W, long long ldw,
                                                                 const float* __restrict__ mu, float* __restrict__ T,
                                                                 float* __restrict__ Tlo, long long ldt,
                                                                 float* __restrict__ usum, int B, int D,
                                                                 unsigned* __restrict__ absmax, int rows_per_cta) {
  __shared__ float s_alpha[RP_ROWS], s_beta[RP_ROWS];
  unsigned amax = 0u;  // MODE 1: bit pattern of max |e|, |u|, |d| (as unsigned, NaN > Inf > finite)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * rows_per_cta;  // rows_per_cta <= RP_ROWS; small batches use fewer so the grid fills the GPU
  const bool vec = ((D & 3) == 0) && ((ldx & 3) == 0) && ((ldg & 3) == 0) && ((ldw & 3) == 0) && ((ldt & 3) == 0) &&
                   (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(W) |
                      reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(mu)) & 15) == 0) &&
                   (MODE == 1 || (reinterpret_cast<uintptr_t>(Tlo) & 15) == 0);
  for (int r = warp; r < rows_per_cta; r += RP_THREADS / 32) {
    const int b = row0 + r;
    if (b >= B) break;
    const float* x = X + b * ldx;
    const float* g = G + b * ldg;
    const float* w = W + b * ldw;
    float vSv = 0.0f, mu_v = 0.0f;
    if (vec) {
      for (int j = lane * 4; j < D; j += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(x + j);
        const float4 gv = *reinterpret_cast<const float4*>(g + j);
        const float4 wv = *reinterpret_cast<const float4*>(w + j);
        const float4 mv = *reinterpret_cast<const float4*>(mu + j);
        vSv += wv.x * gv.x + wv.y * gv.y + wv.z * gv.z + wv.w * gv.w;
        mu_v += (mv.x - xv.x) * gv.x + (mv.y - xv.y) * gv.y + (mv.z - xv.z) * gv.z + (mv.w - xv.w) * gv.w;
      }
    } else {
      for (int j = lane; j < D; j += 32) {
        vSv += w[j] * g[j];
        mu_v += (mu[j] - x[j]) * g[j];
      }
    }
    vSv = warp_sum(vSv);
    mu_v = warp_sum(mu_v);
    if (lane == 0) {
      const float rho = 0.5f * sqrtf(1.0f + 4.0f * (vSv + mu_v * mu_v)) - 0.5f;
      const float alpha = 1.0f / (1.0f + rho);
      const float beta = -alpha * (1.0f + (vSv - mu_v) / (1.0f + rho + mu_v));
      s_alpha[r] = alpha;
      s_beta[r] = beta;
    }
  }
  __syncthreads();
  const int nrows = min(rows_per_cta, B - row0);
  if (vec) {
    for (int j = threadIdx.x * 4; j < D; j += RP_THREADS * 4) {
      const float4 mv = *reinterpret_cast<const float4*>(mu + j);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // four rows per trip: their eight loads are issued before any is consumed (the pass is HBM / L2 latency bound)
      for (int r0 = 0; r0 < nrows; r0 += 4) {
        float4 xv[4], wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long b = row0 + min(r0 + q, nrows - 1);
          xv[q] = *reinterpret_cast<const float4*>(X + b * ldx + j);
          wv[q] = *reinterpret_cast<const float4*>(W + b * ldw + j);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = r0 + q;
          if (r >= nrows) break;
          const long long b = row0 + r;
          const float al = s_alpha[r], be = s_beta[r];
          float4 d, u, e;
          d.x = mv.x - xv[q].x; d.y = mv.y - xv[q].y; d.z = mv.z - xv[q].z; d.w = mv.w - xv[q].w;
          u.x = al * wv[q].x + be * d.x; u.y = al * wv[q].y + be * d.y; u.z = al * wv[q].z + be * d.z; u.w = al * wv[q].w + be * d.w;
          e.x = d.x + u.x; e.y = d.y + u.y; e.z = d.z + u.z; e.w = d.w + u.w;
          acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
          if (MODE == 0) {
            store_split4(T + b * ldt + j, Tlo + b * ldt + j, e);
            store_split4(T + (b + B) * ldt + j, Tlo + (b + B) * ldt + j, u);
            store_split4(T + (b + 2LL * B) * ldt + j, Tlo + (b + 2LL * B) * ldt + j, d);
          } else {
            *reinterpret_cast<float4*>(T + b * ldt + j) = e;
            *reinterpret_cast<float4*>(T + (b + B) * ldt + j) = u;
            *reinterpret_cast<float4*>(T + (b + 2LL * B) * ldt + j) = d;
            amax = max(amax, max4_bits(e));
            amax = max(amax, max4_bits(u));
            amax = max(amax, max4_bits(d));
          }
        }
      }
      atomicAdd(usum + j + 0, acc.x);
      atomicAdd(usum + j + 1, acc.y);
      atomicAdd(usum + j + 2, acc.z);
      atomicAdd(usum + j + 3, acc.w);
    }
  } else {
    for (int j = threadIdx.x; j < D; j += RP_THREADS) {
      const float m = mu[j];
      float acc = 0.0f;
      for (int r = 0; r < nrows; ++r) {
        const long long b = row0 + r;
        const float d = m - X[b * ldx + j];
        const float u = s_alpha[r] * W[b * ldw + j] + s_beta[r] * d;
        acc += u;
        if (MODE == 0) {
          store_split1(T + b * ldt + j, Tlo + b * ldt + j, d + u);
          store_split1(T + (b + B) * ldt + j, Tlo + (b + B) * ldt + j, u);
          store_split1(T + (b + 2LL * B) * ldt + j, Tlo + (b + 2LL * B) * ldt + j, d);
        } else {
          T[b * ldt + j] = d + u;
          T[(b + B) * ldt + j] = u;
          T[(b + 2LL * B) * ldt + j] = d;
          amax = max(amax, max(__float_as_uint(fabsf(d + u)), max(__float_as_uint(fabsf(u)), __float_as_uint(fabsf(d)))));
        }
      }
      atomicAdd(usum + j, acc);
    }
  }
  if (MODE == 1) {
    amax = __reduce_max_sync(0xffffffffu, amax);
    if (lane == 0 && amax != 0u) atomicMax(absmax, amax);
  }
}

// ---- the same row pass for the h3 engine, split in two so that T = [E; U; D] is written ONCE, directly as the fp16 pair:
// (1) per-sample scalars alpha_b, beta_b and a bound on |e|, |u|, |d| over the whole batch (|u_b| <= alpha |w|max +
//     |beta| |d|max per row, |e_b| <= |d|max + that; loose by at most 2x, i.e. one bit of fp16 range) - a warp per row;
// (2) e, u, d from X, W and the scalars, split with the scale of that bound, plus the column sums of u.
// 22 B D bytes of traffic instead of 24 B D (pass) + 24 B D (separate split of an fp32 T).
__global__ void __launch_bounds__(256) gsm_rowscal_kernel(const float* __restrict__ X, long long ldx,
                                                          const float* __restrict__ G, long long ldg,
                                                          const float* __restrict__ W, long long ldw,
                                                          const float* __restrict__ mu, float* __restrict__ ab, int B, int D,
                                                          unsigned* __restrict__ tbound) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const float* x = X + static_cast<long long>(b) * ldx;
  const float* g = G + static_cast<long long>(b) * ldg;
  const float* w = W + static_cast<long long>(b) * ldw;
  const bool vec = ((D & 3) == 0) && ((ldx & 3) == 0) && ((ldg & 3) == 0) && ((ldw & 3) == 0) &&
                   (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(W) |
                      reinterpret_cast<uintptr_t>(mu)) & 15) == 0);
  float vSv = 0.0f, mu_v = 0.0f, wmax = 0.0f, dmax = 0.0f;
  if (vec) {
    for (int j = lane * 4; j < D; j += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(x + j);
      const float4 gv = *reinterpret_cast<const float4*>(g + j);
      const float4 wv = *reinterpret_cast<const float4*>(w + j);
      const float4 mv = *reinterpret_cast<const float4*>(mu + j);
      const float d0 = mv.x - xv.x, d1 = mv.y - xv.y, d2 = mv.z - xv.z, d3 = mv.w - xv.w;
      vSv += wv.x * gv.x + wv.y * gv.y + wv.z * gv.z + wv.w * gv.w;
      mu_v += d0 * gv.x + d1 * gv.y + d2 * gv.z + d3 * gv.w;
      wmax = fmaxf(wmax, fmaxf(fmaxf(fabsf(wv.x), fabsf(wv.y)), fmaxf(fabsf(wv.z), fabsf(wv.w))));
      dmax = fmaxf(dmax, fmaxf(fmaxf(fabsf(d0), fabsf(d1)), fmaxf(fabsf(d2), fabsf(d3))));
    }
  } else {
    for (int j = lane; j < D; j += 32) {
      const float d = mu[j] - x[j];
      vSv += w[j] * g[j];
      mu_v += d * g[j];
      wmax = fmaxf(wmax, fabsf(w[j]));
      dmax = fmaxf(dmax, fabsf(d));
    }
  }
  vSv = warp_sum(vSv);
  mu_v = warp_sum(mu_v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) {
    const float rho = 0.5f * sqrtf(1.0f + 4.0f * (vSv + mu_v * mu_v)) - 0.5f;
    const float alpha = 1.0f / (1.0f + rho);
    const float beta = -alpha * (1.0f + (vSv - mu_v) / (1.0f + rho + mu_v));
    ab[b] = alpha;
    ab[B + b] = beta;
    // NaN / Inf anywhere in the row makes the bound NaN / Inf, which (as a bit pattern) outranks every finite value
    const float bound = dmax + fabsf(alpha) * wmax + fabsf(beta) * dmax;
    atomicMax(tbound, __float_as_uint(fabsf(bound)));
  }
}

__global__ void __launch_bounds__(RP_THREADS) gsm_rowwrite_h3_kernel(const float* __restrict__ X, long long ldx,
                                                                     const float* __restrict__ W, long long ldw,
                                                                     const float* __restrict__ mu, const float* __restrict__ ab,
                                                                     const unsigned* __restrict__ tbound,
                                                                     float* __restrict__ scale_out, __half* __restrict__ Thi,
                                                                     __half* __restrict__ Tlo, long long ldt,
                                                                     float* __restrict__ usum, int B, int D, int rows_per_cta) {
  const float sc = h3_scale_from_absmax(*tbound, 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = sc;
  const int row0 = blockIdx.x * rows_per_cta;
  const int nrows = min(rows_per_cta, B - row0);
  const bool vec = ((D & 3) == 0) && ((ldx & 3) == 0) && ((ldw & 3) == 0) && ((ldt & 3) == 0) &&
                   (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(mu)) & 15) == 0) &&
                   (((reinterpret_cast<uintptr_t>(Thi) | reinterpret_cast<uintptr_t>(Tlo)) & 7) == 0);
  auto put4 = [&](long long row, int j, const float4 v) {
    __half h[4], l[4];
    h3_split1(v.x, sc, h[0], l[0]);
    h3_split1(v.y, sc, h[1], l[1]);
    h3_split1(v.z, sc, h[2], l[2]);
    h3_split1(v.w, sc, h[3], l[3]);
    *reinterpret_cast<uint2*>(Thi + row * ldt + j) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(Tlo + row * ldt + j) = *reinterpret_cast<const uint2*>(l);
  };
  if (vec) {
    for (int j = threadIdx.x * 4; j < D; j += RP_THREADS * 4) {
      const float4 mv = *reinterpret_cast<const float4*>(mu + j);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r0 = 0; r0 < nrows; r0 += 4) {
        float4 xv[4], wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long b = row0 + min(r0 + q, nrows - 1);
          xv[q] = *reinterpret_cast<const float4*>(X + b * ldx + j);
          wv[q] = *reinterpret_cast<const float4*>(W + b * ldw + j);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = r0 + q;
          if (r >= nrows) break;
          const long long b = row0 + r;
          const float al = ab[b], be = ab[B + b];
          float4 d, u, e;
          d.x = mv.x - xv[q].x; d.y = mv.y - xv[q].y; d.z = mv.z - xv[q].z; d.w = mv.w - xv[q].w;
          u.x = al * wv[q].x + be * d.x; u.y = al * wv[q].y + be * d.y; u.z = al * wv[q].z + be * d.z; u.w = al * wv[q].w + be * d.w;
          e.x = d.x + u.x; e.y = d.y + u.y; e.z = d.z + u.z; e.w = d.w + u.w;
          acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
          put4(b, j, e);
          put4(b + B, j, u);
          put4(b + 2LL * B, j, d);
        }
      }
      atomicAdd(usum + j + 0, acc.x);
      atomicAdd(usum + j + 1, acc.y);
      atomicAdd(usum + j + 2, acc.z);
      atomicAdd(usum + j + 3, acc.w);
    }
  } else {
    for (int j = threadIdx.x; j < D; j += RP_THREADS) {
      const float m = mu[j];
      float acc = 0.0f;
      for (int r = 0; r < nrows; ++r) {
        const long long b = row0 + r;
        const float d = m - X[b * ldx + j];
        const float u = ab[b] * W[b * ldw + j] + ab[B + b] * d;
        acc += u;
        h3_split1(d + u, sc, Thi[b * ldt + j], Tlo[b * ldt + j]);
        h3_split1(u, sc, Thi[(b + B) * ldt + j], Tlo[(b + B) * ldt + j]);
        h3_split1(d, sc, Thi[(b + 2LL * B) * ldt + j], Tlo[(b + 2LL * B) * ldt + j]);
      }
      atomicAdd(usum + j, acc);
    }
  }
}

// 16 rows per CTA amortise the column-sum atomics; batch shards too small to give every SM a few CTAs that way use fewer
static inline int rowpass_rows_per_cta(int B) {
  int r = RP_ROWS;
  while (r > 2 && (B + r - 1) / r < 296) r >>= 1;
  return r;
}

// out[j] = a[j] + scale * s[j]
__global__ void vec_axpy_kernel(const float* __restrict__ a, const float* __restrict__ s, float scale,
                                float* __restrict__ out, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (a ? a[j] : 0.0f) + scale * s[j];
}

// C[i,j] = A[i,j] + Bm[i,j]   (n x n; the multi-GPU path applies the all-reduced statistics with it)
__global__ void mat_add_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb,
                               float* __restrict__ C, long long ldc, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < n) C[i * ldc + j] = A[i * lda + j] + Bm[i * ldb + j];
}

// ---------------------------------------------------------------------------------------------- Philox N(0,1)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10: counter = (element group, offset), key = seed.  Four uniforms -> four normals (Box-Muller).
__global__ void philox_normal_kernel(float* __restrict__ Z, long long ldz, int B, int D, unsigned long long seed,
                                     unsigned long long offset) {
  const int groups_per_row = (D + 3) / 4;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(B) * groups_per_row) return;
  const int b = static_cast<int>(gid / groups_per_row);
  const int j = static_cast<int>(gid % groups_per_row) * 4;
  uint32_t c[4] = {static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), static_cast<uint32_t>(offset),
                   static_cast<uint32_t>(offset >> 32)};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u0 = (static_cast<float>(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = (static_cast<float>(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = (static_cast<float>(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
  float* row = Z + static_cast<long long>(b) * ldz;
  if (j + 3 < D && (ldz & 3) == 0) {
    *reinterpret_cast<float4*>(row + j) = make_float4(z[0], z[1], z[2], z[3]);
  } else {
    for (int t = 0; t < 4 && j + t < D; ++t) row[j + t] = z[t];
  }
}

// ---------------------------------------------------------------------------------------------- orchestration

static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

int philox_normal(cudaStream_t stream, float* Z, long long ldz, int B, int D, unsigned long long seed,
                  unsigned long long offset) {
  if (!Z || B <= 0 || D <= 0 || ldz < D) return GSMVI_EINVAL;
  const long long n = static_cast<long long>(B) * ((D + 3) / 4);
  philox_normal_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(Z, ldz, B, D, seed, offset);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int tf32_split(cudaStream_t stream, const float* A, long long lda, float* Hi, float* Lo, long long ldo, int rows, int cols) {
  if (!A || !Hi || !Lo || rows <= 0 || cols <= 0) return GSMVI_EINVAL;
  tf32_split_kernel<<<dim3((cols / 4 + 256) / 256, rows), 256, 0, stream>>>(A, lda, Hi, Lo, ldo, rows, cols);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int sample_mvn(cudaStream_t stream, const float* mu, const float* L, const float* L_lo, long long ldl, const float* Z,
               long long ldz, float* X, long long ldx, int B, int D, int npass) {
  // X[b,i] = mu[i] + sum_{k<=i} Z[b,k] L[i,k]
  GemmOpts o;
  o.npass = npass;
  o.bias_n = mu;
  o.krange = KR_B_LOWER;
  MatView vz{Z, B, D, ldz}, vl{L, D, D, ldl, L_lo};
  return launch_gemm_tf32(stream, B, D, D, vz, vl, X, ldx, o);
}

int gauss_score(cudaStream_t stream, const float* X, long long ldx, const float* P, const float* P_lo, long long ldp,
                const float* c, float* G, long long ldg, int B, int D, int npass) {
  // G = -(X - m) P = -X P + c,  c = P m   (P symmetric: P[n,k] read K-major as-is)
  GemmOpts o;
  o.npass = npass;
  o.alpha = -1.0f;
  o.bias_n = c;
  MatView vx{X, B, D, ldx}, vp{P, D, D, ldp, P_lo};
  return launch_gemm_tf32(stream, B, D, D, vx, vp, G, ldg, o);
}

size_t gsm_update_workspace_bytes(int B, int D) {
  const long long ldw = round_up(D, 32);
  // W [B x ldw] + T_hi, T_lo = split [E; U; D] [2 x 3B x ldw] + usum [ldw]
  return static_cast<size_t>((7LL * B + 1) * ldw) * sizeof(float);
}

int gsm_update(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const float* mu,
               const float* Sigma, const float* Sigma_hi, const float* Sigma_lo, long long lds, float* mu_out,
               float* Sigma_out, long long ldso, int B, int D, int B_total, int mode, float* workspace, int npass) {
  if (!X || !G || !mu || !Sigma || !mu_out || !Sigma_out || !workspace || B <= 0 || D <= 0 || B_total < B)
    return GSMVI_EINVAL;
  const long long ldw = round_up(D, 32);
  float* W = workspace;
  float* T = W + static_cast<long long>(B) * ldw;
  float* Tlo = T + 3LL * B * ldw;
  float* usum = Tlo + 3LL * B * ldw;
  cudaError_t e = cudaMemsetAsync(usum, 0, ldw * sizeof(float), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  // (i) W = G Sigma0   (Sigma0 symmetric)
  {
    GemmOpts o;
    o.npass = npass;
    // pre-split pair (hi, lo) if given, else the raw matrix split in-kernel
    const bool pre = Sigma_hi != nullptr && Sigma_lo != nullptr;
    MatView vg{G, B, D, ldg}, vs{pre ? Sigma_hi : Sigma, D, D, lds, pre ? Sigma_lo : nullptr};
    int rc = launch_gemm_tf32(stream, B, D, D, vg, vs, W, ldw, o);
    if (rc != GSMVI_OK) return rc;
  }
  // (ii) row pass
  const int rpc = rowpass_rows_per_cta(B);
  gsm_rowpass_kernel<0><<<(B + rpc - 1) / rpc, RP_THREADS, 0, stream>>>(X, ldx, G, ldg, W, ldw, mu, T, Tlo, ldw, usum, B, D, nullptr, rpc);
  e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  // (iii) Sigma_out = [Sigma0] - (E^T U + U^T D) / B_total : rows of T are K, so both operands are MN-major views
  {
    GemmOpts o;
    o.npass = npass;
    o.a_mn = o.b_mn = true;
    o.alpha = -1.0f / static_cast<float>(B_total);
    o.tri = true;
    o.mirror = true;
    if (mode == 0) {
      o.beta = 1.0f;
      o.Cin = Sigma;
      o.ldcin = lds;
    }
    // both operands arrive pre-split (T_hi, T_lo from the row pass): no in-kernel conversion for 3-pass modes
    MatView va{T, 2LL * B, D, ldw, Tlo}, vb{T + static_cast<long long>(B) * ldw, 2LL * B, D, ldw, Tlo + static_cast<long long>(B) * ldw};
    int rc = launch_gemm_tf32(stream, D, D, 2 * B, va, vb, Sigma_out, ldso, o);
    if (rc != GSMVI_OK) return rc;
  }
  // mu_out = [mu0 +] usum / B_total
  vec_axpy_kernel<<<(D + 255) / 256, 256, 0, stream>>>(mode == 0 ? mu : nullptr, usum, 1.0f / static_cast<float>(B_total),
                                                      mu_out, D);
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int gsm_apply_stats(cudaStream_t stream, const float* Sigma, long long lds, const float* dSigma, long long ldd,
                    const float* mu, const float* dmu, float* Sigma_out, long long ldso, float* mu_out, int D) {
  mat_add_kernel<<<dim3((D + 255) / 256, D), 256, 0, stream>>>(Sigma, lds, dSigma, ldd, Sigma_out, ldso, D);
  vec_axpy_kernel<<<(D + 255) / 256, 256, 0, stream>>>(mu, dmu, 1.0f, mu_out, D);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}


// ================================================================================================ scaled 3xFP16 path
// The same iteration on the h3 engine (h3_gemm.cuh): every GEMM operand is an fp16 (hi, lo) pair with a power-of-two
// scale.  Producers record max |value| through their epilogue (one atomicMax per warp) and a split pass turns the fp32
// result into the next GEMM's operand; Philox draws are written split directly (|z| < 2^3 by construction).

// Philox4x32-10 normals written as the fp16 pair with the fixed scale 2^11 (|z| <= sqrt(-2 ln 2^-25) = 5.9 < 2^3).
constexpr float Z_H3_SCALE = 2048.0f;
__global__ void philox_normal_h3_kernel(__half* __restrict__ Zhi, __half* __restrict__ Zlo, long long ldz, int B, int D,
                                        unsigned long long seed, unsigned long long offset,
                                        const unsigned long long* __restrict__ offset_dev, float* __restrict__ scale_out) {
  if (offset_dev) offset = *offset_dev;  // counter kept on the device: the launch can be replayed from a CUDA graph
  const int groups_per_row = (D + 3) / 4;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid == 0) *scale_out = Z_H3_SCALE;
  if (gid >= static_cast<long long>(B) * groups_per_row) return;
  const int b = static_cast<int>(gid / groups_per_row);
  const int j = static_cast<int>(gid % groups_per_row) * 4;
  uint32_t c[4] = {static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), static_cast<uint32_t>(offset),
                   static_cast<uint32_t>(offset >> 32)};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u0 = (static_cast<float>(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = (static_cast<float>(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = (static_cast<float>(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
  __half h[4], l[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) h3_split1(z[t], Z_H3_SCALE, h[t], l[t]);
  __half* hrow = Zhi + static_cast<long long>(b) * ldz;
  __half* lrow = Zlo + static_cast<long long>(b) * ldz;
  if (j + 3 < D && (ldz & 3) == 0) {
    *reinterpret_cast<uint2*>(hrow + j) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lrow + j) = *reinterpret_cast<const uint2*>(l);
  } else {
    for (int t = 0; t < 4 && j + t < D; ++t) { hrow[j + t] = h[t]; lrow[j + t] = l[t]; }
  }
}

int philox_normal_h3(cudaStream_t stream, const H3Operand& Z, int B, int D, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* offset_dev) {
  if (!Z.hi || !Z.lo || !Z.scale || B <= 0 || D <= 0 || Z.ld < D) return GSMVI_EINVAL;
  const long long n = static_cast<long long>(B) * ((D + 3) / 4);
  philox_normal_h3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<__half*>(Z.hi), static_cast<__half*>(Z.lo), Z.ld, B, D, seed, offset, offset_dev, Z.scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

static inline HView hview(const H3Operand& o, long long rows, long long cols) {
  return HView{static_cast<const __half*>(o.hi), static_cast<const __half*>(o.lo), rows, cols, o.ld, o.scale};
}

int sample_mvn_h3(cudaStream_t stream, const float* mu, const H3Operand& L, const H3Operand& Z, float* X, long long ldx,
                  unsigned* absmax_x, int B, int D, const H3Operand* Xsplit) {
  // X[b,i] = mu[i] + sum_{k<=i} Z[b,k] L[i,k]
  H3Opts o;
  o.bias_n = mu;
  o.krange = KR_B_LOWER;
  o.absmax_out = absmax_x;
  if (Xsplit) {  // the epilogue also writes X as the fp16 pair with the caller's (a-priori) scale
    o.split_hi = static_cast<__half*>(Xsplit->hi);
    o.split_lo = static_cast<__half*>(Xsplit->lo);
    o.split_ld = Xsplit->ld;
    o.split_scale = Xsplit->scale;
  }
  return launch_gemm_h3(stream, B, D, D, hview(Z, B, D), hview(L, D, D), X, ldx, o);
}

int gauss_score_h3(cudaStream_t stream, const H3Operand& X, const H3Operand& P, const float* c, float* G, long long ldg,
                   unsigned* absmax_g, int B, int D, const H3Operand* Gsplit) {
  // G = -X P + c,  c = P m   (P symmetric: P[n,k] read K-major as-is)
  H3Opts o;
  o.alpha = -1.0f;
  o.bias_n = c;
  o.absmax_out = absmax_g;
  if (Gsplit) {
    o.split_hi = static_cast<__half*>(Gsplit->hi);
    o.split_lo = static_cast<__half*>(Gsplit->lo);
    o.split_ld = Gsplit->ld;
    o.split_scale = Gsplit->scale;
  }
  return launch_gemm_h3(stream, B, D, D, hview(X, B, D), hview(P, D, D), G, ldg, o);
}

size_t gsm_update_h3_workspace_bytes(int B, int D) {
  const long long ldw = round_up(D, 32);
  // W [B x ldw] fp32 + T = [E; U; D] [3B x ldw] fp32 + usum [ldw] fp32 + 32 floats of scalars + T_hi, T_lo [3B x ldw] fp16
  return static_cast<size_t>((4LL * B + 1) * ldw + 32) * sizeof(float) + static_cast<size_t>(6LL * B * ldw) * sizeof(__half);
}

// GSMVI_UPDATE_TIMING=1: CUDA-event stamps between the launches of gsm_update_h3 (synchronises; diagnostics only)
struct UpdTimer {
  bool on;
  cudaStream_t st;
  cudaEvent_t ev[10];
  const char* name[10];
  int n = 0;
  explicit UpdTimer(cudaStream_t s) : st(s) {
    static int v = -1;
    if (v < 0) v = getenv("GSMVI_UPDATE_TIMING") ? 1 : 0;
    on = v == 1;
  }
  void mark(const char* what) {
    if (!on || n >= 10) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void report(int rank) {
    if (!on) return;
    cudaStreamSynchronize(st);
    static int calls = 0;
    if (rank == 0 && (++calls % 16) == 0) {
      fprintf(stderr, "[gsmvi update timing, ms]");
      for (int i = 1; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        fprintf(stderr, " %s %.3f", name[i], ms);
      }
      fprintf(stderr, "\n");
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

struct FusedComm {
  float* const* base;
  float* own_base;
  const gsmvi_comm_layout* lay;
  int rank, world, cur;
  unsigned step;
};

static int gsm_update_h3_impl(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg,
                              const H3Operand& Gh, const float* mu, const float* Sigma, long long lds, const H3Operand& Sh,
                              float* mu_out, float* Sigma_out, long long ldso, unsigned* absmax_sout, int B, int D, int B_total,
                              int mode, void* workspace, const FusedComm* fc) {
  if (!X || !G || !mu || !mu_out || !workspace || B <= 0 || D <= 0 || B_total < B) return GSMVI_EINVAL;
  if (!fc && (!Sigma || !Sigma_out)) return GSMVI_EINVAL;
  const long long ldw = round_up(D, 32);
  float* W = static_cast<float*>(workspace);
  float* T = W + static_cast<long long>(B) * ldw;
  float* usum = T + 3LL * B * ldw;
  float* scal = usum + ldw;  // [0] = |T| max (bit pattern), [1] = T scale
  __half* Thi = reinterpret_cast<__half*>(scal + 32);
  __half* Tlo = Thi + 3LL * B * ldw;
  UpdTimer tmr(stream);
  tmr.mark("start");
  cudaError_t e = cudaMemsetAsync(usum, 0, (ldw + 32) * sizeof(float), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  int rc;
  // (i) W = G Sigma0   (Sigma0 symmetric)
  {
    H3Opts o;
    if ((rc = launch_gemm_h3(stream, B, D, D, hview(Gh, B, D), hview(Sh, D, D), W, ldw, o)) != GSMVI_OK) return rc;
  }
  tmr.mark("W");
  // (ii) row pass in two launches: per-sample scalars + a bound on |T|, then T = [E; U; D] written once as the fp16 pair
  //      (the fp32 T area only holds the 2B scalars), with the column sums of U
  {
    float* ab = T;
    gsm_rowscal_kernel<<<(B + 7) / 8, 256, 0, stream>>>(X, ldx, G, ldg, W, ldw, mu, ab, B, D, reinterpret_cast<unsigned*>(scal));
    tmr.mark("rowscal");
    const int rpc = rowpass_rows_per_cta(B);
    gsm_rowwrite_h3_kernel<<<(B + rpc - 1) / rpc, RP_THREADS, 0, stream>>>(X, ldx, W, ldw, mu, ab, reinterpret_cast<const unsigned*>(scal),
                                                                         scal + 1, Thi, Tlo, ldw, usum, B, D, rpc);
    if ((e = cudaGetLastError()) != cudaSuccess) return static_cast<int>(e);
    tmr.mark("rowwrite");
  }
  // (iii) Sigma_out = [Sigma0] - (E^T U + U^T D) / B_total : rows of T are K, so both operands are MN-major views
  {
    H3Opts o;
    o.a_mn = o.b_mn = true;
    o.alpha = -1.0f / static_cast<float>(B_total);
    o.tri = true;
    o.mirror = true;
    o.absmax_out = absmax_sout;
    if (fc) {
      // multi-GPU: the epilogue pushes every partial tile to its owner rank (reduce-scatter fused into the GEMM)
      o.mirror = false;
      o.absmax_out = nullptr;
      o.push_base = fc->base;
      o.push_stage_off = fc->lay->stage_off;
      o.push_cnt_off = fc->lay->cnt_off;
      o.push_rank = fc->rank;
      o.push_world = fc->world;
      o.push_tpo = fc->lay->tpo;
    } else if (mode == 0) {
      o.beta = 1.0f;
      o.Cin = Sigma;
      o.ldcin = lds;
    }
    HView va{Thi, Tlo, 2LL * B, D, ldw, scal + 1};
    HView vb{Thi + static_cast<long long>(B) * ldw, Tlo + static_cast<long long>(B) * ldw, 2LL * B, D, ldw, scal + 1};
    if ((rc = launch_gemm_h3(stream, D, D, 2 * B, va, vb, Sigma_out, ldso, o)) != GSMVI_OK) return rc;
  }
  tmr.mark("covGEMM");
  if (fc) {  // owners reduce + broadcast the new Sigma tiles, mean increments exchanged, mu_out formed
    rc = comm_reduce_broadcast(stream, fc->base, fc->own_base, *fc->lay, fc->rank, fc->world, D, fc->cur, fc->step, usum,
                               1.0f / static_cast<float>(B_total), mu, mu_out);
    tmr.mark("reduce+finalize+mirror");
    tmr.report(fc->rank);
    return rc;
  }
  vec_axpy_kernel<<<(D + 255) / 256, 256, 0, stream>>>(mode == 0 ? mu : nullptr, usum, 1.0f / static_cast<float>(B_total),
                                                      mu_out, D);
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

// ---- accept / revert on the device (gsmvi/gsm.py:125-129).  The host always flips its buffer pointers as if the update
// had been accepted; when the goodness check raised *bad, this kernel copies the previous state over the rejected
// proposal, so the flipped pointers again name the old (mu, Sigma, L) - a rejected update costs one state copy, an
// accepted one a flag read per CTA, and no iteration waits for the host to read anything back.
struct CommitArgs {
  const int* bad;
  const int* bad2;  // optional second flag (BaM: the solve's own PD flag)
  int* status;      // [0] += 1 per rejected update, [1] <- 1 if accepted else 0
  int n;
  const void* src[GSMVI_COMMIT_MAX];
  void* dst[GSMVI_COMMIT_MAX];
  long long bytes[GSMVI_COMMIT_MAX];
};

__global__ void __launch_bounds__(256) gsm_commit_kernel(const CommitArgs a) {
  const bool bad = (*a.bad != 0) || (a.bad2 != nullptr && *a.bad2 != 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (bad) a.status[0] += 1;
    a.status[1] = bad ? 0 : 1;
  }
  if (!bad) return;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int r = 0; r < a.n; ++r) {
    const long long nb = a.bytes[r];
    if (((reinterpret_cast<uintptr_t>(a.src[r]) | reinterpret_cast<uintptr_t>(a.dst[r]) | static_cast<uintptr_t>(nb)) & 15) == 0) {
      const int4* s4 = static_cast<const int4*>(a.src[r]);
      int4* d4 = static_cast<int4*>(a.dst[r]);
      for (long long i = tid; i < nb / 16; i += nth) d4[i] = s4[i];
    } else {
      const unsigned* s1 = static_cast<const unsigned*>(a.src[r]);
      unsigned* d1 = static_cast<unsigned*>(a.dst[r]);
      for (long long i = tid; i < nb / 4; i += nth) d1[i] = s1[i];
    }
  }
}

int gsm_commit(cudaStream_t stream, const int* bad, const int* bad2, int n, const void* const* src, void* const* dst,
               const long long* bytes, int* status) {
  if (!bad || !status || n < 0 || n > GSMVI_COMMIT_MAX || (n > 0 && (!src || !dst || !bytes))) return GSMVI_EINVAL;
  CommitArgs a;
  a.bad = bad;
  a.bad2 = bad2;
  a.status = status;
  a.n = n;
  long long most = 0;
  for (int r = 0; r < n; ++r) {
    if (!src[r] || !dst[r] || bytes[r] < 0 || (bytes[r] & 3)) return GSMVI_EINVAL;
    a.src[r] = src[r];
    a.dst[r] = dst[r];
    a.bytes[r] = bytes[r];
    most = bytes[r] > most ? bytes[r] : most;
  }
  // the common (accepted) case is one flag read per CTA: a grid that is large enough to stream a 64 MiB state at HBM
  // speed when it does copy, small enough to retire in a microsecond when it does not
  int ctas = static_cast<int>((most / 16 + 255) / 256);
  ctas = ctas < 1 ? 1 : (ctas > 592 ? 592 : ctas);
  gsm_commit_kernel<<<ctas, 256, 0, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int gsm_update_h3(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const H3Operand& Gh,
                  const float* mu, const float* Sigma, long long lds, const H3Operand& Sh, float* mu_out, float* Sigma_out,
                  long long ldso, unsigned* absmax_sout, int B, int D, int B_total, int mode, void* workspace) {
  return gsm_update_h3_impl(stream, X, ldx, G, ldg, Gh, mu, Sigma, lds, Sh, mu_out, Sigma_out, ldso, absmax_sout, B, D, B_total,
                            mode, workspace, nullptr);
}

int gsm_update_h3_fused(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const H3Operand& Gh,
                        const float* mu, const H3Operand& Sh, float* mu_out, float* const* peer_base, float* own_base,
                        const gsmvi_comm_layout& lay, int rank, int world, int cur, unsigned step, int B, int D, int B_total,
                        void* workspace) {
  if (!peer_base || !own_base || world < 1 || rank < 0 || rank >= world || (cur != 0 && cur != 1)) return GSMVI_EINVAL;
  FusedComm fc{peer_base, own_base, &lay, rank, world, cur, step};
  return gsm_update_h3_impl(stream, X, ldx, G, ldg, Gh, mu, nullptr, 0, Sh, mu_out, nullptr, 0, nullptr, B, D, B_total, 1,
                            workspace, &fc);
}

}  // namespace gsmvi
