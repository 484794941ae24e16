// Blocked right-looking Cholesky  Sigma = L L^T  (fp32) with a device-side "is positive definite" flag.
//
// Replaces two host round trips of the reference per iteration: the factorisation inside the sampler
// np.random.multivariate_normal (gsmvi/gsm.py:119, bam.py:193 - an SVD there) and the PD check
// np.linalg.cholesky in _check_goodness (gsm.py:136-150, bam.py:219-233).  One factorisation serves both: the flag
// accepts/rejects the update, and on accept L is the next iteration's sampling factor.
//
// Per 128-column panel, two launches:  (1) a multi-CTA panel kernel - CTA 0 factors the 128x128 diagonal block in shared
// memory / registers and publishes it through a release flag, the other CTAs prefetch their rows of the panel, acquire
// the flag and solve L21 L11^T = A21 one row per thread in registers;  (2) the SYRK trailing update
// A22 -= L21 L21^T on lower tiles, on the tcgen05 3xTF32 GEMM.
#include "dev_once.cuh"
#include "potrf.cuh"
#include "chol_block.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace gsmvi {

constexpr int NB = 128;

// L (lower, incl. diagonal) <- lower triangle of A; strict upper triangle of L <- 0.
__global__ void tril_copy_kernel(const float* __restrict__ A, long long lda, float* __restrict__ L, long long ldl, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < n) L[static_cast<long long>(i) * ldl + j] = (j <= i) ? A[static_cast<long long>(i) * lda + j] : 0.0f;
}

// Shared-memory building blocks of the panel kernel: the 128x128 block is processed in four 32-column sub-panels:
//   (1) warp 0 factors the 32x32 diagonal block, one row per lane in registers, pivots broadcast with shuffles;
//   (2) a thread per row below solves its 32 panel entries against that block by forward substitution (registers);
//   (3) all threads apply the rank-32 update to the trailing lower triangle in 4x4 register tiles.
__device__ long long g_pt[24];
#define PT(i) do { if (threadIdx.x == 0 && gridDim.x > 8) g_pt[i] = clock64(); } while (0)
constexpr int RPC = 32;          // rows of the panel per TRSM CTA (eight threads per row)
constexpr int TPR = 256 / RPC;
constexpr int DS = NB + 4;  // shared-memory leading dimension: rows stay 16-byte aligned, quarter-warps hit distinct banks

// acc[r][c] += sum_k Arows[r * lda_][k] * Brows[c * ldb_][k]   (both row-major over k; kbeg, kend multiples of 4).
// Callers pick the row strides so that the lanes of a quarter-warp touch CONSECUTIVE rows: with a leading dimension of
// 4 (mod 32) words, eight consecutive rows cover all 32 banks for a float4 access (conflict-free).
__device__ __forceinline__ void tile4x4_mac(const float* __restrict__ Arows, int lda_, const float* __restrict__ Brows,
                                            int ldb_, int kbeg, int kend, float (&acc)[4][4]) {
  for (int k = kbeg; k < kend; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(Arows + r * lda_ + k);
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(Brows + c * ldb_ + k);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        acc[r][c] += a[r].x * b[c].x + a[r].y * b[c].y + a[r].z * b[c].z + a[r].w * b[c].w;
  }
}

// Factor the n x n (n <= 128) block held in shared memory `s` (leading dimension DS, lower triangle valid, identity
// padded) in place; dinv[j] <- 1 / L[j][j].  Returns (in every thread) whether a pivot was non-positive / non-finite.
__device__ __forceinline__ bool factor_block_smem(float* s, float* dinv, int* bad_smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int p = 0; p < NB / 32; ++p) {
    const int c0 = 32 * p;
    // ---- (1) 32x32 diagonal block, warp 0, row `lane` in registers
    if (warp == 0) {
      float row[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) row[k] = s[(c0 + lane) * DS + c0 + k];
      int isbad = 0;
      chol32_step<0>(row, lane, dinv + c0, isbad);
#pragma unroll
      for (int k = 0; k < 32; ++k) s[(c0 + lane) * DS + c0 + k] = (k <= lane) ? row[k] : 0.0f;
      if (isbad && lane == 0) *bad_smem = 1;
    }
    __syncthreads();
    // ---- (2) rows below: x L11^T = a  by forward substitution, one thread per row
    if (tid < NB && tid >= c0 + 32) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 t = *reinterpret_cast<const float4*>(s + tid * DS + c0 + k);
        v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float acc0 = v[j], acc1 = 0.0f;
        const float* lrow = s + (c0 + j) * DS + c0;  // broadcast reads
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) {
          acc0 -= v[k] * lrow[k];
          acc1 -= v[k + 1] * lrow[k + 1];
        }
        if (j & 1) acc0 -= v[j - 1] * lrow[j - 1];
        v[j] = (acc0 + acc1) * dinv[c0 + j];
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(s + tid * DS + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
    __syncthreads();
    // ---- (3) trailing update: S[i][k] -= sum_c P[i][c] P[k][c] on the lower triangle of the (NB-c0-32)^2 block.
    // Interleaved 4x4 tiles: thread (ti, tj) owns rows m0 + ti + r*mt and columns m0 + tj + c*mt, so neighbouring
    // lanes read neighbouring rows (bank-conflict-free float4 loads).  Elements above the diagonal are skipped.
    const int m0 = c0 + 32, mt = (NB - m0) / 4;  // mt = 24, 16, 8, 0
    for (int t = tid; t < mt * mt; t += 256) {
      const int ti = t % mt, tj = t / mt;
      float acc[4][4] = {};
      tile4x4_mac(s + (m0 + ti) * DS, mt * DS, s + (m0 + tj) * DS, mt * DS, c0, c0 + 32, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = m0 + ti + r * mt, k = m0 + tj + c * mt;
          if (k <= i) s[i * DS + k] -= acc[r][c];
        }
    }
    __syncthreads();
  }
  return *bad_smem != 0;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Panel kernel: one launch factors the 128-wide panel starting at column j0 of the n x n matrix L (in place).
//   CTA 0            : Cholesky of the nb x nb diagonal block (shared memory / registers), written back with explicit
//                      zeros above the diagonal, then publishes `epoch` through *ready (release).
//   CTA t = 1, 2, ...: TRSM of rows r0 = j0 + nb + 32 (t-1) ...: prefetches its 32 x 128 tile of A21 into shared memory
//                      while CTA 0 works, acquires the flag, loads L11, and solves X L11^T = A21 with eight threads per
//                      row in registers (fp32 FMAs: no explicit inverse, no tensor-core rounding), 32 columns at a time.
// All CTAs must be co-resident for the spin-wait (1 + rest/32 CTAs of one per SM: n <= 4700; larger n would need a
// larger RPC).
// A non-positive or non-finite pivot sets *flag (bit 0); the panel is then garbage and the caller discards L.
__global__ void __launch_bounds__(256, 1) potrf_panel_kernel(float* __restrict__ Lm, long long ld, int n, int j0, int nb,
                                                             int* __restrict__ flag, unsigned* __restrict__ ready,
                                                             unsigned epoch) {
  extern __shared__ __align__(16) float sm[];
  float* s = sm;            // [NB][DS]  diagonal block (CTA 0) / L11 (TRSM CTAs)
  float* at = sm + NB * DS; // [NB][DS]  TRSM CTAs: their rows of A21 -> L21
  __shared__ float dinv[NB];
  __shared__ int bad;
  const int tid = threadIdx.x;
  float* a11 = Lm + static_cast<long long>(j0) * ld + j0;
  const bool vec_ok = (nb == NB) && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a11) & 15) == 0);
  if (tid == 0) bad = 0;

  if (blockIdx.x == 0) {
    PT(0);
    // ---------------- diagonal block
    if (vec_ok) {
      float4 v[16];  // 4096 float4: 16 loads per thread issued back to back, then stored
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
        v[e] = (j4 <= i) ? *reinterpret_cast<const float4*>(a11 + static_cast<long long>(i) * ld + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
        float4 t = v[e];
        if (j4 + 1 > i) t.y = 0.f;
        if (j4 + 2 > i) t.z = 0.f;
        if (j4 + 3 > i) t.w = 0.f;
        *reinterpret_cast<float4*>(s + i * DS + j4) = t;
      }
    } else {
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 7, j = idx & (NB - 1);
        s[i * DS + j] = (i < nb && j <= i) ? a11[static_cast<long long>(i) * ld + j] : ((i == j) ? 1.0f : 0.0f);
      }
    }
    __syncthreads();
    PT(1);
    const bool isbad = factor_block_smem(s, dinv, &bad);
    PT(2);
    if (vec_ok) {
#pragma unroll 4
      for (int e = 0; e < 16; ++e) {
        const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
        *reinterpret_cast<float4*>(a11 + static_cast<long long>(i) * ld + j4) = *reinterpret_cast<const float4*>(s + i * DS + j4);
      }
    } else {
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 7, j = idx & (NB - 1);
        if (i < nb && j < nb) a11[static_cast<long long>(i) * ld + j] = (j <= i) ? s[i * DS + j] : 0.0f;
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      if (isbad) atomicOr(flag, 1);
      st_release_u32(ready, epoch);
    }
    PT(3);
    return;
  }

  // ---------------- TRSM rows (nb == NB whenever rows below exist)
  if (blockIdx.x == 1) PT(8);
  const int r0 = j0 + nb + (blockIdx.x - 1) * RPC;
  const int rows = min(RPC, n - r0);
  float* a21 = Lm + static_cast<long long>(r0) * ld + j0;
  for (int q = tid; q < RPC * 32; q += 256) {  // prefetch this CTA's rows while CTA 0 factors
    const int i = q >> 5, j4 = (q & 31) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < rows) {
      if (vec_ok) t = *reinterpret_cast<const float4*>(a21 + static_cast<long long>(i) * ld + j4);
      else { t.x = a21[static_cast<long long>(i) * ld + j4]; t.y = a21[static_cast<long long>(i) * ld + j4 + 1];
             t.z = a21[static_cast<long long>(i) * ld + j4 + 2]; t.w = a21[static_cast<long long>(i) * ld + j4 + 3]; }
    }
    *reinterpret_cast<float4*>(at + i * DS + j4) = t;
  }
  if (blockIdx.x == 1) PT(9);
  if (tid == 0) {
    const long long t0 = clock64();
    while (ld_acquire_u32(ready) != epoch) {
      __nanosleep(64);
      if (clock64() - t0 > 60000000000LL) { printf("gsmvi: potrf panel watchdog (j0=%d)\n", j0); __trap(); }
    }
  }
  __syncthreads();
  if (blockIdx.x == 1) PT(10);
  for (int q = tid; q < NB * 32; q += 256) {  // L11 (just published; lower triangle + explicit zeros)
    const int i = q >> 5, j4 = (q & 31) * 4;
    float4 t;
    if (vec_ok) t = __ldcg(reinterpret_cast<const float4*>(a11 + static_cast<long long>(i) * ld + j4));
    else { t.x = __ldcg(a11 + static_cast<long long>(i) * ld + j4); t.y = __ldcg(a11 + static_cast<long long>(i) * ld + j4 + 1);
           t.z = __ldcg(a11 + static_cast<long long>(i) * ld + j4 + 2); t.w = __ldcg(a11 + static_cast<long long>(i) * ld + j4 + 3); }
    *reinterpret_cast<float4*>(s + i * DS + j4) = t;
  }
  __syncthreads();
  if (tid < NB) dinv[tid] = 1.0f / s[tid * DS + tid];
  __syncthreads();
  if (blockIdx.x == 1) PT(11);
  {
    // TPR threads per row: thread (row, part) applies the block updates to 32/TPR of the 32 columns of block J, then
    // the part-0 threads run the in-block forward substitution.  L rows are read as broadcast float4.
    constexpr int CPT = 32 / TPR;  // columns per thread in the update phase
    const int row = tid % RPC, part = tid / RPC;
    float* myrow = at + row * DS;
#pragma unroll 1
    for (int J = 0; J < NB / 32; ++J) {
      if (J > 0) {
        float v[CPT];
        {
          const float4 t = *reinterpret_cast<const float4*>(myrow + 32 * J + CPT * part);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
#pragma unroll 1
        for (int I = 0; I < J; ++I) {  // v -= X_I L11[J][I]^T
          float4 xp[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) xp[k] = *reinterpret_cast<const float4*>(myrow + 32 * I + 4 * k);
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const float4* lr = reinterpret_cast<const float4*>(s + (32 * J + CPT * part + j) * DS + 32 * I);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 l = lr[k];
              a0 += xp[k].x * l.x;
              a1 += xp[k].y * l.y;
              a2 += xp[k].z * l.z;
              a3 += xp[k].w * l.w;
            }
            v[j] -= (a0 + a1) + (a2 + a3);
          }
        }
        *reinterpret_cast<float4*>(myrow + 32 * J + CPT * part) = make_float4(v[0], v[1], v[2], v[3]);
        __syncthreads();
      }
      if (part == 0) {  // in-block forward substitution (sequential in j)
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 t = *reinterpret_cast<const float4*>(myrow + 32 * J + k);
          v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float acc0 = v[j], acc1 = 0.0f;
          const float* lrow = s + (32 * J + j) * DS + 32 * J;
#pragma unroll
          for (int k = 0; k + 1 < j; k += 2) {
            acc0 -= v[k] * lrow[k];
            acc1 -= v[k + 1] * lrow[k + 1];
          }
          if (j & 1) acc0 -= v[j - 1] * lrow[j - 1];
          v[j] = (acc0 + acc1) * dinv[32 * J + j];
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(myrow + 32 * J + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (blockIdx.x == 1) PT(12);
  for (int q = tid; q < RPC * 32; q += 256) {
    const int i = q >> 5, j4 = (q & 31) * 4;
    if (i < rows) {
      const float4 t = *reinterpret_cast<const float4*>(at + i * DS + j4);
      if (vec_ok) *reinterpret_cast<float4*>(a21 + static_cast<long long>(i) * ld + j4) = t;
      else { a21[static_cast<long long>(i) * ld + j4] = t.x; a21[static_cast<long long>(i) * ld + j4 + 1] = t.y;
             a21[static_cast<long long>(i) * ld + j4 + 2] = t.z; a21[static_cast<long long>(i) * ld + j4 + 3] = t.w; }
    }
  }
  if (blockIdx.x == 1) PT(13);
}

size_t potrf_workspace_bytes(int n) {
  (void)n;
  return static_cast<size_t>(NB) * NB * sizeof(float);
}

int potrf_lower(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, int n, int* flag,
                float* workspace, int npass) {
  if (n <= 0 || !A || !L || !flag || !workspace) return GSMVI_EINVAL;
  static PerDeviceOnce attr_set;
  const int smem = 2 * NB * DS * sizeof(float);
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  unsigned* ready = reinterpret_cast<unsigned*>(workspace);
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemsetAsync(ready, 0, sizeof(unsigned), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  tril_copy_kernel<<<dim3((n + 255) / 256, n), 256, 0, stream>>>(A, lda, L, ldl, n);
  unsigned epoch = 0;
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = min(NB, n - j0);
    const int rest = n - j0 - nb;
    // panel: diagonal block (CTA 0) + TRSM of the rows below (one CTA per 128 rows)
    potrf_panel_kernel<<<1 + (rest + RPC - 1) / RPC, 256, smem, stream>>>(L, ldl, n, j0, nb, flag, ready, ++epoch);
    if (rest > 0) {
      float* a21 = L + static_cast<long long>(j0 + nb) * ldl + j0;
      float* a22 = L + static_cast<long long>(j0 + nb) * ldl + j0 + nb;
      // SYRK: A22 -= L21 L21^T on the lower tiles (tcgen05 3xTF32 GEMM)
      GemmOpts s;
      s.npass = npass;
      s.alpha = -1.0f;
      s.beta = 1.0f;
      s.Cin = a22;
      s.ldcin = ldl;
      s.tri = true;
      MatView vl{a21, rest, nb, ldl};
      int rc = launch_gemm_tf32(stream, rest, rest, nb, vl, vl, a22, ldl, s);
      if (rc != GSMVI_OK) return rc;
    }
  }
  if (getenv("GSMVI_POTRF_TIMING")) {
    long long h[24];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_pt, sizeof(h));
    fprintf(stderr, "[potrf panel timing] CTA0: load %lld factor %lld store+publish %lld | CTA1: prefetch %lld wait %lld loadL11 %lld solve %lld store %lld | CTA1 end - CTA0 start %lld\n",
            h[1] - h[0], h[2] - h[1], h[3] - h[2], h[9] - h[8], h[10] - h[9], h[11] - h[10], h[12] - h[11], h[13] - h[12], h[13] - h[0]);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
