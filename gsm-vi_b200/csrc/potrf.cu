// Blocked right-looking Cholesky  Sigma = L L^T  (fp32) with a device-side "is positive definite" flag.
//
// Replaces two host round trips of the reference per iteration: the factorisation inside the sampler
// np.random.multivariate_normal (gsmvi/gsm.py:119, bam.py:193 - an SVD there) and the PD check
// np.linalg.cholesky in _check_goodness (gsm.py:136-150, bam.py:219-233).  One factorisation serves both: the flag
// accepts/rejects the update, and on accept L is the next iteration's sampling factor.
//
// Per 128-column panel:  (1) one CTA factors the 128x128 diagonal block in shared memory and also forms its
// inverse; (2) TRSM as a GEMM  L21 = A21 * inv(L11)^T;  (3) SYRK trailing update  A22 -= L21 L21^T, lower tiles
// only - both on the tcgen05 3xTF32 GEMM.
#include "potrf.cuh"

#include <math.h>

namespace gsmvi {

constexpr int NB = 128;

// L (lower, incl. diagonal) <- lower triangle of A; strict upper triangle of L <- 0.
__global__ void tril_copy_kernel(const float* __restrict__ A, long long lda, float* __restrict__ L, long long ldl, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < n) L[static_cast<long long>(i) * ldl + j] = (j <= i) ? A[static_cast<long long>(i) * lda + j] : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// Diagonal-block kernel: factor the n x n (n <= 128) block at `a` (lower triangle read, leading dimension lda) in
// place and write inv(L11) (lower triangular, row-major, leading dimension NB, zero/identity padded) to `linv`.
// A non-positive or non-finite pivot sets *flag (bit 0); the factorisation then continues with NaNs and the caller
// discards the result.
//
// One CTA, 256 threads, everything in shared memory / registers.  The block is processed in four 32-column panels:
//   (1) warp 0 factors the 32x32 diagonal block, one row per lane in registers, pivots broadcast with shuffles;
//   (2) a thread per row below solves its 32 panel entries against that block by forward substitution (registers);
//   (3) all threads apply the rank-32 update to the trailing lower triangle in 4x4 register tiles.
// The inverse is then formed by 32x32 blocks: diagonal blocks by per-column substitution in registers, off-diagonal
// blocks X[I][J] = -X[I][I] sum_K L[I][K] X[K][J] in 4x4 register tiles, by block distance.
constexpr int DS = NB + 4;  // shared-memory leading dimension: rows stay 16-byte aligned, quarter-warps hit distinct banks

__device__ __forceinline__ void tile4x4_mac(const float* __restrict__ Arows, int lda_, const float* __restrict__ Brows,
                                            int ldb_, int kbeg, int kend, float (&acc)[4][4]) {
  // acc[r][c] += sum_k Arows[r][k] * Brows[c][k]   (both row-major over k; kbeg, kend multiples of 4)
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

// One column step of the warp-level 32x32 Cholesky (row `lane` of the block lives in row[0..31]); the recursion on the
// template parameter forces full unrolling so that row[] is only ever indexed statically (stays in registers).
template <int J>
__device__ __forceinline__ void chol32_step(float (&row)[32], int lane, float* dinv_out, int& isbad) {
  if constexpr (J < 32) {
    const float d = __shfl_sync(0xffffffffu, row[J], J);
    if (!(d > 0.0f) || isinf(d)) isbad = 1;
    const float r = 1.0f / sqrtf(d);
    row[J] = (lane == J) ? d * r : row[J] * r;  // l_jj = sqrt(d), l_ij = a_ij / l_jj
    if (lane == J) dinv_out[J] = r;
#pragma unroll
    for (int k = J + 1; k < 32; ++k) {
      const float lk = __shfl_sync(0xffffffffu, row[J], k);
      if (lane >= k) row[k] -= row[J] * lk;
    }
    chol32_step<J + 1>(row, lane, dinv_out, isbad);
  }
}

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(float* __restrict__ a, long long lda, int n,
                                                            float* __restrict__ linv, int* __restrict__ flag) {
  extern __shared__ __align__(16) float sm[];
  float* s = sm;             // [NB][DS]  working block -> L11
  float* x = sm + NB * DS;   // [NB][DS]  inverse, stored TRANSPOSED: x[c][r] = inv(L)[r][c]
  __shared__ float dinv[NB];
  __shared__ int bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, j = idx & (NB - 1);
    s[i * DS + j] = (i < n && j <= i) ? a[static_cast<long long>(i) * lda + j] : ((i == j) ? 1.0f : 0.0f);
    x[i * DS + j] = 0.0f;
  }
  __syncthreads();

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
      if (isbad && lane == 0) bad = 1;
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
        float acc = v[j];
        const float* lrow = s + (c0 + j) * DS + c0;  // broadcast reads
#pragma unroll
        for (int k = 0; k < j; ++k) acc -= v[k] * lrow[k];
        v[j] = acc * dinv[c0 + j];
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(s + tid * DS + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
    __syncthreads();
    // ---- (3) trailing update: S[i][k] -= sum_c P[i][c] P[k][c], lower 4x4 tiles of the (NB-c0-32)^2 block
    const int m0 = c0 + 32, mt = (NB - m0) / 4;  // tiles per side
    for (int t = tid; t < mt * (mt + 1) / 2; t += 256) {
      int ti = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      const int tj = t - ti * (ti + 1) / 2;
      const int i0 = m0 + 4 * ti, k0 = m0 + 4 * tj;
      float acc[4][4] = {};
      tile4x4_mac(s + i0 * DS, DS, s + k0 * DS, DS, c0, c0 + 32, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (k0 + c <= i0 + r) s[(i0 + r) * DS + k0 + c] -= acc[r][c];
    }
    __syncthreads();
  }

  // ---- inverse, diagonal 32x32 blocks: thread (g, c) solves L[g] y = e_c in registers; x holds inv(L)^T
  if (tid < NB) {
    const int g = tid >> 5, c = tid & 31, o = g * 32;
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float acc = (i == c) ? 1.0f : 0.0f;
      const float* lrow = s + (o + i) * DS + o;
#pragma unroll
      for (int k = 0; k < i; ++k) acc -= lrow[k] * y[k];  // y[k] == 0 for k < c
      y[i] = acc * dinv[o + i];
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(x + (o + c) * DS + o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
  }
  __syncthreads();
  // ---- off-diagonal blocks by block distance d (I = J + d).  With XT = inv(L)^T stored row-major in x:
  //   T[r][c]  = sum_k L[I*32+r][k] * XT[J*32+c][k],   k in [J*32, I*32)
  //   X[I][J][r][c] = - sum_kk XT... : inv(L)[I*32+r][I*32+kk] = XT[I*32+kk][I*32+r]  (needs a column walk), so the
  //   second product is taken from a transposed copy of the diagonal block inverse kept in `tmp`.
  float* tmp = sm + 2 * NB * DS;   // [96][36]  T, row-major
  float* dgi = tmp + 96 * 36;      // [NB][36]  inv(L[I][I]) row-major (not transposed), all four diagonal blocks
  for (int idx = tid; idx < NB * 32; idx += 256) {
    const int r = idx >> 5, c = idx & 31, o = (r >> 5) * 32;
    dgi[r * 36 + c] = x[(o + c) * DS + r];  // inv(L)[r][o+c]
  }
  __syncthreads();
  for (int d = 1; d < NB / 32; ++d) {
    const int nblk = NB / 32 - d;
    for (int t = tid; t < nblk * 64; t += 256) {  // 64 4x4 tiles per 32x32 block
      const int b = t >> 6, tr = (t >> 3) & 7, tc = t & 7;
      const int I = b + d, J = b;
      float acc[4][4] = {};
      tile4x4_mac(s + (I * 32 + 4 * tr) * DS, DS, x + (J * 32 + 4 * tc) * DS, DS, J * 32, I * 32, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) tmp[(b * 32 + 4 * tc + c) * 36 + 4 * tr + r] = acc[r][c];  // store T^T: tmp[c][r]
    }
    __syncthreads();
    for (int t = tid; t < nblk * 64; t += 256) {
      const int b = t >> 6, tr = (t >> 3) & 7, tc = t & 7;
      const int I = b + d, J = b;
      // X[I][J][r][c] = - sum_kk inv(L[I][I])[r][kk] * T[kk][c] = - sum_kk dgi[I*32+r][kk] * tmp[c][kk]
      float acc[4][4] = {};
      tile4x4_mac(dgi + (I * 32 + 4 * tr) * 36, 36, tmp + (b * 32 + 4 * tc) * 36, 36, 0, 32, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) x[(J * 32 + 4 * tc + c) * DS + I * 32 + 4 * tr + r] = -acc[r][c];
    }
    __syncthreads();
  }
  // ---- write back L11 (explicit zeros above the diagonal) and inv(L11)
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, j = idx & (NB - 1);
    if (i < n && j < n) a[static_cast<long long>(i) * lda + j] = (j <= i) ? s[i * DS + j] : 0.0f;
    linv[i * NB + j] = (i < n && j < n && j <= i) ? x[j * DS + i] : 0.0f;
  }
  if (tid == 0 && bad) atomicOr(flag, 1);
}

size_t potrf_workspace_bytes(int n) {
  (void)n;
  return static_cast<size_t>(NB) * NB * sizeof(float);
}

int potrf_lower(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, int n, int* flag,
                float* workspace, int npass) {
  if (n <= 0 || !A || !L || !flag || !workspace) return GSMVI_EINVAL;
  static bool attr_set = false;
  const int smem = (2 * NB * DS + 96 * 36 + NB * 36) * sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  tril_copy_kernel<<<dim3((n + 255) / 256, n), 256, 0, stream>>>(A, lda, L, ldl, n);
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = min(NB, n - j0);
    const int rest = n - j0 - nb;
    float* l11 = L + static_cast<long long>(j0) * ldl + j0;
    potrf_diag_kernel<<<1, 256, smem, stream>>>(l11, ldl, nb, workspace, flag);
    if (rest > 0) {
      float* a21 = L + static_cast<long long>(j0 + nb) * ldl + j0;
      float* a22 = L + static_cast<long long>(j0 + nb) * ldl + j0 + nb;
      // TRSM: L21 = A21 * inv(L11)^T, in place (each CTA reads only the rows it then overwrites)
      GemmOpts t;
      t.npass = npass;
      MatView va{a21, rest, nb, ldl}, vb{workspace, nb, nb, NB};
      int rc = launch_gemm_tf32(stream, rest, nb, nb, va, vb, a21, ldl, t);
      if (rc != GSMVI_OK) return rc;
      // SYRK: A22 -= L21 L21^T on the lower tiles
      GemmOpts s;
      s.npass = npass;
      s.alpha = -1.0f;
      s.beta = 1.0f;
      s.Cin = a22;
      s.ldcin = ldl;
      s.tri = true;
      MatView vl{a21, rest, nb, ldl};
      rc = launch_gemm_tf32(stream, rest, rest, nb, vl, vl, a22, ldl, s);
      if (rc != GSMVI_OK) return rc;
    }
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
