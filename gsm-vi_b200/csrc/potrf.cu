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
constexpr int LDS_ = NB + 1;  // padded shared-memory leading dimension (conflict-free column access)

// L (lower, incl. diagonal) <- lower triangle of A; strict upper triangle of L <- 0.
__global__ void tril_copy_kernel(const float* __restrict__ A, long long lda, float* __restrict__ L, long long ldl, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < n) L[static_cast<long long>(i) * ldl + j] = (j <= i) ? A[static_cast<long long>(i) * lda + j] : 0.0f;
}

// Factor the n x n (n <= 128) diagonal block at `a` (lower triangle read, leading dimension lda) in place, write
// inv(L11) (lower triangular, row-major n x n, leading dimension NB) to `linv`.  A non-positive or non-finite pivot
// sets *flag (bit 0) and the factorisation continues with NaNs (the caller discards the result).
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(float* __restrict__ a, long long lda, int n,
                                                            float* __restrict__ linv, int* __restrict__ flag) {
  extern __shared__ float sm[];
  float* s = sm;                   // [NB][LDS_]  working block -> L11
  float* x = sm + NB * LDS_;       // [NB][LDS_]  inverse
  __shared__ float rinv[NB];       // 1 / L[j][j]
  __shared__ int bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int i = idx / NB, j = idx % NB;
    s[i * LDS_ + j] = (i < n && j <= i) ? a[static_cast<long long>(i) * lda + j] : ((i == j) ? 1.0f : 0.0f);
    x[i * LDS_ + j] = 0.0f;
  }
  // ---- unblocked right-looking factorisation; column j is scaled lazily (it is never read after step j)
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const float p = s[j * LDS_ + j];
    if (!(p > 0.0f) || isinf(p)) {
      if (tid == 0) bad = 1;
    }
    const float ip = 1.0f / p;
    for (int i = j + 1 + warp; i < n; i += 8) {
      const float lij = s[i * LDS_ + j] * ip;
      for (int k = j + 1 + lane; k <= i; k += 32) s[i * LDS_ + k] -= lij * s[k * LDS_ + j];
    }
  }
  __syncthreads();
  if (tid < NB) {
    const float p = s[tid * LDS_ + tid];
    rinv[tid] = 1.0f / sqrtf(p);
  }
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int i = idx / NB, j = idx % NB;
    float v;
    if (j < i) v = s[i * LDS_ + j] * rinv[j];
    else if (j == i) v = sqrtf(s[i * LDS_ + i]);
    else v = 0.0f;
    s[i * LDS_ + j] = v;
  }
  __syncthreads();
  // ---- inverse of the unit-padded 128x128 lower-triangular L11 by 32x32 blocks
  // diagonal blocks: thread c of block-group g solves L[g] x = e_c by forward substitution
  if (tid < NB) {
    const int g = tid >> 5, c = tid & 31, o = g * 32;
    for (int i = c; i < 32; ++i) {
      float acc = (i == c) ? 1.0f : 0.0f;
      for (int k = c; k < i; ++k) acc -= s[(o + i) * LDS_ + o + k] * x[(o + k) * LDS_ + o + c];
      x[(o + i) * LDS_ + o + c] = acc / s[(o + i) * LDS_ + o + i];
    }
  }
  __syncthreads();
  // off-diagonal blocks by block distance d:  X[I][J] = -X[I][I] * sum_{K=J}^{I-1} L[I][K] X[K][J]
  float* tmp = sm + 2 * NB * LDS_;  // [32*3][33] scratch for the inner sum of up to 3 blocks
  for (int d = 1; d < 4; ++d) {
    const int nblk = 4 - d;
    // step 1: T = sum_K L[I][K] X[K][J]   (32x32 per block)
    for (int e = tid; e < nblk * 1024; e += blockDim.x) {
      const int b = e >> 10, r = (e >> 5) & 31, c = e & 31;
      const int I = b + d, J = b;
      float acc = 0.0f;
      for (int kk = J * 32; kk < I * 32; ++kk) acc += s[(I * 32 + r) * LDS_ + kk] * x[kk * LDS_ + J * 32 + c];
      tmp[(b * 32 + r) * 33 + c] = acc;
    }
    __syncthreads();
    // step 2: X[I][J] = -X[I][I] * T
    for (int e = tid; e < nblk * 1024; e += blockDim.x) {
      const int b = e >> 10, r = (e >> 5) & 31, c = e & 31;
      const int I = b + d, J = b;
      float acc = 0.0f;
      for (int kk = 0; kk <= r; ++kk) acc += x[(I * 32 + r) * LDS_ + I * 32 + kk] * tmp[(b * 32 + kk) * 33 + c];
      x[(I * 32 + r) * LDS_ + J * 32 + c] = -acc;
    }
    __syncthreads();
  }
  // ---- write back L11 (with explicit zeros above the diagonal) and inv(L11)
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int i = idx / NB, j = idx % NB;
    if (i < n && j < n) a[static_cast<long long>(i) * lda + j] = s[i * LDS_ + j];
    linv[i * NB + j] = (i < n && j < n) ? x[i * LDS_ + j] : 0.0f;
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
  const int smem = (2 * NB * LDS_ + 96 * 33) * sizeof(float);
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
