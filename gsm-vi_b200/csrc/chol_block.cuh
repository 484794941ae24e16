// Warp-level 32x32 Cholesky step shared by the blocked Cholesky panel kernel (potrf.cu) and the ensemble kernel
// (gsm_ensemble.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gsmvi {

// One column step of the warp-level 32x32 Cholesky (row `lane` of the block lives in row[0..31]); the recursion on the
// template parameter forces full unrolling so that row[] is only ever indexed statically (stays in registers).
template <int J>
__device__ __forceinline__ void chol32_step(float (&row)[32], int lane, float* dinv_out, int& isbad) {
  if constexpr (J < 32) {
    const float d = __shfl_sync(0xffffffffu, row[J], J);
    if (!(d > 0.0f) || isinf(d)) isbad = 1;
    float r = rsqrtf(d);
    r = r * (1.5f - 0.5f * d * r * r);          // one Newton step: full fp32 accuracy without the slow sqrt + divide
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

// Blocked variant of the same warp-level 32x32 Cholesky: four 8-column steps.  Each step gathers the 8x8 diagonal
// block into EVERY lane (36 shuffles, issued back to back), factors it redundantly in registers - so the eight
// dependent pivots (rsqrt + one Newton step each) have no cross-lane traffic between them - then every row solves its
// eight entries against that block locally and the remaining columns take a rank-8 update (8 shuffles per column).
// The per-column dependent chain drops from shuffle + rsqrt + shuffle (~110 cycles) to rsqrt + a few FMAs (~60).
// On return lane i holds row i of L in row[0..i]; row[k] for k > i is undefined (callers mask it to zero).
__device__ __forceinline__ void chol32_b8(float (&row)[32], int lane, float* dinv_out, int& isbad) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c0 = 8 * q;
    float d[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) d[r][c] = __shfl_sync(0xffffffffu, row[c0 + c], c0 + r);
    float rinv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dj = d[j][j];
      if (!(dj > 0.0f) || isinf(dj)) isbad = 1;
      float r = rsqrtf(dj);
      r = r * (1.5f - 0.5f * dj * r * r);
      rinv[j] = r;
#pragma unroll
      for (int i = j + 1; i < 8; ++i) d[i][j] *= r;
#pragma unroll
      for (int i = j + 1; i < 8; ++i)
#pragma unroll
        for (int k = j + 1; k <= i; ++k) d[i][k] -= d[i][j] * d[k][j];
    }
    // x L8^T = a for this lane's row (rows inside the block reproduce their own row of L8 for j <= r)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = row[c0 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) x -= row[c0 + k] * d[j][k];
      // the diagonal entry itself: l_jj = d_jj * rinv_j, computed from the un-scaled pivot for full accuracy
      row[c0 + j] = (lane == c0 + j) ? d[j][j] * rinv[j] : x * rinv[j];
      if (lane == c0 + j) dinv_out[c0 + j] = rinv[j];
    }
    // rank-8 update of the remaining columns: row_i[k] -= sum_c l_ic l_kc
#pragma unroll
    for (int k = c0 + 8; k < 32; ++k) {
      float acc0 = row[k], acc1 = 0.0f;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        acc0 -= row[c0 + c] * __shfl_sync(0xffffffffu, row[c0 + c], k);
        acc1 -= row[c0 + c + 1] * __shfl_sync(0xffffffffu, row[c0 + c + 1], k);
      }
      row[k] = acc0 + acc1;
    }
  }
}

// Rolled form of chol32_b8: one loop over the four 8-column steps; the active block always sits in row[0..7] and the
// finished columns are written out and rotated away, so every register index stays static.  Finished columns go to
// out_row[0..31] (this lane's row of the block in shared memory, zeros above the diagonal) and, transposed, to
// dT[col * ldt + lane]; dinv_out[col] <- 1 / l_colcol.  Pivots are NOT checked on the dependent chain: a non-positive or
// non-finite pivot turns its column into NaN/Inf, which the caller sees in the returned diagonal entry (isbad).
__device__ __forceinline__ void chol32_rolled(float (&row)[32], int lane, float* __restrict__ out_row,
                                              float* __restrict__ dT, int ldt, float* __restrict__ dinv_out, int& isbad) {
  float mydiag = 1.0f;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int c0 = 8 * q;
    float d[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) d[r][c] = __shfl_sync(0xffffffffu, row[c], c0 + r);
    float rinv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dj = d[j][j];
      float r;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dj));  // one MUFU; the library rsqrtf adds range fix-up code
      r = r * fmaf(-0.5f * dj, r * r, 1.5f);                     // one Newton step: full fp32 accuracy
      rinv[j] = r;
#pragma unroll
      for (int i = j + 1; i < 8; ++i) d[i][j] *= r;
#pragma unroll
      for (int i = j + 1; i < 8; ++i)
#pragma unroll
        for (int k = j + 1; k <= i; ++k) d[i][k] -= d[i][j] * d[k][j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = row[j];
#pragma unroll
      for (int k = 0; k < j; ++k) x -= row[k] * d[j][k];
      row[j] = (lane == c0 + j) ? d[j][j] * rinv[j] : x * rinv[j];
      if (lane == c0 + j) {
        dinv_out[c0 + j] = rinv[j];
        mydiag = row[j];
      }
    }
    // rank-8 update of the live trailing columns, in three 8-column chunks (q = 0: 24 live columns, q = 2: 8).  A warp
    // issues in order, so the 32 shuffles of four columns go out before their FMAs: one exposed shuffle latency per
    // four columns instead of one per column.
#define GSMVI_CHOL_QUAD(K0)                                                                \
    {                                                                                      \
      float t[4][8];                                                                       \
      _Pragma("unroll") for (int kk = 0; kk < 4; ++kk)                                      \
        _Pragma("unroll") for (int c = 0; c < 8; ++c) t[kk][c] = __shfl_sync(0xffffffffu, row[c], c0 + (K0) + kk); \
      _Pragma("unroll") for (int kk = 0; kk < 4; ++kk) {                                    \
        float acc0 = row[(K0) + kk], acc1 = 0.0f;                                          \
        _Pragma("unroll") for (int c = 0; c < 8; c += 2) {                                  \
          acc0 -= row[c] * t[kk][c];                                                       \
          acc1 -= row[c + 1] * t[kk][c + 1];                                               \
        }                                                                                  \
        row[(K0) + kk] = acc0 + acc1;                                                      \
      }                                                                                    \
    }
#define GSMVI_CHOL_CHUNK(K0) GSMVI_CHOL_QUAD(K0) GSMVI_CHOL_QUAD((K0) + 4)
    if (q < 3) { GSMVI_CHOL_CHUNK(8) }
    if (q < 2) { GSMVI_CHOL_CHUNK(16) }
    if (q < 1) { GSMVI_CHOL_CHUNK(24) }
#undef GSMVI_CHOL_QUAD
#undef GSMVI_CHOL_CHUNK
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = (c0 + j <= lane) ? row[j] : 0.0f;
      dT[(c0 + j) * ldt + lane] = o[j];
    }
    *reinterpret_cast<float4*>(out_row + c0) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(out_row + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
    for (int k = 0; k < 24; ++k) row[k] = row[k + 8];
  }
  if (__any_sync(0xffffffffu, !(mydiag > 0.0f) || !(mydiag < 3.0e38f))) isbad = 1;
}

}  // namespace gsmvi
