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

}  // namespace gsmvi
