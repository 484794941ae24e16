// KLMonitor device reductions (monitor.cu).
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

// out[0] <- sum_b log N(x_b | mu, L L^T) for x_b = mu + L z_b (only Z and diag(L) are read).
int gauss_logq_from_z(cudaStream_t st, const float* Z, long long ldz, int N, int D, const float* L, long long ldl,
                      double* out);
// out[0] <- sum_b log N(x_b | mu, L L^T) for arbitrary rows x_b (forward substitution per sample).
int gauss_logq_from_x(cudaStream_t st, const float* X, long long ldx, int N, int D, const float* mu, const float* L,
                      long long ldl, double* out);

}  // namespace gsmvi
