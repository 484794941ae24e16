// ADVI gradient + Adam step on the device (advi.cu).
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

// One optimiser step of gsmvi/advi.py:69-74 given this iteration's draws Z [B, D] and scores G [B, D] at x = mu + Z L^T:
// gradient of the negative ELBO w.r.t. (mu, lower triangle of L) and the Adam update (t = 1-based step count).
// GtZ [D, ldgz], gsum [D]: scratch; mL, vL [D, ldl], m_mu, v_mu [D]: Adam moments (zero before the first step).
int advi_step(cudaStream_t st, float* L, long long ldl, float* mu, const float* G, long long ldg, const float* Z, long long ldz,
              float* GtZ, long long ldgz, float* gsum, float* mL, float* vL, float* m_mu, float* v_mu, int B, int D, float lr,
              float b1, float b2, float eps, int t, int npass);

}  // namespace gsmvi
