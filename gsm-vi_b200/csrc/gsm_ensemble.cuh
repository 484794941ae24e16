// Ensemble of independent small GSM fits, one persistent CTA per fit (gsm_ensemble.cu).
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

// P [F, D, D], c [F, D] (c_f = P_f m_f), mu [F, D] and Sigma [F, D, D] (in/out), all dense row-major fp32.
// D <= 64, B <= 32.  ztape: optional [F, niter+1, B, D] standard-normal draws (parity runs), else Philox(seed).
// reverts [F]: rejected updates per fit (-1: the initial covariance was not positive definite).
// first_fit: global index of fit 0 of this call (a rank's slice of a larger ensemble draws the same Philox streams as
// the unsharded run).
int gsm_ensemble_fit(cudaStream_t st, const float* P, const float* c, float* mu, float* Sigma, int F, int D, int B,
                     int niter, unsigned long long seed, const float* ztape, int* reverts, int first_fit = 0);

}  // namespace gsmvi
