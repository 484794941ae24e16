// fp64 single-CTA GSM iteration for D <= 64 (gsm_small64.cu): BASELINE configs[0], the reference's numpy example.
#pragma once
#include <cuda_runtime.h>

#include "../../include/gsmvi_b200.h"

namespace gsmvi {

long long gsm_small64_workspace_bytes(int B, int D);
int gsm_small64(cudaStream_t st, int mode, double* mu, double* Sigma, double* L, const float* ztape, unsigned long long seed,
                unsigned long long iter0, double* X, const double* G, const double* P, const double* c, int B, int D,
                int iters, int* status, void* workspace);

}  // namespace gsmvi
