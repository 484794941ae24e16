// Blocked Cholesky with a device-side PD flag (potrf.cu).
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

size_t potrf_workspace_bytes(int n);

// L (n x n, leading dimension ldl, fully written: lower factor + zero upper triangle) <- chol(lower triangle of A).
// *flag <- 0 if every pivot was positive and finite, 1 otherwise (L is then garbage).  A and L must not alias.
// workspace: potrf_workspace_bytes(n) bytes, 16-byte aligned.  npass: 1 (TF32) or 3 (3xTF32) for the TRSM/SYRK GEMMs.
int potrf_lower(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, int n, int* flag,
                float* workspace, int npass);

}  // namespace gsmvi
