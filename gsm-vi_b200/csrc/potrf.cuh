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

// Left-looking variant on the scaled 3xFP16 engine (potrf_h3.cu): also writes L as the fp16 (hi, lo) pair `Lh` with the
// scale derived from max |A_ii|.  zero_upper = 0 skips clearing the blocks above the diagonal (they are never written,
// so a buffer that was zero once stays valid).  workspace: potrf_h3_workspace_bytes(n) bytes, 16-byte aligned.
size_t potrf_h3_workspace_bytes(int n);
// Dry run of potrf_h3's launch loop for an n x n matrix on a device with `sms` SMs: writes one row of eight ints per panel
// (j0, fused launch?, panel CTAs, GEMM CTAs, GEMM row tiles, GEMM splits, partial planes read, helper CTAs) into rows
// [max_rows][8] and returns the number of panels (negative: error).  Host only - no CUDA call is made.
int potrf_h3_plan(int n, int sms, int* rows, int max_rows);
int potrf_h3(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, const gsmvi_h3_operand& Lh, int n,
             int* flag, void* workspace, int zero_upper);

}  // namespace gsmvi
