// FP64 GEMM used by the BaM solve (dgemm.cu).
#pragma once
#include "tc_gemm.cuh"  // KR_* flags, status codes

namespace gsmvi {

constexpr int DGEMM_MAX_PEERS = 8;

struct DgemmOpts {
  double alpha = 1.0, beta = 0.0, diag_add = 0.0;
  const double* Cin = nullptr;
  long long ldcin = 0;
  bool tri = false, mirror = false;
  int krange = KR_FULL;
  // Row-sharded product: A holds rows row0 .. row0+M-1 of the left operand (pass it pre-offset), the result rows are global
  // rows row0 + m (for Cin, diag_add and the stores), and each finished element goes to the ncp full-size result matrices
  // Cp[0..ncp) - this GPU's own and its peers' over NVLink peer memory - instead of C.  tri / mirror / krange not allowed.
  int row0 = 0, ncp = 0;
  double* Cp[DGEMM_MAX_PEERS] = {};
};

// C[M,N] = alpha * op(A) op(B)^T + beta*Cin + diag_add*I.  a_mn/b_mn: operand stored [K, rows].
int launch_dgemm(cudaStream_t stream, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                 long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o);

}  // namespace gsmvi
