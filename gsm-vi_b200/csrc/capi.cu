// C-ABI of libgsmvi_b200.so (declared in include/gsmvi_b200.h). Plain pointers and sizes only.
#include "../../include/gsmvi_b200.h"

#include "tc_gemm.cuh"

using namespace gsmvi;

extern "C" {

int gsmvi_abi_version(void) { return GSMVI_ABI_VERSION; }

int gsmvi_gemm_tf32(const float* A, long long a_rows, long long a_cols, long long lda, int a_mn, const float* B,
                    long long b_rows, long long b_cols, long long ldb, int b_mn, float* C, long long ldc, int M, int N,
                    int K, float alpha, float beta, const float* Cin, long long ldcin, const float* bias_n, int npass,
                    int tri, int mirror, int krange, int neg_from, void* stream) {
  GemmOpts o;
  o.npass = npass;
  o.a_mn = a_mn != 0;
  o.b_mn = b_mn != 0;
  o.alpha = alpha;
  o.beta = beta;
  o.Cin = Cin;
  o.ldcin = ldcin;
  o.bias_n = bias_n;
  o.tri = tri != 0;
  o.mirror = mirror != 0;
  o.krange = krange;
  o.neg_from = neg_from;
  MatView a{A, a_rows, a_cols, lda}, b{B, b_rows, b_cols, ldb};
  return launch_gemm_tf32(static_cast<cudaStream_t>(stream), M, N, K, a, b, C, ldc, o);
}

}  // extern "C"
