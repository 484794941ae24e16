// C-ABI of libgsmvi_b200.so (declared in include/gsmvi_b200.h). Plain pointers and sizes only; no torch types.
#include "../../include/gsmvi_b200.h"

#include "gsm_kernels.cuh"
#include "potrf.cuh"
#include "tc_gemm.cuh"

using namespace gsmvi;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int gsmvi_abi_version(void) { return GSMVI_ABI_VERSION; }

long long gsmvi_workspace_bytes(int kind, int B, int D) {
  switch (kind) {
    case GSMVI_WS_POTRF: return static_cast<long long>(potrf_workspace_bytes(D));
    case GSMVI_WS_GSM_UPDATE: return static_cast<long long>(gsm_update_workspace_bytes(B, D));
    default: return -1;
  }
}

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
  return launch_gemm_tf32(S(stream), M, N, K, a, b, C, ldc, o);
}

int gsmvi_potrf_check(const float* Sigma, long long lds, float* L, long long ldl, int D, int* bad_flag,
                      void* workspace, int npass, void* stream) {
  return potrf_lower(S(stream), Sigma, lds, L, ldl, D, bad_flag, static_cast<float*>(workspace), npass);
}

int gsmvi_philox_normal(float* Z, long long ldz, int B, int D, unsigned long long seed, unsigned long long offset,
                        void* stream) {
  return philox_normal(S(stream), Z, ldz, B, D, seed, offset);
}

int gsmvi_sample(const float* mu, const float* L, long long ldl, const float* Z, long long ldz, float* X,
                 long long ldx, int B, int D, int npass, void* stream) {
  if (!mu || !L || !Z || !X || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return sample_mvn(S(stream), mu, L, ldl, Z, ldz, X, ldx, B, D, npass);
}

int gsmvi_gauss_score(const float* X, long long ldx, const float* P, long long ldp, const float* c, float* G,
                      long long ldg, int B, int D, int npass, void* stream) {
  if (!X || !P || !c || !G || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return gauss_score(S(stream), X, ldx, P, ldp, c, G, ldg, B, D, npass);
}

int gsmvi_gsm_update(const float* X, long long ldx, const float* G, long long ldg, const float* mu,
                     const float* Sigma, long long lds, float* mu_out, float* Sigma_out, long long ldso, int B, int D,
                     int B_total, int mode, void* workspace, int npass, void* stream) {
  return gsm_update(S(stream), X, ldx, G, ldg, mu, Sigma, lds, mu_out, Sigma_out, ldso, B, D, B_total, mode,
                    static_cast<float*>(workspace), npass);
}

int gsmvi_gsm_apply_stats(const float* Sigma, long long lds, const float* dSigma, long long ldd, const float* mu,
                          const float* dmu, float* Sigma_out, long long ldso, float* mu_out, int D, void* stream) {
  if (!Sigma || !dSigma || !mu || !dmu || !Sigma_out || !mu_out || D <= 0) return GSMVI_EINVAL;
  return gsm_apply_stats(S(stream), Sigma, lds, dSigma, ldd, mu, dmu, Sigma_out, ldso, mu_out, D);
}

}  // extern "C"
