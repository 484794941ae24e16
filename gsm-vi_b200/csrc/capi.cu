// C-ABI of libgsmvi_b200.so (declared in include/gsmvi_b200.h). Plain pointers and sizes only; no torch types.
#include "../../include/gsmvi_b200.h"

#include "advi.cuh"
#include "bam_solve.cuh"
#include "comm.cuh"
#include "dgemm.cuh"
#include "gsm_ensemble.cuh"
#include "gsm_kernels.cuh"
#include "gsm_small64.cuh"
#include "h3_gemm.cuh"
#include "monitor.cuh"
#include "oz_gemm.cuh"
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
    case GSMVI_WS_POTRF_H3: return static_cast<long long>(potrf_h3_workspace_bytes(D));
    case GSMVI_WS_GSM_UPDATE_H3: return static_cast<long long>(gsm_update_h3_workspace_bytes(B, D));
    case GSMVI_WS_BAM_STATS: return static_cast<long long>(bam_stats_workspace_bytes(B, D));
    case GSMVI_WS_BAM_SOLVE: return static_cast<long long>(bam_solve_workspace_bytes(B, D, 0));
    case GSMVI_WS_BAM_SOLVE_LOWRANK: return static_cast<long long>(bam_solve_workspace_bytes(B, D, 1));
    default: return -1;
  }
}

int gsmvi_gemm_tf32(const float* A, long long a_rows, long long a_cols, long long lda, int a_mn, const float* B,
                    long long b_rows, long long b_cols, long long ldb, int b_mn, float* C, long long ldc, int M, int N,
                    int K, float alpha, float beta, const float* Cin, long long ldcin, const float* bias_n, int npass,
                    int tri, int mirror, int krange, int neg_from, const float* A_lo, const float* B_lo, void* stream) {
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
  MatView a{A, a_rows, a_cols, lda, A_lo}, b{B, b_rows, b_cols, ldb, B_lo};
  return launch_gemm_tf32(S(stream), M, N, K, a, b, C, ldc, o);
}

int gsmvi_gemm_h3(const void* A_hi, const void* A_lo, const float* scale_a, long long a_rows, long long a_cols, long long lda,
                  int a_mn, const void* B_hi, const void* B_lo, const float* scale_b, long long b_rows, long long b_cols,
                  long long ldb, int b_mn, float* C, long long ldc, int M, int N, int K, float alpha, float beta,
                  const float* Cin, long long ldcin, const float* bias_n, int tri, int mirror, int krange,
                  unsigned* absmax_out, int splits, long long split_stride, void* stream) {
  H3Opts o;
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
  o.absmax_out = absmax_out;
  o.splits = splits;
  o.split_stride = split_stride;
  HView a{static_cast<const __half*>(A_hi), static_cast<const __half*>(A_lo), a_rows, a_cols, lda, scale_a};
  HView b{static_cast<const __half*>(B_hi), static_cast<const __half*>(B_lo), b_rows, b_cols, ldb, scale_b};
  return launch_gemm_h3(S(stream), M, N, K, a, b, C, ldc, o);
}

int gsmvi_h3_pair_kernel(int enable) { return h3_pair_kernel(enable); }

int gsmvi_h3_absmax(const float* A, long long lda, int rows, int cols, unsigned* absmax, void* stream) {
  return h3_absmax(S(stream), A, lda, rows, cols, absmax);
}

int gsmvi_h3_split(const float* A, long long lda, int rows, int cols, const unsigned* absmax, int sqrt_mode,
                   float* scale_out, void* A_hi, void* A_lo, long long ldo, void* stream) {
  return h3_split(S(stream), A, lda, rows, cols, absmax, sqrt_mode, scale_out, static_cast<__half*>(A_hi),
                  static_cast<__half*>(A_lo), ldo);
}

int gsmvi_philox_normal_h3(const gsmvi_h3_operand* Z, int B, int D, unsigned long long seed, unsigned long long offset,
                           const unsigned long long* offset_dev, void* stream) {
  if (!Z) return GSMVI_EINVAL;
  return philox_normal_h3(S(stream), *Z, B, D, seed, offset, offset_dev);
}

int gsmvi_sample_h3(const float* mu, const gsmvi_h3_operand* L, const gsmvi_h3_operand* Z, float* X, long long ldx,
                    unsigned* absmax_x, const gsmvi_h3_operand* X_split, int B, int D, void* stream) {
  if (!mu || !L || !Z || !X || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return sample_mvn_h3(S(stream), mu, *L, *Z, X, ldx, absmax_x, B, D, X_split);
}

int gsmvi_gauss_score_h3(const gsmvi_h3_operand* X, const gsmvi_h3_operand* P, const float* c, float* G, long long ldg,
                         unsigned* absmax_g, const gsmvi_h3_operand* G_split, int B, int D, void* stream) {
  if (!X || !P || !c || !G || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return gauss_score_h3(S(stream), *X, *P, c, G, ldg, absmax_g, B, D, G_split);
}

int gsmvi_h3_bound_scales(const float* mu, int D, const unsigned* sigma_absmax, const unsigned* zmax_bits, float zmax_const,
                          float pnorm, float cmax, float* scale_x, float* scale_g, void* stream) {
  return h3_bound_scales(S(stream), mu, D, sigma_absmax, zmax_bits, zmax_const, pnorm, cmax, scale_x, scale_g);
}

int gsmvi_gsm_update_h3(const float* X, long long ldx, const float* G, long long ldg, const gsmvi_h3_operand* G_split,
                        const float* mu, const float* Sigma, long long lds, const gsmvi_h3_operand* Sigma_split,
                        float* mu_out, float* Sigma_out, long long ldso, unsigned* absmax_sout, int B, int D, int B_total,
                        int mode, void* workspace, void* stream) {
  if (!G_split || !Sigma_split) return GSMVI_EINVAL;
  return gsm_update_h3(S(stream), X, ldx, G, ldg, *G_split, mu, Sigma, lds, *Sigma_split, mu_out, Sigma_out, ldso,
                       absmax_sout, B, D, B_total, mode, workspace);
}

int gsmvi_potrf_check(const float* Sigma, long long lds, float* L, long long ldl, int D, int* bad_flag,
                      void* workspace, int npass, void* stream) {
  return potrf_lower(S(stream), Sigma, lds, L, ldl, D, bad_flag, static_cast<float*>(workspace), npass);
}

long long gsmvi_comm_layout_bytes(int D, int world, gsmvi_comm_layout* lay) { return comm_layout(D, world, lay); }
int gsmvi_comm_alloc(long long bytes, void** dev_ptr_out, unsigned char* handle64_out) {
  return comm_alloc(bytes, dev_ptr_out, handle64_out);
}
int gsmvi_comm_open(const unsigned char* handle64, void** dev_ptr_out) { return comm_open(handle64, dev_ptr_out); }
int gsmvi_comm_close(void* peer_ptr) { return comm_close(peer_ptr); }
int gsmvi_comm_free(void* dev_ptr) { return comm_free(dev_ptr); }

int gsmvi_gsm_update_h3_fused(const float* X, long long ldx, const float* G, long long ldg, const gsmvi_h3_operand* G_split,
                              const float* mu, const gsmvi_h3_operand* Sigma_split, float* mu_out, void* const* peer_base,
                              void* own_base, const gsmvi_comm_layout* lay, int rank, int world, int cur, unsigned step, int B, int D,
                              int B_total, void* workspace, void* stream) {
  if (!G_split || !Sigma_split || !lay) return GSMVI_EINVAL;
  return gsm_update_h3_fused(S(stream), X, ldx, G, ldg, *G_split, mu, *Sigma_split, mu_out,
                             reinterpret_cast<float* const*>(peer_base), static_cast<float*>(own_base), *lay, rank, world, cur,
                             step, B, D, B_total, workspace);
}

int gsmvi_potrf_h3(const float* Sigma, long long lds, float* L, long long ldl, const gsmvi_h3_operand* L_split, int D,
                   int* bad_flag, void* workspace, int zero_upper, void* stream) {
  if (!L_split) return GSMVI_EINVAL;
  return potrf_h3(S(stream), Sigma, lds, L, ldl, *L_split, D, bad_flag, workspace, zero_upper);
}

int gsmvi_potrf_h3_plan(int D, int sms, int* rows, int max_rows) { return potrf_h3_plan(D, sms, rows, max_rows); }

int gsmvi_philox_normal(float* Z, long long ldz, int B, int D, unsigned long long seed, unsigned long long offset,
                        void* stream) {
  return philox_normal(S(stream), Z, ldz, B, D, seed, offset);
}

int gsmvi_tf32_split(const float* A, long long lda, float* A_hi, float* A_lo, long long ldo, int rows, int cols,
                     void* stream) {
  return tf32_split(S(stream), A, lda, A_hi, A_lo, ldo, rows, cols);
}

int gsmvi_sample(const float* mu, const float* L, const float* L_lo, long long ldl, const float* Z, long long ldz,
                 float* X, long long ldx, int B, int D, int npass, void* stream) {
  if (!mu || !L || !Z || !X || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return sample_mvn(S(stream), mu, L, L_lo, ldl, Z, ldz, X, ldx, B, D, npass);
}

int gsmvi_gauss_score(const float* X, long long ldx, const float* P, const float* P_lo, long long ldp, const float* c,
                      float* G, long long ldg, int B, int D, int npass, void* stream) {
  if (!X || !P || !c || !G || B <= 0 || D <= 0) return GSMVI_EINVAL;
  return gauss_score(S(stream), X, ldx, P, P_lo, ldp, c, G, ldg, B, D, npass);
}

int gsmvi_gsm_update(const float* X, long long ldx, const float* G, long long ldg, const float* mu,
                     const float* Sigma, const float* Sigma_hi, const float* Sigma_lo, long long lds, float* mu_out,
                     float* Sigma_out, long long ldso, int B, int D, int B_total, int mode, void* workspace, int npass,
                     void* stream) {
  return gsm_update(S(stream), X, ldx, G, ldg, mu, Sigma, Sigma_hi, Sigma_lo, lds, mu_out, Sigma_out, ldso, B, D, B_total, mode,
                    static_cast<float*>(workspace), npass);
}

int gsmvi_gsm_apply_stats(const float* Sigma, long long lds, const float* dSigma, long long ldd, const float* mu,
                          const float* dmu, float* Sigma_out, long long ldso, float* mu_out, int D, void* stream) {
  if (!Sigma || !dSigma || !mu || !dmu || !Sigma_out || !mu_out || D <= 0) return GSMVI_EINVAL;
  return gsm_apply_stats(S(stream), Sigma, lds, dSigma, ldd, mu, dmu, Sigma_out, ldso, mu_out, D);
}

int gsmvi_dgemm(const double* A, long long lda, int a_mn, const double* B, long long ldb, int b_mn, double* C,
                long long ldc, int M, int N, int K, double alpha, double beta, const double* Cin, long long ldcin,
                double diag_add, int tri, int mirror, int krange, void* stream) {
  DgemmOpts o;
  o.alpha = alpha;
  o.beta = beta;
  o.diag_add = diag_add;
  o.Cin = Cin;
  o.ldcin = ldcin;
  o.tri = tri != 0;
  o.mirror = mirror != 0;
  o.krange = krange;
  return launch_dgemm(S(stream), M, N, K, A, lda, a_mn != 0, B, ldb, b_mn != 0, C, ldc, o);
}

long long gsmvi_dgemm_oz_workspace_bytes(int M, int N, int K, int slices) {
  return static_cast<long long>(oz_workspace_bytes(M, N, K, slices));
}

int gsmvi_dgemm_oz(const double* A, long long lda, int a_mn, const double* B, long long ldb, int b_mn, double* C,
                   long long ldc, int M, int N, int K, double alpha, double beta, const double* Cin, long long ldcin,
                   double diag_add, int tri, int mirror, void* workspace, int slices, void* stream) {
  DgemmOpts o;
  o.alpha = alpha;
  o.beta = beta;
  o.diag_add = diag_add;
  o.Cin = Cin;
  o.ldcin = ldcin;
  o.tri = tri != 0;
  o.mirror = mirror != 0;
  return launch_dgemm_oz(S(stream), M, N, K, A, lda, a_mn != 0, B, ldb, b_mn != 0, C, ldc, o, workspace, slices);
}

int gsmvi_bam_stats(const float* X, long long ldx, const float* G, long long ldg, int B, int D, int B_total,
                    void* stats_workspace, int npass, int stage, void* stream) {
  if (!X || !G || !stats_workspace || B <= 0 || D <= 0 || B_total < B || (stage != 0 && stage != 1)) return GSMVI_EINVAL;
  (void)npass;
  return bam_stats(S(stream), X, ldx, G, ldg, B, D, B_total, static_cast<double*>(stats_workspace), stage);
}

int gsmvi_bam_solve(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0, long long lds0,
                    double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso, void* solve_workspace,
                    int max_ns_iters, int* ns_iters_host, int* bad_flag, int world, int phase, void* stream) {
  if (!stats_workspace || !mu0 || !Sigma0 || !mu_out || !Sigma_out || !solve_workspace || !bad_flag || B <= 0 || D <= 0 ||
      world < 1 || phase < 0 || phase > 2)
    return GSMVI_EINVAL;
  return bam_solve_full(S(stream), static_cast<const double*>(stats_workspace), B, D, B_total, mu0, Sigma0, lds0, reg, jitter, mu_out,
                        Sigma_out, ldso, static_cast<double*>(solve_workspace), max_ns_iters, ns_iters_host, bad_flag, world,
                        phase);
}

int gsmvi_potrf64(double* A, long long lda, int n, int* bad_flag, void* stream) {
  return potrf64(S(stream), A, lda, n, bad_flag);
}

int gsmvi_bam_solve_sharded(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0,
                            long long lds0, double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso,
                            void* solve_workspace, int max_ns_iters, int* ns_iters_host, int* bad_flag,
                            const gsmvi_bam_shard* shard, void* stream) {
  if (!stats_workspace || !mu0 || !Sigma0 || !mu_out || !Sigma_out || !solve_workspace || !bad_flag || B <= 0 || D <= 0 ||
      !shard || shard->world < 2 || shard->world > 8 || shard->rank < 0 || shard->rank >= shard->world)
    return GSMVI_EINVAL;
  return bam_solve_full(S(stream), static_cast<const double*>(stats_workspace), B, D, B_total, mu0, Sigma0, lds0, reg, jitter, mu_out,
                        Sigma_out, ldso, static_cast<double*>(solve_workspace), max_ns_iters, ns_iters_host, bad_flag,
                        shard->world, 2, shard);
}

int gsmvi_bam_solve_lowrank(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0,
                            long long lds0, double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso,
                            void* solve_workspace, int max_ns_iters, int* ns_iters_host, int* bad_flag, void* stream) {
  if (!stats_workspace || !mu0 || !Sigma0 || !mu_out || !Sigma_out || !solve_workspace || !bad_flag || B <= 0 || D <= 0)
    return GSMVI_EINVAL;
  return bam_solve_lowrank(S(stream), static_cast<const double*>(stats_workspace), B, D, B_total, mu0, Sigma0, lds0, reg,
                           jitter, mu_out, Sigma_out, ldso, static_cast<double*>(solve_workspace), max_ns_iters,
                           ns_iters_host, bad_flag);
}

int gsmvi_gauss_logq_reduce(const float* Z_or_X, long long ld, int N, int D, const float* mu, const float* L,
                            long long ldl, int from_z, double* out, void* stream) {
  if (from_z) return gauss_logq_from_z(S(stream), Z_or_X, ld, N, D, L, ldl, out);
  return gauss_logq_from_x(S(stream), Z_or_X, ld, N, D, mu, L, ldl, out);
}

int gsmvi_gsm_ensemble_fit(const float* P, const float* c, float* mu, float* Sigma, int F, int D, int B, int niter,
                           unsigned long long seed, const float* z_tape, int* reverts, int first_fit, void* stream) {
  return gsm_ensemble_fit(S(stream), P, c, mu, Sigma, F, D, B, niter, seed, z_tape, reverts, first_fit);
}

int gsmvi_advi_step(float* L, long long ldl, float* mu, const float* G, long long ldg, const float* Z, long long ldz,
                    float* GtZ, long long ldgz, float* gsum, float* mL, float* vL, float* m_mu, float* v_mu, int B, int D,
                    float lr, float b1, float b2, float eps, int t, int npass, void* stream) {
  return advi_step(S(stream), L, ldl, mu, G, ldg, Z, ldz, GtZ, ldgz, gsum, mL, vL, m_mu, v_mu, B, D, lr, b1, b2, eps, t, npass);
}

int gsmvi_gsm_commit(const int* bad_flag, const int* bad_flag2, int n, const void* const* src_host, void* const* dst_host,
                     const long long* bytes_host, int* status, void* stream) {
  return gsm_commit(S(stream), bad_flag, bad_flag2, n, src_host, dst_host, bytes_host, status);
}

long long gsmvi_gsm_small64_workspace_bytes(int B, int D) { return gsm_small64_workspace_bytes(B, D); }

int gsmvi_gsm_small64(int mode, double* mu, double* Sigma, double* L, const float* z_tape, unsigned long long seed,
                      unsigned long long iter0, double* X, const double* G, const double* P, const double* c, int B, int D,
                      int iters, int* status, void* workspace, void* stream) {
  return gsm_small64(S(stream), mode, mu, Sigma, L, z_tape, seed, iter0, X, G, P, c, B, D, iters, status, workspace);
}

}  // extern "C"
