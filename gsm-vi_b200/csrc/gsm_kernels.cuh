// GSM iteration pieces (gsm_kernels.cu): Philox normals, sampling, dense-Gaussian score, fused GSM update.
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

int philox_normal(cudaStream_t stream, float* Z, long long ldz, int B, int D, unsigned long long seed,
                  unsigned long long offset);
// hi = tf32_rn(a), lo = tf32_rn(a - hi).  Passing (hi, lo) as (X, X_lo) below lets the GEMM skip converting that operand.
int tf32_split(cudaStream_t stream, const float* A, long long lda, float* Hi, float* Lo, long long ldo, int rows, int cols);
int sample_mvn(cudaStream_t stream, const float* mu, const float* L, const float* L_lo, long long ldl, const float* Z,
               long long ldz, float* X, long long ldx, int B, int D, int npass);
int gauss_score(cudaStream_t stream, const float* X, long long ldx, const float* P, const float* P_lo, long long ldp,
                const float* c, float* G, long long ldg, int B, int D, int npass);
size_t gsm_update_workspace_bytes(int B, int D);
int gsm_update(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const float* mu,
               const float* Sigma, const float* Sigma_hi, const float* Sigma_lo, long long lds, float* mu_out, float* Sigma_out, long long ldso, int B, int D,
               int B_total, int mode, float* workspace, int npass);
int gsm_apply_stats(cudaStream_t stream, const float* Sigma, long long lds, const float* dSigma, long long ldd,
                    const float* mu, const float* dmu, float* Sigma_out, long long ldso, float* mu_out, int D);

// accept / revert on the device: if *bad (or *bad2) is set, copy src[r] -> dst[r] (bytes[r], multiples of 4) for r < n
int gsm_commit(cudaStream_t stream, const int* bad, const int* bad2, int n, const void* const* src, void* const* dst,
               const long long* bytes, int* status);

// ---- scaled 3xFP16 path (h3_gemm.cuh): operands are gsmvi_h3_operand (fp16 hi / lo + device scale)
typedef gsmvi_h3_operand H3Operand;
int philox_normal_h3(cudaStream_t stream, const H3Operand& Z, int B, int D, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* offset_dev = nullptr);
int sample_mvn_h3(cudaStream_t stream, const float* mu, const H3Operand& L, const H3Operand& Z, float* X, long long ldx,
                  unsigned* absmax_x, int B, int D, const H3Operand* Xsplit = nullptr);
int gauss_score_h3(cudaStream_t stream, const H3Operand& X, const H3Operand& P, const float* c, float* G, long long ldg,
                   unsigned* absmax_g, int B, int D, const H3Operand* Gsplit = nullptr);
size_t gsm_update_h3_workspace_bytes(int B, int D);
int gsm_update_h3(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const H3Operand& Gh,
                  const float* mu, const float* Sigma, long long lds, const H3Operand& Sh, float* mu_out, float* Sigma_out,
                  long long ldso, unsigned* absmax_sout, int B, int D, int B_total, int mode, void* workspace);

int gsm_update_h3_fused(cudaStream_t stream, const float* X, long long ldx, const float* G, long long ldg, const H3Operand& Gh,
                        const float* mu, const H3Operand& Sh, float* mu_out, float* const* peer_base, float* own_base,
                        const gsmvi_comm_layout& lay, int rank, int world, int cur, unsigned step, int B, int D, int B_total,
                        void* workspace);

}  // namespace gsmvi
