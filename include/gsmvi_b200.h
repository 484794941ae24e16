/* gsmvi_b200.h - C ABI of libgsmvi_b200.so: the B200-native GSM / BaM hot path of modichirag/GSM-VI.
 *
 * The reference has no FFI boundary (its boundary is the Python API); these entry points are
 * what a replacement of its per-iteration host/XLA calls binds.  Each one names the reference code it replaces
 * (paths relative to the reference repository).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions: every pointer is a DEVICE pointer (fp32, row-major, leading dimension in elements, a multiple of 4,
 * base 16-byte aligned) unless its name ends in _host.  Every function is ordered on `stream` (a cudaStream_t),
 * never synchronises, never allocates, and returns 0 on success, <0 for an invalid argument (GSMVI_E*), >0 for a
 * cudaError_t.  `npass` selects tensor-core precision: 3 = 3xTF32 with round-to-nearest split (fp32-grade, default),
 * 2 = 3xTF32 with truncation split (faster, ~8x larger product error), 1 = single-pass TF32.
 */
#ifndef GSMVI_B200_H
#define GSMVI_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define GSMVI_ABI_VERSION 1

#define GSMVI_OK 0
#define GSMVI_EINVAL (-1)     /* bad size / null pointer */
#define GSMVI_EALIGN (-2)     /* pointer or leading dimension not 16-byte aligned */
#define GSMVI_EDRIVER (-3)    /* cuTensorMapEncodeTiled unavailable / failed */
#define GSMVI_EWORKSPACE (-4)

/* workspace kinds for gsmvi_workspace_bytes */
#define GSMVI_WS_POTRF 1
#define GSMVI_WS_GSM_UPDATE 2
#define GSMVI_WS_BAM_STATS 3
#define GSMVI_WS_BAM_SOLVE 4
#define GSMVI_WS_BAM_SOLVE_LOWRANK 5
#define GSMVI_WS_GSM_UPDATE_H3 6
#define GSMVI_WS_POTRF_H3 7

/* One operand of the scaled 3xFP16 engine: fp16 arrays hi / lo (row-major, leading dimension ld in elements, a multiple
 * of 8) and the device float holding the power-of-two scale they were stored with (see gsmvi_h3_split). */
typedef struct gsmvi_h3_operand {
  void* hi;
  void* lo;
  float* scale;
  long long ld;
} gsmvi_h3_operand;

/* Layout of a rank's peer-mapped exchange buffer (offsets in 4-byte words from its base), gsmvi_comm_layout_bytes. */
typedef struct gsmvi_comm_layout {
  long long stage_off;  /* staging: world x tpo dense 128 x 128 partial tiles, slot [source rank][owned tile] */
  long long s_off[2];   /* the two Sigma buffers (D x lds), used alternately as current / next */
  long long dmu_off;    /* world x lds: every rank's mean increment */
  long long cnt_off;    /* unsigned counters: [0, tpo) partial arrivals per owned tile, [tpo] final tiles, [tpo+1] mean */
  long long lds;
  int tiles_m, tpo;
} gsmvi_comm_layout;

int gsmvi_abi_version(void);

/* Bytes of device scratch the call of that kind needs for batch B, dimension D (-1: unknown kind). */
long long gsmvi_workspace_bytes(int kind, int B, int D);

/* General tensor-core contraction C[M,N] = alpha * op(A) op(B)^T + beta * Cin + bias_n (diagnostics and tests; the
 * engine under every call below).  a_mn/b_mn: operand stored [K, rows] instead of [rows, K].  A_lo / B_lo: optional
 * pre-split low parts (see gsmvi_tf32_split; B_lo alone, or both; the operand itself is then the hi part). */
int gsmvi_gemm_tf32(const float* A, long long a_rows, long long a_cols, long long lda, int a_mn, const float* B,
                    long long b_rows, long long b_cols, long long ldb, int b_mn, float* C, long long ldc, int M, int N,
                    int K, float alpha, float beta, const float* Cin, long long ldcin, const float* bias_n, int npass,
                    int tri, int mirror, int krange, int neg_from, const float* A_lo, const float* B_lo, void* stream);

/* Scaled 3xFP16 contraction (same contract as gsmvi_gemm_tf32 at twice the tensor-pipe rate): operands arrive pre-split
 * as fp16 pairs A_hi = rn_f16(A s), A_lo = rn_f16((A s - A_hi) 2^11) with a power-of-two scale s held in a device float
 * (gsmvi_h3_split); fp16 products are exact in the fp32 accumulator, so hi*hi + hi*lo + lo*hi carries 22 significand
 * bits like 3xTF32.  Leading dimensions of the fp16 arrays are in elements, multiples of 8.
 * absmax_out: optional device word, atomicMax of the bit patterns of |C| (feeds the next gsmvi_h3_split; zero it first).
 * splits > 1: split-K, split s writes its raw partial product to C + s * split_stride (beta, bias, mirror not allowed). */
int gsmvi_gemm_h3(const void* A_hi, const void* A_lo, const float* scale_a, long long a_rows, long long a_cols, long long lda,
                  int a_mn, const void* B_hi, const void* B_lo, const float* scale_b, long long b_rows, long long b_cols,
                  long long ldb, int b_mn, float* C, long long ldc, int M, int N, int K, float alpha, float beta,
                  const float* Cin, long long ldcin, const float* bias_n, int tri, int mirror, int krange,
                  unsigned* absmax_out, int splits, long long split_stride, void* stream);

/* Which kernel serves gsmvi_gemm_h3 and the GSM entry points built on it (gsmvi_sample_h3, gsmvi_gauss_score_h3,
 * gsmvi_gsm_update_h3: gsmvi/gsm.py:11-27, 53-54, 119) when both can express the launch: 1 = the persistent 2-CTA kernel
 * (tcgen05.mma.cta_group::2, 256 x 128 tiles, double-buffered TMEM chunks; default), 0 = the one-CTA 128 x 128 kernel,
 * any other value = query only.  Returns the previous setting.  Host-only, no device work. */
int gsmvi_h3_pair_kernel(int enable);

/* *absmax <- max(*absmax, max |A|) as a float bit pattern (device word; zero it first). */
int gsmvi_h3_absmax(const float* A, long long lda, int rows, int cols, unsigned* absmax, void* stream);

/* The fp16 split of an operand: *scale_out <- 2^(14 - e) where *absmax = f 2^e (sqrt_mode: of sqrt(*absmax), the bound
 * on a Cholesky factor's entries given max Sigma_ii), A_hi / A_lo as above. */
int gsmvi_h3_split(const float* A, long long lda, int rows, int cols, const unsigned* absmax, int sqrt_mode,
                   float* scale_out, void* A_hi, void* A_lo, long long ldo, void* stream);

/* The GSM iteration on the scaled 3xFP16 engine (same reference lines as the fp32-operand calls below).
 * gsmvi_philox_normal_h3: Z written directly as the fp16 pair (fixed scale 2^11); offset_dev (optional device word)
 *   overrides `offset`, so that the launch can sit in a replayed CUDA graph while the counter advances on the device.
 * gsmvi_sample_h3 / gsmvi_gauss_score_h3: as gsmvi_sample / gsmvi_gauss_score with pre-split operands; *absmax_x /
 *   *absmax_g (device words, zeroed by the caller) receive the bit pattern of max |X| / max |G| for the next split.
 * gsmvi_gsm_update_h3: as gsmvi_gsm_update; G is given both as fp32 (row pass) and split (W = G Sigma), Sigma both as
 *   fp32 (epilogue) and split; *absmax_sout receives max |Sigma_out| (lower tiles; Sigma_out is symmetric).
 *   workspace: gsmvi_workspace_bytes(GSMVI_WS_GSM_UPDATE_H3, B, D). */
int gsmvi_philox_normal_h3(const gsmvi_h3_operand* Z, int B, int D, unsigned long long seed, unsigned long long offset,
                           const unsigned long long* offset_dev, void* stream);
int gsmvi_sample_h3(const float* mu, const gsmvi_h3_operand* L, const gsmvi_h3_operand* Z, float* X, long long ldx,
                    unsigned* absmax_x, const gsmvi_h3_operand* X_split, int B, int D, void* stream);
int gsmvi_gauss_score_h3(const gsmvi_h3_operand* X, const gsmvi_h3_operand* P, const float* c, float* G, long long ldg,
                         unsigned* absmax_g, const gsmvi_h3_operand* G_split, int B, int D, void* stream);
/* X_split / G_split (optional): the epilogue also writes the result as the fp16 pair with the scale ALREADY in
 * *X_split->scale - an a-priori bound from gsmvi_h3_bound_scales:  |x| <= max|mu| + zmax sqrt(D) sqrt(max Sigma_ii),
 * |g| <= xbound * pnorm + cmax (pnorm = max_j sum_k |P_kj|, cmax = max|c|) - which saves the max|.| pass and the separate
 * split kernel.  zmax_bits: device word with the bit pattern of max|z| (NULL: zmax_const, 5.9 for the Philox draws). */
int gsmvi_h3_bound_scales(const float* mu, int D, const unsigned* sigma_absmax, const unsigned* zmax_bits, float zmax_const,
                          float pnorm, float cmax, float* scale_x, float* scale_g, void* stream);
int gsmvi_gsm_update_h3(const float* X, long long ldx, const float* G, long long ldg, const gsmvi_h3_operand* G_split,
                        const float* mu, const float* Sigma, long long lds, const gsmvi_h3_operand* Sigma_split,
                        float* mu_out, float* Sigma_out, long long ldso, unsigned* absmax_sout, int B, int D, int B_total,
                        int mode, void* workspace, void* stream);

/* Multi-GPU (one process per GPU): batch statistics exchanged through NVLink peer memory, fused with the covariance GEMM
 * instead of an all-reduce after it (replaces nothing in the reference, which is single-device).
 * gsmvi_comm_layout_bytes: fill *lay for dimension D and `world` ranks, return the buffer size in bytes.
 * gsmvi_comm_alloc: cudaMalloc + zero + CUDA IPC handle (64 bytes, to be exchanged between the processes);
 * gsmvi_comm_open / _close: map / unmap a peer's buffer; gsmvi_comm_free: release the own one.
 * gsmvi_gsm_update_h3_fused: gsmvi_gsm_update_h3 for a batch shard with the exchange fused in: the covariance GEMM's
 *   epilogue pushes each partial tile to its owner rank, the owner adds the partials in rank order to the current Sigma
 *   (buffer `cur` inside its comm buffer) and stores the new tile into buffer 1 - cur of every rank; mu_out = mu + the
 *   summed mean increments (only the lower triangle travels; the upper one is mirrored locally).  peer_base: DEVICE
 *   array of `world` comm-buffer pointers (own buffer at [rank], also passed as the plain pointer own_base); `step` must
 *   increase by one per call on every rank (counters are monotonic).  Every rank must make the same sequence of calls. */
long long gsmvi_comm_layout_bytes(int D, int world, gsmvi_comm_layout* lay);
int gsmvi_comm_alloc(long long bytes, void** dev_ptr_out, unsigned char* handle64_out);
int gsmvi_comm_open(const unsigned char* handle64, void** dev_ptr_out);
int gsmvi_comm_close(void* peer_ptr);
int gsmvi_comm_free(void* dev_ptr);
int gsmvi_gsm_update_h3_fused(const float* X, long long ldx, const float* G, long long ldg, const gsmvi_h3_operand* G_split,
                              const float* mu, const gsmvi_h3_operand* Sigma_split, float* mu_out, void* const* peer_base,
                              void* own_base, const gsmvi_comm_layout* lay, int rank, int world, int cur, unsigned step, int B, int D,
                              int B_total, void* workspace, void* stream);

/* L <- chol(Sigma) (lower, upper triangle zeroed), *bad_flag <- 0 if Sigma is positive definite else 1.
 * Replaces GSM._check_goodness / BaM._check_goodness (gsmvi/gsm.py:136-150, gsmvi/bam.py:219-233: host
 * np.linalg.cholesky) and the factorisation inside the sampler (gsmvi/gsm.py:119, gsmvi/bam.py:193). */
int gsmvi_potrf_check(const float* Sigma, long long lds, float* L, long long ldl, int D, int* bad_flag,
                      void* workspace, int npass, void* stream);

/* Same contract as gsmvi_potrf_check on the scaled 3xFP16 engine: left-looking panels (one long-K split-K update GEMM +
 * one panel kernel each; for 512 <= D <= ~17k one fused launch per panel whose spare CTAs run the next panel's update
 * GEMM and reduce the next diagonal tile, while CTA 0 forms the previous panel's own K = 128 term on its tensor core -
 * environment GSMVI_POTRF_LOOKAHEAD=0 selects the two-launch form, GSMVI_POTRF_LATE_MMA=0 the round-2a form with 16
 * helper CTAs, GSMVI_POTRF_CHAIN=1 flag-chained launches), L also written as the fp16 pair *L_split (scale from
 * max |Sigma_ii|) for the sampler's and the later panels' TMA loads.  zero_upper = 0: the blocks above the diagonal are left untouched (valid for a buffer that
 * was zeroed once).  workspace: gsmvi_workspace_bytes(GSMVI_WS_POTRF_H3, 0, D). */
int gsmvi_potrf_h3(const float* Sigma, long long lds, float* L, long long ldl, const gsmvi_h3_operand* L_split, int D,
                   int* bad_flag, void* workspace, int zero_upper, void* stream);

/* Dry run of gsmvi_potrf_h3's launch loop (the goodness check of gsmvi/gsm.py:136-150) for a D x D matrix on a device
 * with `sms` SMs; no CUDA call is made.  One row of eight ints per 128-column panel is written to rows[max_rows][8]:
 * first column, fused launch (1/0), panel CTAs (diagonal block + row owners), GEMM CTAs, GEMM row tiles, GEMM split-K
 * factor, partial planes the panel reads, helper CTAs.  Returns the number of panels (< 0: error).  The kernels spin
 * on each other inside a launch, so panel CTAs + GEMM CTAs <= sms is the invariant that matters; tests check it here. */
int gsmvi_potrf_h3_plan(int D, int sms, int* rows, int max_rows);

/* Z[B,D] <- N(0,1), Philox4x32-10 keyed by seed, counter (element, offset).  Replaces the host RNG of
 * np.random.seed / np.random.multivariate_normal (gsmvi/gsm.py:117-119). */
int gsmvi_philox_normal(float* Z, long long ldz, int B, int D, unsigned long long seed, unsigned long long offset,
                        void* stream);

/* A_hi = tf32_rn(A), A_lo = tf32_rn(A - A_hi): the round-to-nearest 3xTF32 split of a reused operand, done once.
 * The calls below take an optional X_lo beside an operand X (same shape and leading dimension, NULL = none): when it is
 * given, X must be the matching hi part and TMA loads hi and lo tiles directly instead of splitting inside the GEMM. */
int gsmvi_tf32_split(const float* A, long long lda, float* A_hi, float* A_lo, long long ldo, int rows, int cols,
                     void* stream);

/* X[B,D] = mu + Z L^T : samples of N(mu, L L^T).  Replaces np.random.multivariate_normal (gsmvi/gsm.py:119,
 * gsmvi/bam.py:193, gsmvi/monitors.py:106). */
int gsmvi_sample(const float* mu, const float* L, const float* L_lo, long long ldl, const float* Z, long long ldz,
                 float* X, long long ldx, int B, int D, int npass, void* stream);

/* G[B,D] = -(X - m) P = -X P + c with c = P m: batched score of a dense Gaussian target.  Replaces lp_g of the
 * benchmark targets (examples/example_gsm_numpy.py:24-29, examples/example_gsm.py:34-35). */
int gsmvi_gauss_score(const float* X, long long ldx, const float* P, const float* P_lo, long long ldp, const float* c,
                      float* G, long long ldg, int B, int D, int npass, void* stream);

/* Fused GSM batch update.  Replaces gsm_update (gsmvi/gsm.py:31-58; _gsm_update_single gsmvi/gsm.py:8-28).
 * mode 0: mu_out = mu + mean_b u_b, Sigma_out = Sigma + (D^T D - E^T E)/B_total (B_total == B).
 * mode 1: partial statistics of a batch shard: mu_out = sum_b u_b / B_total, Sigma_out = (D^T D - E^T E)/B_total,
 *         to be summed over shards (all-reduce) and applied with gsmvi_gsm_apply_stats.
 * Sigma_hi / Sigma_lo: optional gsmvi_tf32_split of Sigma (both or neither; same leading dimension) for W = G Sigma.
 * workspace: gsmvi_workspace_bytes(GSMVI_WS_GSM_UPDATE, B, D). Sigma_out must not alias Sigma. */
int gsmvi_gsm_update(const float* X, long long ldx, const float* G, long long ldg, const float* mu,
                     const float* Sigma, const float* Sigma_hi, const float* Sigma_lo, long long lds, float* mu_out,
                     float* Sigma_out, long long ldso, int B, int D, int B_total, int mode, void* workspace, int npass,
                     void* stream);

/* Sigma_out = Sigma + dSigma, mu_out = mu + dmu (after the all-reduce of mode-1 statistics). */
int gsmvi_gsm_apply_stats(const float* Sigma, long long lds, const float* dSigma, long long ldd, const float* mu,
                          const float* dmu, float* Sigma_out, long long ldso, float* mu_out, int D, void* stream);

/* fp64 contraction C = alpha * op(A) op(B)^T + beta * Cin + diag_add * I (the engine of the BaM solve; exported for
 * tests and diagnostics). */
int gsmvi_dgemm(const double* A, long long lda, int a_mn, const double* B, long long ldb, int b_mn, double* C,
                long long ldc, int M, int N, int K, double alpha, double beta, const double* Cin, long long ldcin,
                double diag_add, int tri, int mirror, int krange, void* stream);

/* The same fp64 contraction on the int8 tensor cores (Ozaki splitting: `slices` signed 7-bit digits per operand row,
 * 2..8; exact int32 accumulation of every digit-pair product, fp64 recombination; K <= 8192).  The BaM solve uses it for
 * its D^3-sized products (GSMVI_OZ_SLICES=0 in the environment falls back to the FP64-pipe kernel).
 * workspace: gsmvi_dgemm_oz_workspace_bytes(M, N, K, slices) bytes, 1 KiB aligned. */
long long gsmvi_dgemm_oz_workspace_bytes(int M, int N, int K, int slices);
int gsmvi_dgemm_oz(const double* A, long long lda, int a_mn, const double* B, long long ldb, int b_mn, double* C,
                   long long ldc, int M, int N, int K, double alpha, double beta, const double* Cin, long long ldcin,
                   double diag_add, int tri, int mirror, void* workspace, int slices, void* stream);

/* BaM batch statistics.  Replaces gsmvi/bam.py:49-57 (xbar, C, gbar; vmapped outer products there).  Inputs are the
 * fp32 samples / scores; the statistics are fp64 (V = S0 + reg*C multiplies C's rounding by reg, ~100 early on).
 * stats_workspace (gsmvi_workspace_bytes(GSMVI_WS_BAM_STATS, B, D)) layout, in DOUBLES with ld = roundup(D, 8):
 *   [Xc; Gc] (2B x ld) | C (D x ld) | xbar (ld) | gbar (ld).
 * stage 0: xbar, gbar <- column SUMS of this shard (all-reduce them across shards, then)
 * stage 1: xbar, gbar /= B_total; Xc = X - xbar, Gc = G - gbar; C = Xc^T Xc / B_total (shard partial: all-reduce the C
 *          region across shards).  Gamma (gsmvi/bam.py:57) is never formed: the solve uses U's exact factor built from
 *          Gc, so that rounding cannot break U's rank / PSD structure.  npass is ignored (kept for ABI stability). */
int gsmvi_bam_stats(const float* X, long long ldx, const float* G, long long ldg, int B, int D, int B_total,
                    void* stats_workspace, int npass, int stage, void* stream);

/* BaM covariance/mean update from the statistics.  Replaces bam_update's solve (gsmvi/bam.py:59-67, get_sqrt
 * gsmvi/bam.py:19-28: host scipy.linalg.sqrtm) plus the jitter and symmetrisation of BaM.fit (gsmvi/bam.py:198-199):
 *   V = L L^T, M = I + 4 L^T U L (formed as I + 4 W W^T, W = L^T Q, Q Q^T = U exactly), N = M^(1/2) (fp64 scaled
 *   Newton-Schulz), S = 2 L (I+N)^-1 L^T, Sigma_out = S + jitter I,
 *   mu_out = mu0/(1+reg) + reg/(1+reg) (S gbar + xbar).
 * fp64 internally; SYNCHRONISES `stream` once per Newton-Schulz iteration (residual read).  *ns_iters_host <- iterations
 * run (host int); *bad_flag (device int) <- 1 if V or I+N was not positive definite (outputs are then garbage).
 * solve_workspace: gsmvi_workspace_bytes(GSMVI_WS_BAM_SOLVE, B, D) bytes, 16-byte aligned.
 * Sharded batch: world = number of shards, B = this shard's rows.  phase 0 = whole solve (world must be 1);
 * phase 1 = up to this shard's partial M_r = I/world + 4 W_r W_r^T, stored as D x roundup(D,8) doubles at
 * solve_workspace + 3*D*roundup(D,8) doubles - all-reduce (sum) it across shards - then phase 2 = the rest. */
int gsmvi_bam_solve(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0,
                    long long lds0, double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso, void* solve_workspace,
                    int max_ns_iters, int* ns_iters_host, int* bad_flag, int world, int phase, void* stream);

/* fp64 Cholesky used twice inside the BaM solve (V = L L^T and I + N = R R^T; the reference's np.linalg.solve /
 * scipy sqrtm path of gsmvi/bam.py:63-65 is LAPACK fp64): A (n x n doubles, leading dimension lda, lower triangle read) is
 * overwritten by its lower Cholesky factor, upper triangle zeroed; *bad_flag |= 1 on a non-positive pivot.  Exported for
 * tests and timing. */
int gsmvi_potrf64(double* A, long long lda, int n, int* bad_flag, void* stream);

/* Tensor-parallel form of gsmvi_bam_solve's phase 2 for a batch-sharded fit (one process per GPU; SURVEY.md section 8f-1;
 * replaces, like gsmvi_bam_solve, the solve of gsmvi/bam.py:59-67 with its host sqrtm, gsmvi/bam.py:19-28).  Every rank's
 * solve_workspace is a peer-mapped allocation of the same size (gsmvi_comm_alloc / gsmvi_comm_open, zero-initialised);
 * shard->peer_ws[r] is rank r's workspace as mapped into THIS process (own at [rank]).  The three D x D products of every
 * Newton-Schulz iteration, T = L R^-T and S = 2 T T^T are computed by rows of the result across the ranks; each GEMM's
 * epilogue stores its rows into every rank's workspace over NVLink peer memory (the all-gather is fused into the GEMM) and
 * barrier kernels (release / acquire on a counter word inside the workspaces) separate producers from consumers.  Residuals
 * are reduced in a fixed order, so every rank reads the same bits, runs the same number of iterations and ends with
 * bit-identical (mu_out, Sigma_out).  *shard->epoch_host (host word, zero before the first call on these workspaces) counts
 * barriers across calls.  phase must be 2 (after the all-reduce of the phase-1 partials M_r); every rank makes the call. */
typedef struct gsmvi_bam_shard {
  int rank, world;          /* world <= 8 */
  void* peer_ws[8];
  unsigned* epoch_host;
} gsmvi_bam_shard;
int gsmvi_bam_solve_sharded(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0,
                            long long lds0, double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso,
                            void* solve_workspace, int max_ns_iters, int* ns_iters_host, int* bad_flag,
                            const gsmvi_bam_shard* shard, void* stream);

/* Low-rank BaM update (K = B + 1 < D).  Replaces bam_lowrank_update (gsmvi/bam.py:72-114) and compute_Q
 * (gsmvi/bam.py:10-17: host ARPACK svds) with the exact factor Q = [sqrt(reg/B) Gc^T, sqrt(reg/(1+reg)) gbar] of U:
 *   A = V Q, H = A^T Q + I/4, BB = (I/2 + H^(1/2))^2, S = V - A BB^-1 A^T.  Same conventions as gsmvi_bam_solve;
 * workspace kind GSMVI_WS_BAM_SOLVE_LOWRANK. */
int gsmvi_bam_solve_lowrank(const void* stats_workspace, int B, int D, int B_total, const float* mu0, const float* Sigma0,
                            long long lds0, double reg, double jitter, float* mu_out, float* Sigma_out, long long ldso,
                            void* solve_workspace, int max_ns_iters, int* ns_iters_host, int* bad_flag, void* stream);

/* *out (device double) <- sum_b log N(x_b | mu, L L^T), b < N.  Replaces MultivariateNormal(mu, cov).log_prob inside
 * reverse_kl / forward_kl of the monitor (gsmvi/monitors.py:10-22, 107-113).
 * from_z = 1: rows of Z_or_X are the standard-normal draws z_b of samples x_b = mu + L z_b (reverse KL: only |z|^2 and
 *             log det are needed; mu unused);  from_z = 0: rows are arbitrary points x_b (forward KL). */
int gsmvi_gauss_logq_reduce(const float* Z_or_X, long long ld, int N, int D, const float* mu, const float* L,
                            long long ldl, int from_z, double* out, void* stream);

/* Ensemble of F independent GSM fits on dense-Gaussian targets, D <= 64, B <= 32: the complete loop of GSM.fit
 * (gsmvi/gsm.py:107-129: sample, score -(x - m) P, gsm_update gsmvi/gsm.py:31-58, Cholesky check gsmvi/gsm.py:136-150,
 * accept/revert) for niter + 1 iterations inside one kernel, one CTA per fit, state in shared memory.
 * P [F,D,D], c [F,D] (c_f = P_f m_f), mu [F,D] and Sigma [F,D,D] in/out, dense row-major.  z_tape: optional
 * [F, niter+1, B, D] draws (else Philox(seed)).  reverts [F] <- rejected updates (-1: initial Sigma not PD).
 * first_fit: global index of fit 0 (ensembles split across GPUs - replicas only, no communication - keep the Philox
 * streams of the unsharded run). */
int gsmvi_gsm_ensemble_fit(const float* P, const float* c, float* mu, float* Sigma, int F, int D, int B, int niter,
                           unsigned long long seed, const float* z_tape, int* reverts, int first_fit, void* stream);

/* ADVI baseline on the same sampler / score kernels.  One optimiser step of gsmvi/advi.py:69-74:
 * the gradient of neg_elbo (gsmvi/advi.py:31-45; jax.value_and_grad there) with respect to (mu, lower triangle of the
 * scale factor L) from this iteration's draws Z [B,D] and the scores G [B,D] = grad log p at x = mu + Z L^T,
 *   d/dmu = -sum_b g_b,   d/dL_ij = -(G^T Z)_ij - B delta_ij / L_ii  (i >= j),
 * followed by the Adam update of optax.adam (lr, b1, b2, eps; t = 1-based step).  G^T Z runs on the tensor-core GEMM.
 * GtZ [D,ldgz], gsum [D]: scratch.  mL, vL [D,ldl], m_mu, v_mu [D]: Adam moments, zero before the first step. */
int gsmvi_advi_step(float* L, long long ldl, float* mu, const float* G, long long ldg, const float* Z, long long ldz,
                    float* GtZ, long long ldgz, float* gsum, float* mL, float* vL, float* m_mu, float* v_mu, int B, int D,
                    float lr, float b1, float b2, float eps, int t, int npass, void* stream);

/* Accept / revert of an iteration's proposal on the device.  Replaces the host-side branch of gsmvi/gsm.py:125-129 and
 * gsmvi/bam.py:208-212 (`if is_good: mean, cov = mean_new, cov_new else: revert`), which costs the reference a device ->
 * host copy of the covariance and a host Cholesky per iteration.  The caller always exchanges its (current, proposal)
 * buffer pointers; when *bad_flag (the Cholesky's PD flag; optionally also *bad_flag2) is non-zero this call copies the n
 * regions src[r] -> dst[r] (the previous state over the rejected proposal; bytes[r] multiples of 4; src / dst / bytes are
 * HOST arrays of n <= GSMVI_COMMIT_MAX entries holding device pointers), so the exchanged pointers name the old state
 * again.  status (device int[2]): [0] += 1 per rejected update, [1] <- 1 if accepted else 0.  Nothing is read back. */
#define GSMVI_COMMIT_MAX 12
int gsmvi_gsm_commit(const int* bad_flag, const int* bad_flag2, int n, const void* const* src_host, void* const* dst_host,
                     const long long* bytes_host, int* status, void* stream);

/* fp64 GSM iteration for small problems (D <= 64), one CTA, state resident on the device, accept / revert predicated
 * on the device (no host read between iterations).  Replaces the loop body of gsmvi/gsm_numpy.py:107-127 (numpy fp64:
 * np.random.multivariate_normal, lp_g, gsm_update gsm_numpy.py:27-55, np.linalg.cholesky check gsm_numpy.py:131-145) -
 * BASELINE configs[0], examples/example_gsm_numpy.py - and gsmvi/gsm.py:107-129 under jax_enable_x64.
 * State (DOUBLES, dense row-major): mu [D], Sigma [D,D], L [D,D] = chol(Sigma).
 * mode GSMVI_SMALL64_INIT:   L <- chol(Sigma); status[1] <- 1 if Sigma is not positive definite.
 * mode GSMVI_SMALL64_FULL:   `iters` complete iterations with the built-in dense-Gaussian score g = -x P + c
 *                            (P [D,D] symmetric precision, c = P m; examples/example_gsm_numpy.py:24-29).
 * mode GSMVI_SMALL64_SAMPLE: X [B,D] <- mu + Z L^T (gsm.py:117-119) - then the caller evaluates its lp_g on X - and
 * mode GSMVI_SMALL64_UPDATE: the update (gsm.py:122), check (gsm.py:125) and commit from X and the caller's G [B,D].
 * Draws: z_tape [iters, B, D] (fp32, the slice for these iterations) or NULL for Philox4x32-10 (seed, counter iter0 + it).
 * status (device int[3]): [0] += rejected updates, [1] see INIT, [2] <- 1 if the last update was accepted.
 * workspace: gsmvi_gsm_small64_workspace_bytes(B, D) bytes, 8-byte aligned. */
#define GSMVI_SMALL64_INIT 0
#define GSMVI_SMALL64_FULL 1
#define GSMVI_SMALL64_SAMPLE 2
#define GSMVI_SMALL64_UPDATE 3
long long gsmvi_gsm_small64_workspace_bytes(int B, int D);
int gsmvi_gsm_small64(int mode, double* mu, double* Sigma, double* L, const float* z_tape, unsigned long long seed,
                      unsigned long long iter0, double* X, const double* G, const double* P, const double* c, int B, int D,
                      int iters, int* status, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif
