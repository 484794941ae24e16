// BaM covariance solve on the device.
//
// Reference (gsmvi/bam.py:59-67):  U = reg*Gamma + reg/(1+reg) gbar gbar^T,  V = S0 + reg*C + reg/(1+reg) (mu0-xbar)(mu0-xbar)^T,
//   S = 2 (I + sqrtm(I + 4UV)^T)^{-1} V^T   (solves S U S + S = V; host scipy.linalg.sqrtm of a non-symmetric matrix),
//   mu = mu0/(1+reg) + reg/(1+reg) (S gbar + xbar).
// Device restatement (oracle/gsmvi_oracle.py bam_update): V = L L^T, M = I + 4 L^T U L (SPD, eigenvalues >= 1),
//   N = M^{1/2} by a GEMM-bound coupled Newton-Schulz iteration, I + N = R R^T, T = L R^{-T}, S = 2 T T^T - symmetric
//   PSD by construction, algebraically identical.  Low-rank variant (bam.py:72-114) with the exact factor
//   Q = [sqrt(reg/B) (G-gbar)^T, sqrt(reg/(1+reg)) gbar] of U instead of a truncated SVD (Q Q^T = U, S is invariant).
// Samples and scores come from the fp32 tensor-core path; the statistics and everything D x D / K x K here are fp64
// (dgemm.cu): V = S0 + reg*C amplifies C's rounding by reg, and the solve mixes scales of 1 and ~1e9.
#include "dev_once.cuh"
#include "bam_solve.cuh"
#include "oz_gemm.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dgemm.cuh"

namespace gsmvi {

constexpr int NB64 = 64;
constexpr int L64S = NB64 + 1;

// ------------------------------------------------------------------------------------------------ small kernels

// fp64 diagonal block (n <= 64): in-place Cholesky, one CTA.  This kernel sits on the critical path of every 64-column
// panel of the fp64 Cholesky (twice per BaM update), so it is built for latency AND instruction count (measured on B200,
// tools/probes/: a dependent DFMA is 23 cycles, rsqrt 83, a shared-memory hand-over through __syncthreads 72 - but a
// per-column right-looking sweep over 256 threads issued ~1200 cycles of mostly predicated-off work per column, 95 us per
// block).  The matrix lives in registers as 4 x 4 blocks (thread (ty, tx) = rows 4 ty.., columns 4 tx..; blocks above the
// diagonal idle) and the factorisation advances a BLOCK column at a time, 16 steps of: the diagonal thread factors its
// 4 x 4 block in registers; the threads below solve their block against it; everybody to the right applies the rank-4
// update.  Two barriers per step, every FMA issued is a useful one.  No inverse is formed: the panel solve is
// trsm64_tile_kernel below.
__global__ void __launch_bounds__(256, 1) potrf64_diag_kernel(double* __restrict__ a, long long lda, int n,
                                                              int* __restrict__ flag) {
  __shared__ double lcol[2][NB64][4];  // the finished block column (rows x 4), double-buffered
  __shared__ double drec[2][4];        // reciprocals of its diagonal entries
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  if (tid == 0) bad = 0;
  double r[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int i = 4 * ty + ii, k = 4 * tx + kk;
      r[ii][kk] = (i < n && k <= i) ? a[static_cast<long long>(i) * lda + k] : ((i == k) ? 1.0 : 0.0);
    }
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < NB64 / 4; ++jb) {
    double (*lc)[4] = lcol[jb & 1];
    double* dr = drec[jb & 1];
    if (ty == jb && tx == jb) {
      // 4 x 4 diagonal block, in registers: four dependent pivots, nothing else waits on anything but these
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double p = r[c][c];
        if (!(p > 0.0) || isinf(p)) bad = 1;
        const double rs = rsqrt(p);
        r[c][c] = p * rs;
        dr[c] = rs;
#pragma unroll
        for (int i = c + 1; i < 4; ++i) r[i][c] *= rs;
#pragma unroll
        for (int k = c + 1; k < 4; ++k)
#pragma unroll
          for (int i = k; i < 4; ++i) r[i][k] -= r[i][c] * r[k][c];
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk > ii) r[ii][kk] = 0.0;
          lc[4 * jb + ii][kk] = r[ii][kk];
        }
    }
    __syncthreads();
    if (tx == jb && ty > jb) {
      // X Ld^T = R for this thread's 4 x 4 block: column c of X from the columns before it
      double ld_[4][4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ld_[ii][kk] = lc[4 * jb + ii][kk];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double rc = dr[c];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          double v = r[ii][c];
#pragma unroll
          for (int k = 0; k < c; ++k) v -= r[ii][k] * ld_[c][k];
          r[ii][c] = v * rc;
        }
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) lc[4 * ty + ii][kk] = r[ii][kk];
    }
    __syncthreads();
    if (tx > jb && ty >= tx) {
      double li[4][4], lk[4][4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          li[ii][c] = lc[4 * ty + ii][c];
          lk[ii][c] = lc[4 * tx + ii][c];
        }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          double v = r[ii][kk];
#pragma unroll
          for (int c = 0; c < 4; ++c) v -= li[ii][c] * lk[kk][c];
          r[ii][kk] = v;
        }
    }
    // (no barrier here: the next step writes the other buffer, and the barrier after ITS diagonal phase orders every
    // reader of this one before the step after that overwrites it)
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int i = 4 * ty + ii, k = 4 * tx + kk;
      if (i < n && k < n) a[static_cast<long long>(i) * lda + k] = (k <= i) ? r[ii][kk] : 0.0;
    }
  __syncthreads();
  if (tid == 0 && bad) atomicOr(flag, 1);
}

// X <- X Ld^-T for a 64-row tile of X per CTA (X: m x nb, nb <= 64; Ld: nb x nb lower triangular, both fp64 row-major): the
// panel solve of the blocked Cholesky (L21 = A21 L11^-T) and the block step of the blocked triangular solve, without an
// explicit inverse.  Same register blocking as the diagonal kernel: 16 block-column steps of (threads of block column jb
// solve their 4 x 4 block against Ld's diagonal block and publish it; threads to the right subtract its contribution),
// one barrier per step, row tiles independent across CTAs.
__global__ void __launch_bounds__(256) trsm64_tile_kernel(double* __restrict__ X, long long ldx, int m,
                                                          const double* __restrict__ Ld, long long ldl, int nb) {
  __shared__ double ls[NB64][NB64 + 1];
  __shared__ double xcol[2][NB64][4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const long long r0 = static_cast<long long>(blockIdx.x) * NB64;
  for (int idx = tid; idx < NB64 * NB64; idx += 256) {
    const int i = idx / NB64, k = idx % NB64;
    ls[i][k] = (i < nb && k <= i) ? Ld[static_cast<long long>(i) * ldl + k] : ((i == k) ? 1.0 : 0.0);
  }
  double r[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const long long i = r0 + 4 * ty + ii;
      const int k = 4 * tx + kk;
      r[ii][kk] = (i < m && k < nb) ? X[i * ldx + k] : 0.0;
    }
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < NB64 / 4; ++jb) {
    double (*xc)[4] = xcol[jb & 1];
    if (tx == jb) {
      double ld_[4][4], rc[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ld_[ii][kk] = ls[4 * jb + ii][4 * jb + kk];
        rc[ii] = 1.0 / ld_[ii][ii];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          double v = r[ii][c];
#pragma unroll
          for (int k = 0; k < c; ++k) v -= r[ii][k] * ld_[c][k];
          r[ii][c] = v * rc[c];
        }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) xc[4 * ty + ii][kk] = r[ii][kk];
    }
    __syncthreads();
    if (tx > jb) {
      double xi[4][4], lk[4][4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xi[ii][c] = xc[4 * ty + ii][c];
          lk[ii][c] = ls[4 * tx + ii][4 * jb + c];
        }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          double v = r[ii][kk];
#pragma unroll
          for (int c = 0; c < 4; ++c) v -= xi[ii][c] * lk[kk][c];
          r[ii][kk] = v;
        }
    }
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const long long i = r0 + 4 * ty + ii;
      const int k = 4 * tx + kk;
      if (i < m && k < nb) X[i * ldx + k] = r[ii][kk];
    }
}

// zero the strict upper triangle (and optionally symmetrise from the lower one)
__global__ void tril64_kernel(double* __restrict__ A, long long lda, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < n && j > i) A[i * lda + j] = 0.0;
}

// V64 = S0 + reg*C + w dm dm^T   (w = reg/(1+reg), dm = mu0 - xbar)   bam.py:60.   U (bam.py:59) is never formed: it
// only enters through its exact factor Q (bam_build_q_kernel), which keeps U's rank / PSD structure out of reach of
// rounding (rounding Gamma's entries independently perturbs the unit eigenvalues of I + 4 L^T U L by ~4 reg |L|^2 eps).
__global__ void bam_v_kernel(const double* __restrict__ C, long long ldc, const float* __restrict__ S0, long long lds,
                             const double* __restrict__ xbar, const float* __restrict__ mu0, double reg,
                             double* __restrict__ V, long long ldv, int D) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j >= D) return;
  const double w = reg / (1.0 + reg);
  const double di = static_cast<double>(mu0[i]) - xbar[i], dj = static_cast<double>(mu0[j]) - xbar[j];
  V[i * ldv + j] = static_cast<double>(S0[i * lds + j]) + reg * C[i * ldc + j] + w * di * dj;
}

// A = scale * A (+ diag_add on the diagonal)
// B = scale * A + diag_add * I (out of place)
__global__ void affine_diag64_kernel(const double* __restrict__ A, double* __restrict__ Bm, long long ld, int n, double scale,
                                     double diag_add) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < n) Bm[i * ld + j] = scale * A[i * ld + j] + ((i == j) ? diag_add : 0.0);
}

__global__ void scale_diag64_kernel(double* __restrict__ A, long long lda, int n, double scale, double diag_add) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < n) A[i * lda + j] = scale * A[i * lda + j] + ((i == j) ? diag_add : 0.0);
}

__global__ void or_flag_kernel(int* flag, int bits) { atomicOr(flag, bits); }

__global__ void set_identity64_kernel(double* __restrict__ A, long long lda, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < n) A[i * lda + j] = (i == j) ? 1.0 : 0.0;
}

// out[0] += sum_ij (A[i][j] - d*delta_ij)^2 ; out[1] = max(out[1], max_i sum_j |A[i][j]|)  (inf-norm, via atomics on bits)
__global__ void frob_inf64_kernel(const double* __restrict__ A, long long lda, int n, double d, double* __restrict__ out) {
  const long long i = blockIdx.x;
  double ss = 0.0, rs = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = A[i * lda + j];
    const double e = v - ((i == j) ? d : 0.0);
    ss += e * e;
    rs += fabs(v);
  }
  __shared__ double s1[256], s2[256];
  s1[threadIdx.x] = ss;
  s2[threadIdx.x] = rs;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(out, s1[0]);
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(out + 1), static_cast<unsigned long long>(__double_as_longlong(s2[0])));
  }
}

// Deterministic form of the reduction above (same quantities): one CTA per row writes the row's sum of squares and
// absolute row sum to rowbuf[2 i], rowbuf[2 i + 1]; a single CTA then adds / maximises them in a fixed order.  Replicated
// ranks of a sharded solve take their stopping decisions from this number, so it must not depend on atomic ordering.
__global__ void frob_rows64_kernel(const double* __restrict__ A, long long lda, int n, double d, double* __restrict__ rowbuf) {
  const long long i = blockIdx.x;
  double ss = 0.0, rs = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = A[i * lda + j];
    const double e = v - ((i == j) ? d : 0.0);
    ss += e * e;
    rs += fabs(v);
  }
  __shared__ double s1[256], s2[256];
  s1[threadIdx.x] = ss;
  s2[threadIdx.x] = rs;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    rowbuf[2 * i] = s1[0];
    rowbuf[2 * i + 1] = s2[0];
  }
}
__global__ void frob_final64_kernel(const double* __restrict__ rowbuf, int n, double* __restrict__ out) {
  double ss = 0.0, mx = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    ss += rowbuf[2 * i];
    mx = fmax(mx, rowbuf[2 * i + 1]);  // NaN rows: fmax drops NaN, the sum of squares keeps it
  }
  __shared__ double s1[256], s2[256];
  s1[threadIdx.x] = ss;
  s2[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] = fmax(s2[threadIdx.x], s2[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = s1[0];
    out[1] = (s1[0] != s1[0]) ? s1[0] : s2[0];  // a NaN anywhere poisons the norm too
  }
}

// ---- tensor-parallel solve across the ranks of a sharded fit: the D x D products of the
// Newton-Schulz iteration (bam.py:63, get_sqrt bam.py:19-28) are split by rows of the result, each rank's GEMM epilogue
// stores its rows into every rank's copy of the result (dgemm.cu, peer memory over NVLink), and a barrier separates a
// product from its consumers.  Every rank therefore holds bit-identical iterates and takes identical decisions.
struct ShardPtrs {
  unsigned* cnt[DGEMM_MAX_PEERS];  // the barrier word inside every rank's solve workspace
};
__device__ __forceinline__ unsigned ld_acquire_sys64(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Every rank adds one to every rank's barrier word and waits until its own has seen all `world` arrivals of this barrier
// (the word only grows: target = world * number of barriers so far).  The peers' stores of the kernels before were fenced
// system-wide by their epilogues; the release / acquire pair orders them before everything after the barrier.
__global__ void ns_barrier_kernel(ShardPtrs p, int rank, int world, unsigned target, long long timeout_cycles) {
  if (threadIdx.x < world) {
    __threadfence_system();
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p.cnt[threadIdx.x]), "r"(1u) : "memory");
  }
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys64(p.cnt[rank]) - target) < 0) {
      __nanosleep(200);
      if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
        printf("gsmvi: BaM solve barrier watchdog (rank %d, have %u want %u)\n", rank, ld_acquire_sys64(p.cnt[rank]), target);
        __trap();
      }
    }
  }
}
// rows [row0, row0 + rows) of an n-column fp64 matrix (leading dimension ld, ld even) from this rank's copy into the same
// place of every peer's copy: 16-byte stores, consecutive threads on consecutive addresses
struct BcastArgs {
  const double* src;
  double* dst[DGEMM_MAX_PEERS];
  int ndst;
  long long count2;  // number of double2 elements
};
__global__ void __launch_bounds__(256) peer_bcast_kernel(const BcastArgs a) {
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  const double2* s2 = reinterpret_cast<const double2*>(a.src);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.count2; i += nth) {
    const double2 v = s2[i];
    for (int d = 0; d < a.ndst; ++d) reinterpret_cast<double2*>(a.dst[d])[i] = v;
  }
  __threadfence_system();
}

// S32 = (float)(scale*S64) + jitter*I, symmetric by construction of S64 ; bam.py:198-199
__global__ void bam_finish_cov_kernel(const double* __restrict__ S64, long long lds64, float* __restrict__ S32,
                                      long long lds32, int D, double scale, double jitter) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j < D) {
    // (cov + cov^T)/2 of bam.py:199: S64 is already exactly symmetric (mirrored stores), so this is the identity
    const double v = 0.5 * (scale * S64[i * lds64 + j] + scale * S64[j * lds64 + i]) + ((i == j) ? jitter : 0.0);
    S32[i * lds32 + j] = static_cast<float>(v);
  }
}

// mu_out = mu0/(1+reg) + reg/(1+reg) (S gbar + xbar)   bam.py:67 ; one warp per row, fp64 S
__global__ void bam_mean_kernel(const double* __restrict__ S64, long long lds, double scale, const double* __restrict__ gbar,
                                const double* __restrict__ xbar, const float* __restrict__ mu0, double reg,
                                float* __restrict__ mu_out, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= D) return;
  double acc = 0.0;
  for (int j = lane; j < D; j += 32) acc += S64[static_cast<long long>(row) * lds + j] * gbar[j];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0)
    mu_out[row] = static_cast<float>(static_cast<double>(mu0[row]) / (1.0 + reg) +
                                     reg / (1.0 + reg) * (scale * acc + xbar[row]));
}

// column sums of X and G in fp64: sx[j] += sum_b X[b][j] (over this CTA's rows)
constexpr int CS_ROWS = 32;
__global__ void colsum2_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ G, long long ldg,
                               int B, int D, double* __restrict__ sx, double* __restrict__ sg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * CS_ROWS;
  if (j >= D) return;
  double ax = 0.0, ag = 0.0;
  const int r1 = min(r0 + CS_ROWS, B);
  for (long long b = r0; b < r1; ++b) {
    ax += X[b * ldx + j];
    ag += G[b * ldg + j];
  }
  atomicAdd(sx + j, ax);
  atomicAdd(sg + j, ag);
}

// T = [X - xbar ; G - gbar]  ([2B, D] row-major, fp64)
__global__ void center2_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ G, long long ldg,
                               int B, int D, const double* __restrict__ xbar, const double* __restrict__ gbar,
                               double* __restrict__ T, long long ldt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (j >= D) return;
  T[b * ldt + j] = static_cast<double>(X[b * ldx + j]) - xbar[j];
  T[(b + B) * ldt + j] = static_cast<double>(G[b * ldg + j]) - gbar[j];
}

__global__ void scale_vec2_kernel(double* __restrict__ a, double* __restrict__ b, double s, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    a[j] *= s;
    b[j] *= s;
  }
}

// Q64[i][k] = sqrt(reg/Btot) * Gc[k][i]  (k < B) ; Q64[i][B] = sqrt(reg/(1+reg)) * gbar[i]   -> Q stored [D, K] row-major
// With the batch sharded over `world` ranks every rank builds the columns of its own samples and the gbar column scaled
// by 1/sqrt(world), so that the shard Gram matrices sum to U.
__global__ void bam_build_q_kernel(const double* __restrict__ Gc, long long ldgc, const double* __restrict__ gbar, int B,
                                   int D, double reg, int Btot, int world, double* __restrict__ Q, long long ldq) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (k > B) return;
  Q[i * ldq + k] = (k < B) ? sqrt(reg / Btot) * Gc[static_cast<long long>(k) * ldgc + i]
                           : sqrt(reg / ((1.0 + reg) * world)) * gbar[i];
}

// ------------------------------------------------------------------------------------------------ fp64 building blocks

static inline dim3 grid2(int n, int rows) { return dim3((n + 255) / 256, rows); }
static inline cudaError_t last() { return cudaGetLastError(); }
#define GSMVI_TRY(x)                \
  do {                              \
    int rc__ = (x);                 \
    if (rc__ != GSMVI_OK) return rc__; \
  } while (0)
#define GSMVI_CUDA(x)                                          \
  do {                                                         \
    cudaError_t e__ = (x);                                     \
    if (e__ != cudaSuccess) return static_cast<int>(e__);      \
  } while (0)

// The large fp64 products of the solve CAN run on the int8 tensor cores (oz_gemm.cuh): GSMVI_OZ_SLICES = 2..8 in the
// environment picks the digit count.  It is OFF by default: the solve amplifies product errors by ~1e8, and the
// fixed-point-per-row error of the digit split (9e-16 normwise at 8 digits) costs one to two digits against true fp64
// (single update, D = 2048, kappa = 1e2: relF 1.4e-7 vs 2.7e-8; D = 1024, kappa = 1e4: 2.7e-4 vs 2.9e-6, over the
// 1e-4 bar) for a 1.4x faster product (profiles/r01_oz_probe.json).
struct OzCtx {
  void* ws = nullptr;
  int slices = 0;
};
static int oz_slices_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GSMVI_OZ_SLICES");
    v = e ? atoi(e) : 0;
    if (v != 0 && (v < 2 || v > 8)) v = 8;
  }
  return v;
}
static int dgemm_big(cudaStream_t st, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                     long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o, const OzCtx& oz) {
  if (oz.ws && oz.slices >= 2 && M >= 1024 && N >= 1024 && K >= 512 && K <= 8192)
    return launch_dgemm_oz(st, M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, o, oz.ws, oz.slices);
  return launch_dgemm(st, M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, o);
}

// Helper stream + events of the look-ahead below, one set per device, created on first use.
struct LookAhead {
  cudaStream_t side = nullptr;
  cudaEvent_t slab = nullptr, rest = nullptr;
  // row-chunked triangular solve: three more streams, a fork event and one join event per stream
  cudaStream_t chunk[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
};
static int lookahead_get(LookAhead** out) {
  static LookAhead per_dev[64];
  int d = 0;
  cudaGetDevice(&d);
  if (d < 0 || d >= 64) d = 0;
  LookAhead& la = per_dev[d];
  if (!la.side) {
    GSMVI_CUDA(cudaStreamCreateWithFlags(&la.side, cudaStreamNonBlocking));
    GSMVI_CUDA(cudaEventCreateWithFlags(&la.slab, cudaEventDisableTiming));
    GSMVI_CUDA(cudaEventCreateWithFlags(&la.rest, cudaEventDisableTiming));
    GSMVI_CUDA(cudaEventCreateWithFlags(&la.fork, cudaEventDisableTiming));
    for (int c = 0; c < 3; ++c) {
      GSMVI_CUDA(cudaStreamCreateWithFlags(&la.chunk[c], cudaStreamNonBlocking));
      GSMVI_CUDA(cudaEventCreateWithFlags(&la.join[c], cudaEventDisableTiming));
    }
  }
  *out = &la;
  return GSMVI_OK;
}

// In-place blocked Cholesky of the lower triangle of A (n x n fp64); upper triangle zeroed.  flag |= 1 on a bad pivot.
// Right-looking over 64-column panels with a one-panel look-ahead: the trailing update of panel k is split into the slab
// that panel k+1 consists of (on `st`, the critical path) and the rest of the trailing matrix (on a helper stream), so the
// diagonal kernel and the panel solve of k+1 run while the bulk of update k is still in flight.  The chain per panel is
// then diagonal block + panel solve + slab instead of + the whole HBM-bound update (9.2 -> ~4 ms at n = 4096).
static int potrf64_inplace(cudaStream_t st, double* A, long long lda, int n, int* flag) {
  LookAhead* la = nullptr;
  const bool look = n >= 8 * NB64;
  if (look) GSMVI_TRY(lookahead_get(&la));
  bool rest_pending = false;
  tril64_kernel<<<grid2(n, n), 256, 0, st>>>(A, lda, n);
  for (int j0 = 0; j0 < n; j0 += NB64) {
    const int nb = min(NB64, n - j0), rest = n - j0 - nb;
    double* a11 = A + static_cast<long long>(j0) * lda + j0;
    potrf64_diag_kernel<<<1, 256, 0, st>>>(a11, lda, nb, flag);
    if (rest > 0) {
      double* a21 = A + static_cast<long long>(j0 + nb) * lda + j0;
      double* a22 = A + static_cast<long long>(j0 + nb) * lda + j0 + nb;
      trsm64_tile_kernel<<<(rest + NB64 - 1) / NB64, 256, 0, st>>>(a21, lda, rest, a11, lda, nb);  // L21 = A21 L11^-T
      DgemmOpts s;  // A22 -= L21 L21^T (lower)
      s.alpha = -1.0;
      s.beta = 1.0;
      s.ldcin = lda;
      const int slab = min(NB64, rest);
      if (!look || rest <= 2 * NB64) {
        if (rest_pending) {  // the previous panel's bulk update must have landed before this one adds to the same blocks
          GSMVI_CUDA(cudaStreamWaitEvent(st, la->rest, 0));
          rest_pending = false;
        }
        s.Cin = a22;
        s.tri = true;
        GSMVI_TRY(launch_dgemm(st, rest, rest, nb, a21, lda, false, a21, lda, false, a22, lda, s));
      } else {
        // (a) the next panel's columns: rows j0+nb .. n, columns j0+nb .. j0+nb+slab  (after the previous bulk update,
        //     which also touched them)
        if (rest_pending) GSMVI_CUDA(cudaStreamWaitEvent(st, la->rest, 0));
        s.Cin = a22;
        GSMVI_TRY(launch_dgemm(st, rest, slab, nb, a21, lda, false, a21, lda, false, a22, lda, s));
        GSMVI_CUDA(cudaEventRecord(la->slab, st));
        // (b) everything to the right of the slab, lower triangle, on the helper stream: rows / columns j0+nb+slab .. n
        const int r2 = rest - slab;
        double* b21 = a21 + static_cast<long long>(slab) * lda;
        double* b22 = a22 + static_cast<long long>(slab) * lda + slab;
        GSMVI_CUDA(cudaStreamWaitEvent(la->side, la->slab, 0));
        DgemmOpts s2 = s;
        s2.Cin = b22;
        s2.tri = true;
        GSMVI_TRY(launch_dgemm(la->side, r2, r2, nb, b21, lda, false, b21, lda, false, b22, lda, s2));
        GSMVI_CUDA(cudaEventRecord(la->rest, la->side));
        rest_pending = true;
      }
    }
  }
  if (rest_pending) GSMVI_CUDA(cudaStreamWaitEvent(st, la->rest, 0));
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// T (m x n) <- Bm R^{-T}  i.e. solve T R^T = Bm for lower-triangular R (n x n).
// Bm and T may alias.  Right-looking over 64-column blocks: T_j <- T_j R_jj^-T (trsm64_tile_kernel), then the trailing columns
// T[:, j0+64:] -= T_j R[j0+64:, j]^T.  (The left-looking form - one m x 64 x j0 product per block - has only m / 64 tiles
// per launch and a K loop as long as the matrix: 8.2 ms for 2048 x 4096 against ~2 ms for this one, whose launches are
// wide, K = 64 products.)
static int trsm64_right_lt(cudaStream_t st, const double* Bm, long long ldb, const double* R, long long ldr, double* T,
                           long long ldt, int m, int n) {
  if (T != Bm)
    GSMVI_CUDA(cudaMemcpy2DAsync(T, ldt * sizeof(double), Bm, ldb * sizeof(double), n * sizeof(double), m,
                                 cudaMemcpyDeviceToDevice, st));
  for (int j0 = 0; j0 < n; j0 += NB64) {
    const int nb = min(NB64, n - j0), rest = n - j0 - nb;
    double* tj = T + j0;
    trsm64_tile_kernel<<<(m + NB64 - 1) / NB64, 256, 0, st>>>(tj, ldt, m, R + static_cast<long long>(j0) * ldr + j0, ldr, nb);
    if (rest > 0) {
      DgemmOpts o;
      o.alpha = -1.0;
      o.beta = 1.0;
      o.Cin = tj + nb;
      o.ldcin = ldt;
      GSMVI_TRY(launch_dgemm(st, m, rest, nb, tj, ldt, false, R + static_cast<long long>(j0 + nb) * ldr + j0, ldr, false, tj + nb,
                             ldt, o));
    }
  }
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// The rows of T = Bm R^{-T} are independent and every launch of the solve above is a short, latency-bound kernel (64
// dependent steps), so a tall solve is cut into four row chunks that run the same chain concurrently on four streams: the
// chain is as long as before, four of them overlap (7.2 -> ~2.5 ms for 4096 x 4096).
static int trsm64_right_lt_chunked(cudaStream_t st, const double* Bm, long long ldb, const double* R, long long ldr, double* T,
                                   long long ldt, int m, int n) {
  if (m < 1024) return trsm64_right_lt(st, Bm, ldb, R, ldr, T, ldt, m, n);
  LookAhead* la = nullptr;
  GSMVI_TRY(lookahead_get(&la));
  const int per = ((m + 3) / 4 + NB64 - 1) / NB64 * NB64;
  GSMVI_CUDA(cudaEventRecord(la->fork, st));
  for (int c = 0; c < 4; ++c) {
    const int r0 = c * per;
    if (r0 >= m) break;
    const int rows = min(per, m - r0);
    cudaStream_t sc = c == 0 ? st : la->chunk[c - 1];
    if (c > 0) GSMVI_CUDA(cudaStreamWaitEvent(sc, la->fork, 0));
    GSMVI_TRY(trsm64_right_lt(sc, Bm + static_cast<long long>(r0) * ldb, ldb, R, ldr, T + static_cast<long long>(r0) * ldt, ldt,
                              rows, n));
    if (c > 0) {
      GSMVI_CUDA(cudaEventRecord(la->join[c - 1], sc));
      GSMVI_CUDA(cudaStreamWaitEvent(st, la->join[c - 1], 0));
    }
  }
  return GSMVI_OK;
}

// Debug trace (GSMVI_DEBUG=1): synchronise and print ||A||_F, ||A||_inf and the PD flag after a stage.
static bool dbg_on() {
  static int on = -1;
  if (on < 0) on = getenv("GSMVI_DEBUG") ? 1 : 0;
  return on == 1;
}
static void dbg_stage(cudaStream_t st, const char* name, const double* A, long long ld, int n, double* scal, const int* flag) {
  if (!dbg_on()) return;
  double h[2];
  int f = -1;
  cudaMemsetAsync(scal, 0, 2 * sizeof(double), st);
  frob_inf64_kernel<<<n, 256, 0, st>>>(A, ld, n, 0.0, scal);
  cudaMemcpyAsync(h, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&f, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  fprintf(stderr, "[gsmvi debug] %-22s n=%d frob=%.6e inf=%.6e flag=%d cuda=%d\n", name, n, sqrt(h[0]), h[1], f, (int)e);
}

// Sharding context of a solve: world == 1 is the single-GPU solve; otherwise peer_ws[r] is rank r's solve workspace
// (same layout on every rank, own at [rank]) and *epoch_host counts the barriers this workspace has been through.
struct ShardCtx {
  int rank = 0, world = 1;
  double* peer_ws[DGEMM_MAX_PEERS] = {};
  double* own_ws = nullptr;
  long long cnt_off = 0;  // offset (in doubles) of the barrier word inside a workspace
  unsigned* epoch_host = nullptr;
  long long timeout_cycles = 0;
};

static int shard_barrier(cudaStream_t st, const ShardCtx& sh) {
  if (sh.world <= 1) return GSMVI_OK;
  ShardPtrs p;
  for (int r = 0; r < DGEMM_MAX_PEERS; ++r)
    p.cnt[r] = r < sh.world ? reinterpret_cast<unsigned*>(sh.peer_ws[r] + sh.cnt_off) : nullptr;
  *sh.epoch_host += 1;
  ns_barrier_kernel<<<1, 32, 0, st>>>(p, sh.rank, sh.world, static_cast<unsigned>(sh.world) * (*sh.epoch_host), sh.timeout_cycles);
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// rows of an n-row result owned by this rank: 128-row tiles dealt out in contiguous runs
static inline void shard_rows(const ShardCtx& sh, int n, int* row0, int* rows) {
  const int tiles = (n + 127) / 128, per = (tiles + sh.world - 1) / sh.world;
  const int r0 = sh.rank * per * 128;
  *row0 = r0 < n ? r0 : n;
  *rows = r0 < n ? ((r0 + per * 128 < n ? r0 + per * 128 : n) - r0) : 0;
}

// C = alpha * A op(B)^T + diag_add I with A K-major [n, K]: the full product on one GPU, or this rank's rows of it stored
// into every rank's copy of C (C must lie inside the solve workspace).  No barrier here: the caller places it.
static int dgemm_sharded(cudaStream_t st, const ShardCtx& sh, int n, int N, int K, const double* A, long long lda, const double* B,
                         long long ldb, bool b_mn, double* C, long long ldc, DgemmOpts o, const OzCtx& oz) {
  if (sh.world <= 1) return dgemm_big(st, n, N, K, A, lda, false, B, ldb, b_mn, C, ldc, o, oz);
  int row0, rows;
  shard_rows(sh, n, &row0, &rows);
  if (rows <= 0) return GSMVI_OK;
  o.row0 = row0;
  o.ncp = sh.world;
  for (int r = 0; r < sh.world; ++r) o.Cp[r] = sh.peer_ws[r] + (C - sh.own_ws);
  return launch_dgemm(st, rows, N, K, A + static_cast<long long>(row0) * lda, lda, false, B, ldb, b_mn, nullptr, ldc, o);
}

// ||A - d I||_F^2 and ||A||_inf into out[0], out[1], deterministically (rowbuf: 2 n doubles of scratch)
static int frob_inf64(cudaStream_t st, const double* A, long long ld, int n, double d, double* rowbuf, double* out) {
  frob_rows64_kernel<<<n, 256, 0, st>>>(A, ld, n, d, rowbuf);
  frob_final64_kernel<<<1, 256, 0, st>>>(rowbuf, n, out);
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// Coupled Newton-Schulz square root of the SPD matrix in Y (n x n, overwritten):  on return Y ~= M^{1/2}.
//   Y0 = M/c, Z0 = I;  T = Z Y;  P = a I + b T;  Y <- Y P;  Z <- P Z        
// c = ||M||_inf >= lambda_max.  Plain Newton-Schulz is (a, b) = (3/2, -1/2); while the spectrum of T is still wide
// the scaled coefficients of Chen & Chow are used: with lo a lower bound on sqrt(lambda_min(T)),
// alpha = sqrt(3 / (1 + lo + lo^2)), (a, b) = (3 alpha / 2, -alpha^3 / 2), lo <- a lo + b lo^3, which roughly halves the
// iteration count (19 instead of 38 at kappa(M) = 3.5e11).  lam_min is a lower bound on lambda_min(M).
// All three products are taken exactly as written (no symmetry shortcut such as Z Y^T): the coupled iteration is
// stable under rounding only while Y and Z keep their exact coupling - symmetrising either iterate makes it diverge
// at kappa ~ 1e11 (measured).  Needs 4 scratch n x n buffers and 2 n doubles (rowbuf).  Synchronises the
// stream once per iteration to read the residual ||I - Z Y||_F (two launches, fixed summation order: every rank of a
// sharded solve reads the same bits and stops at the same iteration).
// Sharded (sh.world > 1): each of the three products is computed by rows across the ranks, the epilogues store into every
// rank's buffers over NVLink, and two barriers per iteration separate producers from consumers.
static int ns_sqrt64(cudaStream_t st, double* Y, long long ld, int n, double* Z, double* P, double* Y2, double* Z2,
                     double* scal_dev, double* rowbuf, int max_iter, double tol, double lam_min, int* iters_out, int* flag,
                     const OzCtx& oz, const ShardCtx& sh) {
  double h[2];
  GSMVI_TRY(frob_inf64(st, Y, ld, n, 0.0, rowbuf, scal_dev));
  GSMVI_CUDA(cudaMemcpyAsync(h, scal_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GSMVI_CUDA(cudaStreamSynchronize(st));
  const double c = h[1];
  if (!(c > 0.0) || isinf(c) || isnan(c)) {
    *iters_out = -1;
    or_flag_kernel<<<1, 1, 0, st>>>(flag, 2);  // not a matrix a square root exists for: the update must be rejected
    return GSMVI_OK;
  }
  scale_diag64_kernel<<<grid2(n, n), 256, 0, st>>>(Y, ld, n, 1.0 / c, 0.0);
  set_identity64_kernel<<<grid2(n, n), 256, 0, st>>>(Z, ld, n);
  double lo = sqrt(fmin(fmax(lam_min / c, 1e-300), 1.0));
  int it = 0;
  bool last_round = false;
  double prev_res = 1e300, exit_res = 1e300;
  bool stagnated = false;
  for (; it < max_iter; ++it) {
    double a = 1.5, b = -0.5;
    if (lo < 0.9 && !last_round) {
      const double al = sqrt(3.0 / (1.0 + lo + lo * lo));
      a = 1.5 * al;
      b = -0.5 * al * al * al;
      lo = a * lo + b * lo * lo * lo;
    }
    DgemmOpts o;
    if (it == 0) {
      // Z0 = I: the products Z Y and P Z are Y and P themselves - one GEMM instead of three
      affine_diag64_kernel<<<grid2(n, n), 256, 0, st>>>(Y, P, ld, n, b, a);                      // P = a I + b Y
      GSMVI_TRY(frob_inf64(st, P, ld, n, a + b, rowbuf, scal_dev));
      GSMVI_TRY(dgemm_sharded(st, sh, n, n, n, Y, ld, P, ld, true, Y2, ld, o, oz));                // Y2 = Y P
      GSMVI_CUDA(cudaMemcpy2DAsync(Z2, ld * sizeof(double), P, ld * sizeof(double), n * sizeof(double), n,
                                   cudaMemcpyDeviceToDevice, st));                               // Z2 = P
      GSMVI_TRY(shard_barrier(st, sh));
    } else {
      DgemmOpts p;  // P = a I + b Z Y   (B operand MN-major: the product is exactly Z Y)
      p.alpha = b;
      p.diag_add = a;
      GSMVI_TRY(dgemm_sharded(st, sh, n, n, n, Z, ld, Y, ld, true, P, ld, p, oz));
      GSMVI_TRY(shard_barrier(st, sh));
      GSMVI_TRY(frob_inf64(st, P, ld, n, a + b, rowbuf, scal_dev));  // ||P - (a+b) I||_F = |b| ||I - Z Y||_F
      GSMVI_TRY(dgemm_sharded(st, sh, n, n, n, Y, ld, P, ld, true, Y2, ld, o, oz));   // Y2 = Y P
      // the last round only needs Y: its Z update is skipped
      if (!last_round) GSMVI_TRY(dgemm_sharded(st, sh, n, n, n, P, ld, Z, ld, true, Z2, ld, o, oz));   // Z2 = P Z
      GSMVI_TRY(shard_barrier(st, sh));
    }
    double* t = Y; Y = Y2; Y2 = t;
    t = Z; Z = Z2; Z2 = t;
    if (last_round) { ++it; if (!stagnated) exit_res = prev_res * prev_res; break; }  // one more quadratic step
    GSMVI_CUDA(cudaMemcpyAsync(h, scal_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    GSMVI_CUDA(cudaStreamSynchronize(st));
    const double res = sqrt(h[0]) / fabs(b);
    if (dbg_on()) fprintf(stderr, "[gsmvi debug]   NS it=%d a=%.4f b=%.4f res=%.6e\n", it, a, b, res);
    exit_res = res;
    if (isnan(res) || isinf(res)) { ++it; break; }
    if (res < tol) { ++it; break; }            // this update already used P with ||I - ZY|| < tol: converged
    if (res < 1e-4 && a == 1.5) last_round = true;   // quadratic convergence: one more plain update reaches ~1e-9
    if (res < 1e-2 && res > 0.5 * prev_res) last_round = stagnated = true;  // stagnating at the rounding floor
    prev_res = res;
  }
  *iters_out = it;
  // Out of iterations, or a non-finite residual: Y is not M^(1/2), and S = 2 T T^T built from it would still pass the PD
  // check (it is PSD by construction) - so the solve itself raises bit 1 of the flag and the update is rejected
  // (the reference's sqrtm would have thrown, which BaM.fit's retry loop catches, bam.py:188-206).
  if (!(exit_res < (stagnated ? 1e-4 : 1e-6))) or_flag_kernel<<<1, 1, 0, st>>>(flag, 2);
  // result lives in the current Y; if that is the caller's Y2 buffer, copy back (odd number of swaps)
  if (it % 2 == 1) GSMVI_CUDA(cudaMemcpy2DAsync(Y2, ld * sizeof(double), Y, ld * sizeof(double), n * sizeof(double), n,
                                                cudaMemcpyDeviceToDevice, st));
  scale_diag64_kernel<<<grid2(n, n), 256, 0, st>>>((it % 2 == 1) ? Y2 : Y, ld, n, sqrt(c), 0.0);
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// ------------------------------------------------------------------------------------------------ public pieces

int potrf64(cudaStream_t st, double* A, long long lda, int n, int* flag) {
  if (!A || !flag || n <= 0 || lda < n) return GSMVI_EINVAL;
  return potrf64_inplace(st, A, lda, n, flag);
}

static inline long long rup(long long v, long long m) { return (v + m - 1) / m * m; }

size_t bam_stats_workspace_bytes(int B, int D) {
  const long long ld = rup(D, 8);
  // fp64: T = [Xc; Gc] [2B x ld] + C [D x ld] + xbar, gbar [2 ld]
  return static_cast<size_t>((2LL * B + D + 2) * ld) * sizeof(double);
}

// Layout of the full solve's workspace (doubles): 6 D x ld buffers | Q, W [D x ldk] | two sets of diagonal-block inverses |
// rowbuf [2 max(D, K)] | barrier word (8 doubles) | scalars (16) | (1 KiB aligned) int8 digit planes of the optional
// tensor-core products.  bam_solve_cnt_offset() is where a sharded solve's barrier word lives (same on every rank).
static inline long long bam_full_fixed_doubles(int B, int D) {
  const long long ld = rup(D, 8);
  const long long nblk = (D + NB64 - 1) / NB64;
  const long long K = B + 1, ldk = rup(K, 8);
  return 6 * D * ld + 2 * D * ldk + 2 * nblk * NB64 * NB64;
}
long long bam_solve_cnt_offset(int B, int D) {
  const long long K = B + 1;
  return bam_full_fixed_doubles(B, D) + 2 * rup(D > K ? D : K, 8);
}

size_t bam_solve_workspace_bytes(int B, int D, int lowrank) {
  const long long ld = rup(D, 8);
  const long long K = B + 1, ldk = rup(K, 8), kblk = (K + NB64 - 1) / NB64;
  if (!lowrank) {
    size_t b = static_cast<size_t>(bam_solve_cnt_offset(B, D) + 8 + 16) * sizeof(double);
    if (D >= 1024) b = static_cast<size_t>(rup(static_cast<long long>(b), 1024)) + 1024 + oz_workspace_bytes(D, D, D > K ? D : K, 8);
    return b;
  }
  // V [D x ld], S [D x ld], Q, A = VQ, W = A F [D x ldk], 6 K x K buffers, inverses, rowbuf [2 K], scalars
  return static_cast<size_t>(2 * D * ld + 3 * D * ldk + 6 * K * ldk + kblk * NB64 * NB64 + 2 * ldk + 16) * sizeof(double);
}

int bam_stats(cudaStream_t st, const float* X, long long ldx, const float* G, long long ldg, int B, int D, int Btot,
              double* ws, int stage) {
  // fp64 statistics from the fp32 samples / scores.  V = S0 + reg*C multiplies C's rounding by reg (~100 early in the
  // paper's schedule), so C, the means and the centred rows are kept in fp64; X and G stay fp32.
  // stage 0: column sums into xbar/gbar (unnormalised; all-reduce them across shards before stage 1)
  // stage 1: xbar,gbar /= Btot; T = [X - xbar; G - gbar]; C = Xc^T Xc / Btot (partial over this shard)
  const long long ld = rup(D, 8);
  double* T = ws;
  double* C = T + 2LL * B * ld;
  double* xbar = C + static_cast<long long>(D) * ld;
  double* gbar = xbar + ld;
  if (stage == 0) {
    GSMVI_CUDA(cudaMemsetAsync(xbar, 0, 2 * ld * sizeof(double), st));
    colsum2_kernel<<<dim3((D + 255) / 256, (B + CS_ROWS - 1) / CS_ROWS), 256, 0, st>>>(X, ldx, G, ldg, B, D, xbar, gbar);
    GSMVI_CUDA(last());
    return GSMVI_OK;
  }
  scale_vec2_kernel<<<(D + 255) / 256, 256, 0, st>>>(xbar, gbar, 1.0 / static_cast<double>(Btot), D);
  center2_kernel<<<grid2(D, B), 256, 0, st>>>(X, ldx, G, ldg, B, D, xbar, gbar, T, ld);
  DgemmOpts o;
  o.alpha = 1.0 / static_cast<double>(Btot);
  o.tri = true;
  o.mirror = true;
  GSMVI_TRY(launch_dgemm(st, D, D, B, T, ld, true, T, ld, true, C, ld, o));  // C = Xc^T Xc / Btot (rows of T are K)
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

// GSMVI_BAM_TIMING=1: CUDA-event stamps between the stages of the solve, printed per call (synchronises; diagnostics only)
struct SolveTimer {
  bool on;
  cudaStream_t st;
  cudaEvent_t ev[12];
  const char* name[12];
  int n = 0;
  explicit SolveTimer(cudaStream_t s) : st(s) {
    static int v = -1;
    if (v < 0) v = getenv("GSMVI_BAM_TIMING") ? 1 : 0;
    on = v == 1;
  }
  void mark(const char* what) {
    if (!on || n >= 12) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void report(int rank, int iters) {
    if (!on) return;
    cudaStreamSynchronize(st);
    if (rank == 0) {
      fprintf(stderr, "[gsmvi bam solve timing, ms] ns_iters %d |", iters);
      for (int i = 1; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        fprintf(stderr, " %s %.3f", name[i], ms);
      }
      fprintf(stderr, "\n");
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

int bam_solve_full(cudaStream_t st, const double* stats_ws, int B, int D, int Btot, const float* mu0, const float* S0,
                   long long lds0, double reg, double jitter, float* mu_out, float* S_out, long long ldso, double* ws,
                   int max_ns, int* ns_iters_host, int* flag, int world, int phase, const BamShard* shard) {
  // phase 0: everything.  Sharded batch (world > 1): phase 1 stops after this rank's partial
  // M_r = I/world + 4 W_r W_r^T (D x ld doubles at ws + 3 D ld: sum it over ranks), phase 2 resumes at the square root.
  // With `shard` (peer-mapped workspaces) phase 2 is tensor-parallel: Newton-Schulz products, T = L R^-T and S = 2 T T^T
  // are computed by rows across the ranks; without it every rank repeats the whole of phase 2.
  const long long ld = rup(D, 8);
  const int K = B + 1;
  const long long ldk = rup(K, 8);
  const double* Tc = stats_ws;  // [Xc; Gc]
  const double* C = stats_ws + 2LL * B * ld;
  const double* xbar = C + static_cast<long long>(D) * ld;
  const double* gbar = xbar + ld;
  double* b0 = ws;                  // Z2
  double* b1 = b0 + D * ld;         // V -> L (kept)
  double* b2 = b1 + D * ld;         // Z
  double* b3 = b2 + D * ld;         // M      -> Y -> N -> R
  double* b4 = b3 + D * ld;         // P      -> T = L R^{-T}
  double* b5 = b4 + D * ld;         // Y2     -> T T^T
  double* Q = b5 + D * ld;          // [D x ldk]
  double* W = Q + D * ldk;          // [D x ldk]  L^T Q
  const long long nblk = (D + NB64 - 1) / NB64;
  double* rowbuf = W + D * ldk + 2 * nblk * NB64 * NB64;  // (the 2 nblk 64 x 64 blocks before it are reserved, unused)
  double* cntw = ws + bam_solve_cnt_offset(B, D);
  double* scal = cntw + 8;
  OzCtx oz;
  if (D >= 1024 && oz_slices_env() >= 2) {
    oz.ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(scal + 16) + 1023) & ~static_cast<uintptr_t>(1023));
    oz.slices = oz_slices_env();
  }
  ShardCtx sh;
  if (shard && world > 1 && phase == 2) {
    if (shard->world != world || world > DGEMM_MAX_PEERS || !shard->epoch_host) return GSMVI_EINVAL;
    sh.rank = shard->rank;
    sh.world = world;
    sh.own_ws = ws;
    sh.cnt_off = bam_solve_cnt_offset(B, D);
    sh.epoch_host = shard->epoch_host;
    for (int r = 0; r < world; ++r) sh.peer_ws[r] = static_cast<double*>(shard->peer_ws[r]);
    if (sh.peer_ws[sh.rank] != ws) return GSMVI_EINVAL;
    static long long tmo = -1;
    if (tmo < 0) {
      const char* e = getenv("GSMVI_COMM_TIMEOUT_S");
      const double sec = e ? atof(e) : 300.0;
      tmo = sec <= 0.0 ? 0 : static_cast<long long>(sec * 2.0e9);
    }
    sh.timeout_cycles = tmo;
    oz = OzCtx();  // the int8 path has no sharded form
  }
  SolveTimer tmr(st);
  tmr.mark("start");
  if (phase != 2) {
  GSMVI_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  bam_v_kernel<<<grid2(D, D), 256, 0, st>>>(C, ld, S0, lds0, xbar, mu0, reg, b1, ld, D);
  dbg_stage(st, "V", b1, ld, D, scal, flag);
  GSMVI_TRY(potrf64_inplace(st, b1, ld, D, flag));                          // V = L L^T
  dbg_stage(st, "L=chol(V)", b1, ld, D, scal, flag);
  tmr.mark("V+chol(V)");
  bam_build_q_kernel<<<dim3((K + 255) / 256, D), 256, 0, st>>>(Tc + static_cast<long long>(B) * ld, ld, gbar, B, D, reg,
                                                               Btot, world, Q, ldk);  // sum over ranks of Q Q^T = U
  DgemmOpts o;
  o.krange = KR_A_UPPER;  // A operand is L^T given as MN-major L: (L^T)[i][k] = L[k][i] == 0 for k < i
  GSMVI_TRY(dgemm_big(st, D, K, D, b1, ld, true, Q, ldk, true, W, ldk, o, oz));     // W = L^T Q
  DgemmOpts m;
  m.alpha = 4.0;
  m.diag_add = 1.0 / world;
  m.tri = true;
  m.mirror = true;
  GSMVI_TRY(dgemm_big(st, D, D, K, W, ldk, false, W, ldk, false, b3, ld, m, oz));   // M = I + 4 W W^T  (= I + 4 L^T U L)
  dbg_stage(st, "M", b3, ld, D, scal, flag);
  tmr.mark("W,M");
  }
  if (phase == 1) {
    tmr.report(shard ? shard->rank : 0, 0);
    GSMVI_CUDA(last());
    return GSMVI_OK;
  }
  int iters = 0;
  GSMVI_TRY(ns_sqrt64(st, b3, ld, D, b2, b4, b5, b0, scal, rowbuf, max_ns, 1e-11, 1.0, &iters, flag, oz, sh));  // b3 = N = M^{1/2}
  if (ns_iters_host) *ns_iters_host = iters;
  dbg_stage(st, "N=sqrt(M)", b3, ld, D, scal, flag);
  tmr.mark("newton-schulz");
  scale_diag64_kernel<<<grid2(D, D), 256, 0, st>>>(b3, ld, D, 1.0, 1.0);          // I + N
  GSMVI_TRY(potrf64_inplace(st, b3, ld, D, flag));                          // I + N = R R^T
  dbg_stage(st, "R=chol(I+N)", b3, ld, D, scal, flag);
  tmr.mark("chol(I+N)");
  if (sh.world > 1) {
    // rows of T = L R^{-T} are independent: this rank solves its rows, sends them to every rank, then computes its rows of
    // S / 2 = T T^T with the all-gather fused into the GEMM epilogue
    int row0, rows;
    shard_rows(sh, D, &row0, &rows);
    if (rows > 0) {
      const long long off = static_cast<long long>(row0) * ld;
      GSMVI_TRY(trsm64_right_lt_chunked(st, b1 + off, ld, b3, ld, b4 + off, ld, rows, D));
      BcastArgs ba;
      ba.src = b4 + off;
      ba.ndst = 0;
      for (int r = 0; r < sh.world; ++r)
        if (r != sh.rank) ba.dst[ba.ndst++] = sh.peer_ws[r] + (b4 - ws) + off;
      ba.count2 = static_cast<long long>(rows) * ld / 2;
      peer_bcast_kernel<<<296, 256, 0, st>>>(ba);
    }
    GSMVI_TRY(shard_barrier(st, sh));
    tmr.mark("trsm+bcast");
    DgemmOpts s;
    GSMVI_TRY(dgemm_sharded(st, sh, D, D, D, b4, ld, b4, ld, false, b5, ld, s, oz));  // rows of T T^T
    GSMVI_TRY(shard_barrier(st, sh));
    tmr.mark("T T^T");
  } else {
    GSMVI_TRY(trsm64_right_lt_chunked(st, b1, ld, b3, ld, b4, ld, D, D));     // T = L R^{-T}
    dbg_stage(st, "T=L R^-T", b4, ld, D, scal, flag);
    tmr.mark("trsm");
    DgemmOpts s;
    s.tri = true;
    s.mirror = true;
    GSMVI_TRY(dgemm_big(st, D, D, D, b4, ld, false, b4, ld, false, b5, ld, s, oz));   // b5 = T T^T ; S = 2 b5
    tmr.mark("T T^T");
  }
  bam_mean_kernel<<<(D + 7) / 8, 256, 0, st>>>(b5, ld, 2.0, gbar, xbar, mu0, reg, mu_out, D);
  bam_finish_cov_kernel<<<grid2(D, D), 256, 0, st>>>(b5, ld, S_out, ldso, D, 2.0, jitter);
  tmr.mark("mean+finish");
  tmr.report(sh.rank, iters);
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

int bam_solve_lowrank(cudaStream_t st, const double* stats_ws, int B, int D, int Btot, const float* mu0, const float* S0,
                      long long lds0, double reg, double jitter, float* mu_out, float* S_out, long long ldso, double* ws,
                      int max_ns, int* ns_iters_host, int* flag) {
  // bam.py:102-112 with Q Q^T = U exactly:  A = V Q;  H = Q^T V Q + I/4;  BB = (I/2 + H^{1/2})^2;  S = V - A BB^{-1} A^T.
  // BB^{-1} = F^2 with F = (I/2 + H^{1/2})^{-1} = T1 T1^T, T1 = R2^{-T}, (I/2 + H^{1/2}) = R2 R2^T; S = V - (A F)(A F)^T.
  const long long ld = rup(D, 8);
  const int K = B + 1;
  const long long ldk = rup(K, 8), kblk = (K + NB64 - 1) / NB64;
  const double* Tc = stats_ws;  // [Xc; Gc]
  const double* C = stats_ws + 2LL * B * ld;
  const double* xbar = C + static_cast<long long>(D) * ld;
  const double* gbar = xbar + ld;
  double* V = ws;
  double* Sb = V + D * ld;
  double* Q = Sb + D * ld;
  double* A = Q + D * ldk;
  double* W = A + D * ldk;
  double* k0 = W + D * ldk;  // H -> Y
  double* k1 = k0 + K * ldk;
  double* k2 = k1 + K * ldk;
  double* k3 = k2 + K * ldk;
  double* k4 = k3 + K * ldk;
  double* k5 = k4 + K * ldk;
  double* rowbuf = k5 + K * ldk + kblk * NB64 * NB64;  // (kblk 64 x 64 blocks before it reserved, unused)
  double* scal = rowbuf + 2 * ldk;
  GSMVI_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  bam_v_kernel<<<grid2(D, D), 256, 0, st>>>(C, ld, S0, lds0, xbar, mu0, reg, V, ld, D);
  bam_build_q_kernel<<<dim3((K + 255) / 256, D), 256, 0, st>>>(Tc + static_cast<long long>(B) * ld, ld, gbar, B, D, reg,
                                                               Btot, 1, Q, ldk);
  DgemmOpts o;
  GSMVI_TRY(launch_dgemm(st, D, K, D, V, ld, false, Q, ldk, true, A, ldk, o));     // A = V Q   (V symmetric)
  DgemmOpts h;
  h.diag_add = 0.25;
  h.tri = true;
  h.mirror = true;
  GSMVI_TRY(launch_dgemm(st, K, K, D, A, ldk, true, Q, ldk, true, k0, ldk, h));    // H = A^T Q + I/4
  int iters = 0;
  GSMVI_TRY(ns_sqrt64(st, k0, ldk, K, k1, k2, k3, k4, scal, rowbuf, max_ns, 1e-11, 0.25, &iters, flag, OzCtx(), ShardCtx()));
  if (ns_iters_host) *ns_iters_host = iters;
  scale_diag64_kernel<<<grid2(K, K), 256, 0, st>>>(k0, ldk, K, 1.0, 0.5);          // I/2 + H^{1/2}
  GSMVI_TRY(potrf64_inplace(st, k0, ldk, K, flag));                          // = R2 R2^T
  set_identity64_kernel<<<grid2(K, K), 256, 0, st>>>(k1, ldk, K);
  GSMVI_TRY(trsm64_right_lt(st, k1, ldk, k0, ldk, k1, ldk, K, K));           // T1 = R2^{-T}
  DgemmOpts f;
  f.tri = true;
  f.mirror = true;
  GSMVI_TRY(launch_dgemm(st, K, K, K, k1, ldk, false, k1, ldk, false, k2, ldk, f));  // F = T1 T1^T
  GSMVI_TRY(launch_dgemm(st, D, K, K, A, ldk, false, k2, ldk, false, W, ldk, o));    // W = A F  (F symmetric)
  DgemmOpts s;
  s.alpha = -1.0;
  s.beta = 1.0;
  s.Cin = V;
  s.ldcin = ld;
  s.tri = true;
  s.mirror = true;
  GSMVI_TRY(launch_dgemm(st, D, D, K, W, ldk, false, W, ldk, false, Sb, ld, s));   // S = V - W W^T
  bam_mean_kernel<<<(D + 7) / 8, 256, 0, st>>>(Sb, ld, 1.0, gbar, xbar, mu0, reg, mu_out, D);
  bam_finish_cov_kernel<<<grid2(D, D), 256, 0, st>>>(Sb, ld, S_out, ldso, D, 1.0, jitter);
  GSMVI_CUDA(last());
  return GSMVI_OK;
}

}  // namespace gsmvi
