// FP64 GEMM for the BaM quadratic-matrix-equation solve:  C = alpha * op(A) op(B)^T + beta * Cin + diag_add * I.
//
// Why fp64: the solve S = 2 L (I + (I + 4 L^T U L)^{1/2})^{-1} L^T (gsmvi/bam.py:59-65 symmetrised) mixes scales of
// 1 and ~1e9; in fp32 the "I +" is rounded away and the update is garbage or NaN (SURVEY.md section 7 hard part 1,
// section 9 table C; the reference itself runs BaM under jax_enable_x64, examples/example_bam.py:14-15).  tcgen05 has
// no f64 kind, and on B200 the FP64 CUDA-core and DMMA rates are the same, so this is a register-tiled DFMA kernel:
// 128x128 CTA tile, BK = 16, 256 threads each owning an 8x8 micro-tile, double-buffered through registers.  A thread's
// rows are the pairs {2ty, 2ty+1} + 32a and its columns {2tx, 2tx+1} + 32b (a, b = 0..3), so the 16 lanes that differ
// in tx read 16 consecutive 16-byte words per shared-memory load (conflict-free) and lanes that share ty broadcast:
// 12 wavefronts per 64 DFMA (the first version, 8x4 tiles with 32-byte-strided reads, needed 0.85 per DFMA and ran the
// FP64 pipe at 41%, profiles/r01_ncu_dgemm_summary.txt).
// Operand conventions match tc_gemm.cuh: K-major operand = [rows, K] row-major, MN-major = [K, rows] row-major.
#include "dgemm.cuh"

namespace gsmvi {

constexpr int DBK = 16, DTHREADS = 256;

struct DgemmArgs {
  int M, N, K;
  double alpha, beta, diag_add;
  const double* A;
  long long lda;
  const double* B;
  long long ldb;
  const double* Cin;
  long long ldcin;
  double* C;
  long long ldc;
  int tri, mirror, krange;
};

// Load a ROWS x DBK tile (rows r0.., k from k0) into registers: each thread takes ROWS*DBK/256 elements.
template <int ROWS, bool MN>
__device__ __forceinline__ void dload(const double* __restrict__ P, long long ld, int nrows, int K, int r0, int k0,
                                      double (&reg)[ROWS * DBK / DTHREADS]) {
  constexpr int PER = ROWS * DBK / DTHREADS;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int idx = threadIdx.x + e * DTHREADS;
    int r, k;
    if (MN) {  // consecutive threads walk rows (contiguous in memory)
      r = idx % ROWS;
      k = idx / ROWS;
    } else {   // consecutive threads walk k (contiguous in memory)
      k = idx % DBK;
      r = idx / DBK;
    }
    const int gr = r0 + r, gk = k0 + k;
    double v = 0.0;
    if (gr < nrows && gk < K) v = MN ? P[static_cast<long long>(gk) * ld + gr] : P[static_cast<long long>(gr) * ld + gk];
    reg[e] = v;
  }
}

template <int ROWS, bool MN>
__device__ __forceinline__ void dstore(double* __restrict__ S, const double (&reg)[ROWS * DBK / DTHREADS]) {
  constexpr int PER = ROWS * DBK / DTHREADS;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int idx = threadIdx.x + e * DTHREADS;
    int r, k;
    if (MN) {
      r = idx % ROWS;
      k = idx / ROWS;
    } else {
      k = idx % DBK;
      r = idx / DBK;
    }
    S[k * (ROWS + 2) + r] = reg[e];  // smem layout [k][row], +2 padding
  }
}

// DBM = DBN = 128 (8x8 per thread) for large problems, 64 (4x4 per thread, 2 CTAs/SM) when 128-tiles would not fill the GPU.
template <int DBM, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(DTHREADS, (DBM == 128) ? 1 : 2) dgemm_kernel(const DgemmArgs a) {
  constexpr int DBN = DBM, TM = DBM / 16;
  __shared__ __align__(16) double As[DBK * (DBM + 2)];
  __shared__ __align__(16) double Bs[DBK * (DBN + 2)];
  int tm, tn;
  {
    const int tiles_n = (a.N + DBN - 1) / DBN;
    tm = blockIdx.x / tiles_n;
    tn = blockIdx.x % tiles_n;
    if (a.tri && tn > tm) return;  // lower tiles only
  }
  const int m0 = tm * DBM, n0 = tn * DBN;
  int k_begin = 0, k_end = a.K;
  if (a.krange & KR_A_LOWER) k_end = min(k_end, m0 + DBM);
  if (a.krange & KR_B_LOWER) k_end = min(k_end, n0 + DBN);
  if (a.krange & KR_A_UPPER) k_begin = max(k_begin, m0);
  if (a.krange & KR_B_UPPER) k_begin = max(k_begin, n0);
  k_begin = (k_begin / DBK) * DBK;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  double acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.0;

  double ra[DBM * DBK / DTHREADS], rb[DBN * DBK / DTHREADS];
  if (k_begin < k_end) {
    dload<DBM, A_MN>(a.A, a.lda, a.M, a.K, m0, k_begin, ra);
    dload<DBN, B_MN>(a.B, a.ldb, a.N, a.K, n0, k_begin, rb);
  }
  for (int k0 = k_begin; k0 < k_end; k0 += DBK) {
    __syncthreads();
    dstore<DBM, A_MN>(As, ra);
    dstore<DBN, B_MN>(Bs, rb);
    __syncthreads();
    if (k0 + DBK < k_end) {
      dload<DBM, A_MN>(a.A, a.lda, a.M, a.K, m0, k0 + DBK, ra);
      dload<DBN, B_MN>(a.B, a.ldb, a.N, a.K, n0, k0 + DBK, rb);
    }
#pragma unroll
    for (int k = 0; k < DBK; ++k) {
      double av[TM], bv[TM];
      const double* ap = As + k * (DBM + 2) + 2 * ty;
      const double* bp = Bs + k * (DBN + 2) + 2 * tx;
#pragma unroll
      for (int q = 0; q < TM / 2; ++q) {
        const double2 ta = *reinterpret_cast<const double2*>(ap + 32 * q);
        av[2 * q] = ta.x;
        av[2 * q + 1] = ta.y;
        const double2 tb = *reinterpret_cast<const double2*>(bp + 32 * q);
        bv[2 * q] = tb.x;
        bv[2 * q + 1] = tb.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
  }
  // epilogue: element (i, j) of the micro-tile is row m0 + 32 (i/2) + 2 ty + (i&1), column n0 + 32 (j/2) + 2 tx + (j&1)
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + 32 * (i >> 1) + 2 * ty + (i & 1);
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int n = n0 + 32 * (j >> 1) + 2 * tx + (j & 1);
      if (n >= a.N) continue;
      if (a.tri && n > m) continue;
      double v = a.alpha * acc[i][j];
      if (a.beta != 0.0) v += a.beta * a.Cin[static_cast<long long>(m) * a.ldcin + n];
      if (m == n) v += a.diag_add;
      a.C[static_cast<long long>(m) * a.ldc + n] = v;
      if (a.mirror && n != m) a.C[static_cast<long long>(n) * a.ldc + m] = v;
    }
  }
}

int launch_dgemm(cudaStream_t stream, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                 long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o) {
  if (M <= 0 || N <= 0 || K < 0 || !C || (K > 0 && (!A || !B))) return GSMVI_EINVAL;
  if (o.beta != 0.0 && !o.Cin) return GSMVI_EINVAL;
  DgemmArgs a;
  a.M = M; a.N = N; a.K = K;
  a.alpha = o.alpha; a.beta = o.beta; a.diag_add = o.diag_add;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb;
  a.Cin = o.Cin; a.ldcin = o.ldcin; a.C = C; a.ldc = ldc;
  a.tri = o.tri ? 1 : 0; a.mirror = o.mirror ? 1 : 0; a.krange = o.krange;
  const long long tiles128 = static_cast<long long>((M + 127) / 128) * ((N + 127) / 128);
  const bool big = (o.tri ? tiles128 / 2 : tiles128) >= 120;  // enough 128x128 tiles to fill 148 SMs
#define GSMVI_DG(BMV)                                                                            \
  {                                                                                              \
    const int grid = ((M + BMV - 1) / BMV) * ((N + BMV - 1) / BMV);                              \
    if (!a_mn && !b_mn) dgemm_kernel<BMV, false, false><<<grid, DTHREADS, 0, stream>>>(a);        \
    else if (a_mn && !b_mn) dgemm_kernel<BMV, true, false><<<grid, DTHREADS, 0, stream>>>(a);     \
    else if (!a_mn && b_mn) dgemm_kernel<BMV, false, true><<<grid, DTHREADS, 0, stream>>>(a);     \
    else dgemm_kernel<BMV, true, true><<<grid, DTHREADS, 0, stream>>>(a);                         \
  }
  if (big) GSMVI_DG(128) else GSMVI_DG(64)
#undef GSMVI_DG
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
