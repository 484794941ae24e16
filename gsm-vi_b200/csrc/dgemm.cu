// FP64 GEMM for the BaM quadratic-matrix-equation solve:  C = alpha * op(A) op(B)^T + beta * Cin + diag_add * I.
//
// Why fp64: the solve S = 2 L (I + (I + 4 L^T U L)^{1/2})^{-1} L^T (gsmvi/bam.py:59-65 symmetrised) mixes scales of
// 1 and ~1e9; in fp32 the "I +" is rounded away and the update is garbage or NaN (SURVEY.md section 7 hard part 1,
// section 9 table C; the reference itself runs BaM under jax_enable_x64, examples/example_bam.py:14-15).  tcgen05 has
// no f64 kind, so the products run on the FP64 tensor-core path (mma.sync.m16n8k16.f64): 128 x 128 CTA tile, eight warps
// of 64 x 32, BK = 16.  Two kernels: dgemm_pipe_kernel moves the operand tiles with cp.async through a three-stage ring
// (30.9 TFLOP/s at 4096^3, cuBLAS DGEMM on the same part 35.3; the default for every large aligned product);
// dgemm_mma16_kernel stages through registers (any alignment, 64 x 64 tiles for small problems).  The first two
// generations (DFMA register tiles, m8n8k4) were removed after round 1.
// Row-sharded mode (tensor-parallel Newton-Schulz across GPUs): the epilogue stores every result element into this GPU's
// and its NVLink peers' copies of the result matrix, so the all-gather of a sharded product rides on the GEMM itself.
// Operand conventions match tc_gemm.cuh: K-major operand = [rows, K] row-major, MN-major = [K, rows] row-major.
#include "dev_once.cuh"
#include "dgemm.cuh"

#include <stdint.h>
#include <stdlib.h>

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
  // row-sharded products (tensor-parallel Newton-Schulz, bam_solve.cu): the M rows computed here are rows row0 .. row0+M-1
  // of the full result (diag_add, tri, Cin and the stores use the global row), and every finished element is stored into
  // the ncp full-size result matrices Cp[] - this GPU's and its NVLink peers' (the all-gather rides on the epilogue)
  int row0, ncp;
  double* Cp[DGEMM_MAX_PEERS];
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

// ------------------------------------------------------------------------------------------------ DMMA variant
// Same tiling and register-staged double buffering, but the inner product runs on the FP64 tensor-core path
// (mma.sync m8n8k4): a warp owns a (DBM/2) x (DBN/4) sub-tile as 8x8 accumulator blocks, operand fragments are one
// 8-byte shared-memory load per thread and feed 4 (A) / 8 (B) MMAs each, so the kernel needs 12 shared-memory wavefronts
// per 32 MMAs where the DFMA version needs 12 per 64 FMAs and is co-limited by shared-memory bandwidth.
// smem layout [k][row] with row stride DBM + 4 doubles: the 4 (k) x 8 (row) fragment touches 32 distinct 8-byte slots.
// m16n8k16: A fragment a_i at row g + 8 (i & 1), k = t + 4 (i >> 1); B fragment b_i at k = t + 4 i, column g;
// C fragment c0, c1 at row g, columns 2t, 2t+1 and c2, c3 at row g + 8 (g = lane / 4, t = lane % 4).
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, "
      "{%12, %13, %14, %15}, {%0, %1, %2, %3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

template <int ROWS, bool MN>
__device__ __forceinline__ void dstore4(double* __restrict__ S, const double (&reg)[ROWS * DBK / DTHREADS]) {
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
    if (MN) S[k * (ROWS + 4) + r] = reg[e];  // rows arrive contiguous: [k][row], conflict-free stores
    else S[r * (DBK + 4) + k] = reg[e];       // k arrives contiguous: [row][k] with stride 20 keeps stores AND fragments conflict-free
  }
}
// element (row r, k) of a staged operand tile
template <int ROWS, bool MN>
__device__ __forceinline__ int sidx(int r, int k) {
  return MN ? k * (ROWS + 4) + r : r * (DBK + 4) + k;
}

// Epilogue shared by the kernels below: fragment layout of m16n8k16 (c0, c1 at row g, columns 2t, 2t+1; c2, c3 at row g+8).
// Cin is read in batches BEFORE any store of the batch: Cin may alias C (in-place rank-k updates of the Cholesky / TRSM), so
// the compiler cannot hoist a load above an earlier store by itself, and 64 dependent load -> store round trips per thread
// made a K = 64 update cost ~60 us where its arithmetic needs 3.
template <int BI, int BJ>
__device__ __forceinline__ void dgemm_epilogue(const DgemmArgs& a, const double (&acc)[BI][BJ][4], int m0, int n0, int wr, int wc,
                                               int fg, int ft) {
  const bool vec_cin = a.beta != 0.0 && (a.ldcin & 1) == 0 && (reinterpret_cast<uintptr_t>(a.Cin) & 15) == 0;
#pragma unroll
  for (int i = 0; i < BI; ++i) {
    double cin[2][BJ][2];
    if (a.beta != 0.0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wr + 16 * i + fg + 8 * h;
        const long long gm = a.row0 + m;
#pragma unroll
        for (int j = 0; j < BJ; ++j) {
          const int n = n0 + wc + 8 * j + 2 * ft;
          cin[h][j][0] = cin[h][j][1] = 0.0;
          if (m < a.M && n < a.N) {
            const double* src = a.Cin + gm * a.ldcin + n;
            if (vec_cin && n + 1 < a.N) {
              const double2 t = *reinterpret_cast<const double2*>(src);
              cin[h][j][0] = t.x;
              cin[h][j][1] = t.y;
            } else {
              cin[h][j][0] = src[0];
              if (n + 1 < a.N) cin[h][j][1] = src[1];
            }
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + wr + 16 * i + fg + 8 * h;
      if (m >= a.M) continue;
      const long long gm = a.row0 + m;  // row of the full result
#pragma unroll
      for (int j = 0; j < BJ; ++j) {
        const int n = n0 + wc + 8 * j + 2 * ft;
        double v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          v[u] = a.alpha * acc[i][j][2 * h + u];
          if (a.beta != 0.0) v[u] += a.beta * cin[h][j][u];
          if (gm == n + u) v[u] += a.diag_add;
        }
        if (a.ncp > 0) {
          // every destination gets the pair as one 16-byte store where alignment allows (peer stores cross NVLink)
          const bool pair = (n + 1 < a.N) && ((a.ldc & 1) == 0);
          for (int d = 0; d < a.ncp; ++d) {
            double* dst = a.Cp[d] + gm * a.ldc + n;
            if (pair && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
            } else {
              if (n < a.N) dst[0] = v[0];
              if (n + 1 < a.N) dst[1] = v[1];
            }
          }
          continue;
        }
        double* dst = a.C + gm * a.ldc + n;
        if (!a.tri && !a.mirror && n + 1 < a.N && (a.ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
          continue;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (n + u >= a.N) continue;
          if (a.tri && n + u > gm) continue;
          dst[u] = v[u];
          if (a.mirror && n + u != gm) a.C[static_cast<long long>(n + u) * a.ldc + gm] = v[u];
        }
      }
    }
  }
  // peer stores must be performed system-wide before this grid counts as finished: the barrier kernel that follows
  // signals the peers with a release that is only cumulative over what this GPU has already made visible
  if (a.ncp > 1) __threadfence_system();
}

template <int DBM, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(DTHREADS, (DBM == 128) ? 1 : 2) dgemm_mma16_kernel(const DgemmArgs a) {
  constexpr int DBN = DBM;
  constexpr int WM = DBM / 2, WN = DBN / 4;   // warp sub-tile (8 warps as 2 x 4)
  constexpr int BI = WM / 16, BJ = WN / 8;    // 16x8 blocks per warp
  static_assert(DBK == 16, "one m16n8k16 step per staged k-block");
  __shared__ __align__(16) double As[DBM * (DBK + 4)];  // covers both [row][k+4] and [k][row+4]
  __shared__ __align__(16) double Bs[DBN * (DBK + 4)];
  int tm, tn;
  {
    const int tiles_n = (a.N + DBN - 1) / DBN;
    tm = blockIdx.x / tiles_n;
    tn = blockIdx.x % tiles_n;
    if (a.tri && tn > tm) return;
  }
  const int m0 = tm * DBM, n0 = tn * DBN;
  int k_begin = 0, k_end = a.K;
  if (a.krange & KR_A_LOWER) k_end = min(k_end, m0 + DBM);
  if (a.krange & KR_B_LOWER) k_end = min(k_end, n0 + DBN);
  if (a.krange & KR_A_UPPER) k_begin = max(k_begin, m0);
  if (a.krange & KR_B_UPPER) k_begin = max(k_begin, n0);
  k_begin = (k_begin / DBK) * DBK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 2) * WM, wc = (warp & 3) * WN;
  const int fg = lane >> 2, ft = lane & 3;
  double acc[BI][BJ][4];
#pragma unroll
  for (int i = 0; i < BI; ++i)
#pragma unroll
    for (int j = 0; j < BJ; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[i][j][u] = 0.0;
  double ra[DBM * DBK / DTHREADS], rb[DBN * DBK / DTHREADS];
  if (k_begin < k_end) {
    dload<DBM, A_MN>(a.A, a.lda, a.M, a.K, m0, k_begin, ra);
    dload<DBN, B_MN>(a.B, a.ldb, a.N, a.K, n0, k_begin, rb);
  }
  for (int k0 = k_begin; k0 < k_end; k0 += DBK) {
    __syncthreads();
    dstore4<DBM, A_MN>(As, ra);
    dstore4<DBN, B_MN>(Bs, rb);
    __syncthreads();
    if (k0 + DBK < k_end) {
      dload<DBM, A_MN>(a.A, a.lda, a.M, a.K, m0, k0 + DBK, ra);
      dload<DBN, B_MN>(a.B, a.ldb, a.N, a.K, n0, k0 + DBK, rb);
    }
    double af[BI][8], bf[BJ][4];
#pragma unroll
    for (int i = 0; i < BI; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) af[i][e] = As[sidx<DBM, A_MN>(wr + 16 * i + fg + 8 * (e & 1), ft + 4 * (e >> 1))];
#pragma unroll
    for (int j = 0; j < BJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) bf[j][e] = Bs[sidx<DBN, B_MN>(wc + 8 * j + fg, ft + 4 * e)];
#pragma unroll
    for (int i = 0; i < BI; ++i)
#pragma unroll
      for (int j = 0; j < BJ; ++j) dmma16816(acc[i][j], af[i], bf[j]);
  }
  dgemm_epilogue<BI, BJ>(a, acc, m0, n0, wr, wc, fg, ft);
}

// ------------------------------------------------------------------------------------------------ pipelined DMMA variant
// Same m16n8k16 inner product, but the operand tiles travel global -> shared memory with cp.async (16-byte chunks, no
// register staging) through a three-stage ring, so two k-blocks of loads are always in flight behind the MMAs and the
// only barrier per k-block is the one that hands a filled stage over.  Needs 16-byte aligned operands with an even
// leading dimension (every call of the BaM solve; anything else takes dgemm_mma16_kernel).
constexpr int PSTAGES = 4;  // two k-blocks are computed per barrier while the two after them are in flight
constexpr int PSTAGE_DOUBLES = 2 * 128 * (DBK + 4);  // A and B tiles, the larger ([row][k+4]) of the two layouts each
constexpr int PIPE_SMEM_BYTES = PSTAGES * PSTAGE_DOUBLES * static_cast<int>(sizeof(double));

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// one 128 x DBK operand tile (rows r0.., k from k0) into its stage buffer; out-of-range elements are zero-filled
template <bool MN>
__device__ __forceinline__ void pipe_load(double* __restrict__ S, const double* __restrict__ P, long long ld, int nrows, int K,
                                          int r0, int k0) {
#pragma unroll
  for (int e = 0; e < 128 * DBK / 2 / DTHREADS; ++e) {  // 16-byte chunks: 4 per thread
    const int c = threadIdx.x + e * DTHREADS;
    int r, k, nr, nk;
    if (MN) {  // chunk = two consecutive rows at one k
      k = c / 64;
      r = 2 * (c % 64);
      nr = nrows - (r0 + r);
      nk = K - (k0 + k);
      const int bytes = (nk > 0 && nr > 0) ? (nr >= 2 ? 16 : 8) : 0;
      const double* src = bytes ? P + static_cast<long long>(k0 + k) * ld + r0 + r : P;
      cp_async16(S + k * (128 + 4) + r, src, bytes);
    } else {   // chunk = two consecutive k of one row
      r = c / 8;
      k = 2 * (c % 8);
      nr = nrows - (r0 + r);
      nk = K - (k0 + k);
      const int bytes = (nk > 0 && nr > 0) ? (nk >= 2 ? 16 : 8) : 0;
      const double* src = bytes ? P + static_cast<long long>(r0 + r) * ld + k0 + k : P;
      cp_async16(S + r * (DBK + 4) + k, src, bytes);
    }
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(DTHREADS, 1) dgemm_pipe_kernel(const DgemmArgs a) {
  constexpr int DBM = 128, DBN = 128;
  constexpr int WM = DBM / 2, WN = DBN / 4;
  constexpr int BI = WM / 16, BJ = WN / 8;
  extern __shared__ __align__(16) double psm[];
  int tm, tn;
  {
    const int tiles_n = (a.N + DBN - 1) / DBN;
    tm = blockIdx.x / tiles_n;
    tn = blockIdx.x % tiles_n;
    if (a.tri && tn > tm) return;
  }
  const int m0 = tm * DBM, n0 = tn * DBN;
  int k_begin = 0, k_end = a.K;
  if (a.krange & KR_A_LOWER) k_end = min(k_end, m0 + DBM);
  if (a.krange & KR_B_LOWER) k_end = min(k_end, n0 + DBN);
  if (a.krange & KR_A_UPPER) k_begin = max(k_begin, m0);
  if (a.krange & KR_B_UPPER) k_begin = max(k_begin, n0);
  k_begin = (k_begin / DBK) * DBK;
  const int nkb = (k_end > k_begin) ? (k_end - k_begin + DBK - 1) / DBK : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 2) * WM, wc = (warp & 3) * WN;
  const int fg = lane >> 2, ft = lane & 3;
  double acc[BI][BJ][4];
#pragma unroll
  for (int i = 0; i < BI; ++i)
#pragma unroll
    for (int j = 0; j < BJ; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[i][j][u] = 0.0;

  auto stage_a = [&](int s) { return psm + s * PSTAGE_DOUBLES; };
  auto stage_b = [&](int s) { return psm + s * PSTAGE_DOUBLES + 128 * (DBK + 4); };
  // Two k-blocks per hand-over: with two warps per scheduler a barrier per 16-deep k-block left the FP64 tensor pipe idle
  // 16% of the cycles (ncu: 83.8% active); the ring holds the pair being computed and the pair in flight.
  auto load_block = [&](int kbi) {
    if (kbi < nkb) {
      pipe_load<A_MN>(stage_a(kbi % PSTAGES), a.A, a.lda, a.M, a.K, m0, k_begin + kbi * DBK);
      pipe_load<B_MN>(stage_b(kbi % PSTAGES), a.B, a.ldb, a.N, a.K, n0, k_begin + kbi * DBK);
    }
  };
  auto compute_block = [&](int kbi) {
    const double* As = stage_a(kbi % PSTAGES);
    const double* Bs = stage_b(kbi % PSTAGES);
    double af[BI][8], bf[BJ][4];
#pragma unroll
    for (int i = 0; i < BI; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) af[i][e] = As[sidx<DBM, A_MN>(wr + 16 * i + fg + 8 * (e & 1), ft + 4 * (e >> 1))];
#pragma unroll
    for (int j = 0; j < BJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) bf[j][e] = Bs[sidx<DBN, B_MN>(wc + 8 * j + fg, ft + 4 * e)];
#pragma unroll
    for (int i = 0; i < BI; ++i)
#pragma unroll
      for (int j = 0; j < BJ; ++j) dmma16816(acc[i][j], af[i], bf[j]);
  };
  load_block(0);
  load_block(1);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int kb = 0; kb < nkb; kb += 2) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // k-blocks kb, kb+1 have landed
    __syncthreads();  // ... for every thread, and everyone is done with the two stages that are refilled next
    load_block(kb + 2);
    load_block(kb + 3);
    asm volatile("cp.async.commit_group;" ::: "memory");
    compute_block(kb);
    if (kb + 1 < nkb) compute_block(kb + 1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  dgemm_epilogue<BI, BJ>(a, acc, m0, n0, wr, wc, fg, ft);
}

int launch_dgemm(cudaStream_t stream, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                 long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o) {
  if (M <= 0 || N <= 0 || K < 0 || (K > 0 && (!A || !B))) return GSMVI_EINVAL;
  if (o.ncp == 0 && !C) return GSMVI_EINVAL;
  if (o.beta != 0.0 && !o.Cin) return GSMVI_EINVAL;
  if (o.ncp < 0 || o.ncp > DGEMM_MAX_PEERS || o.row0 < 0) return GSMVI_EINVAL;
  if ((o.ncp > 0 || o.row0 > 0) && (o.tri || o.mirror || o.krange != KR_FULL)) return GSMVI_EINVAL;
  DgemmArgs a;
  a.M = M; a.N = N; a.K = K;
  a.alpha = o.alpha; a.beta = o.beta; a.diag_add = o.diag_add;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb;
  a.Cin = o.Cin; a.ldcin = o.ldcin; a.C = C; a.ldc = ldc;
  a.tri = o.tri ? 1 : 0; a.mirror = o.mirror ? 1 : 0; a.krange = o.krange;
  a.row0 = o.row0; a.ncp = o.ncp;
  for (int d = 0; d < DGEMM_MAX_PEERS; ++d) a.Cp[d] = d < o.ncp ? o.Cp[d] : nullptr;
  for (int d = 0; d < o.ncp; ++d)
    if (!a.Cp[d]) return GSMVI_EINVAL;
  const long long tiles128 = static_cast<long long>((M + 127) / 128) * ((N + 127) / 128);
  const bool big = (o.tri ? tiles128 / 2 : tiles128) >= 120;  // enough 128x128 tiles to fill 148 SMs
  const bool aligned = ((lda | ldb) & 1) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
  if (big && aligned && K > 0) {
    // cp.async three-stage pipeline, 128 x 128 tiles
    static PerDeviceOnce attr_set;
    if (!attr_set.get()) {
      cudaError_t ae = cudaFuncSetAttribute(dgemm_pipe_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES);
      if (ae == cudaSuccess) ae = cudaFuncSetAttribute(dgemm_pipe_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES);
      if (ae == cudaSuccess) ae = cudaFuncSetAttribute(dgemm_pipe_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES);
      if (ae == cudaSuccess) ae = cudaFuncSetAttribute(dgemm_pipe_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES);
      if (ae != cudaSuccess) return static_cast<int>(ae);
      attr_set.set();
    }
    const int grid = ((M + 127) / 128) * ((N + 127) / 128);
    if (!a_mn && !b_mn) dgemm_pipe_kernel<false, false><<<grid, DTHREADS, PIPE_SMEM_BYTES, stream>>>(a);
    else if (a_mn && !b_mn) dgemm_pipe_kernel<true, false><<<grid, DTHREADS, PIPE_SMEM_BYTES, stream>>>(a);
    else if (!a_mn && b_mn) dgemm_pipe_kernel<false, true><<<grid, DTHREADS, PIPE_SMEM_BYTES, stream>>>(a);
    else dgemm_pipe_kernel<true, true><<<grid, DTHREADS, PIPE_SMEM_BYTES, stream>>>(a);
  } else {
    // register-staged kernel: any alignment; 64 x 64 tiles (two CTAs per SM) when 128-tiles would not fill the GPU
#define GSMVI_DG(BMV)                                                                                  \
  {                                                                                                    \
    const int grid = ((M + BMV - 1) / BMV) * ((N + BMV - 1) / BMV);                                    \
    if (!a_mn && !b_mn) dgemm_mma16_kernel<BMV, false, false><<<grid, DTHREADS, 0, stream>>>(a);        \
    else if (a_mn && !b_mn) dgemm_mma16_kernel<BMV, true, false><<<grid, DTHREADS, 0, stream>>>(a);     \
    else if (!a_mn && b_mn) dgemm_mma16_kernel<BMV, false, true><<<grid, DTHREADS, 0, stream>>>(a);     \
    else dgemm_mma16_kernel<BMV, true, true><<<grid, DTHREADS, 0, stream>>>(a);                         \
  }
    if (big) GSMVI_DG(128) else GSMVI_DG(64)
#undef GSMVI_DG
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
