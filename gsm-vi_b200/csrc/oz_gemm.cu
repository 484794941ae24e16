// FP64 GEMM on the int8 tensor cores (see oz_gemm.cuh): digit slicers, the tcgen05 kind::i8 kernel, the launcher.
#include "dev_once.cuh"
#include "oz_gemm.cuh"

#include <cuda.h>
#include <math.h>

#include "ptx.cuh"

namespace gsmvi {

namespace {

constexpr int OZ_BM = 128, OZ_BN = 128;
constexpr int OZ_BK = 128;                      // int8 per 128-byte swizzle row
constexpr int OZ_UMMA_K = 32;                   // kind::i8: 32 bytes of K per instruction
constexpr int OZ_TILE_BYTES = OZ_BM * OZ_BK;    // 16 KiB per operand tile
constexpr int OZ_STAGE_BYTES = 2 * OZ_TILE_BYTES;
constexpr int OZ_STAGES = 6;
constexpr int OZ_SMEM_BYTES = 1024 + BAR_BYTES + OZ_STAGES * OZ_STAGE_BYTES;
constexpr int OZ_THREADS = 320;
constexpr int OZ_MAX_SLICES = 8;

static inline long long rup(long long v, long long m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ digit slicers

// peel `s` signed 7-bit digits off v (|v| < 1): v = sum_t q_t 2^(-7t) + remainder; digit t of element e goes to byte e of w[t]
__device__ __forceinline__ void oz_digits(double v, int s, int e, unsigned long long (&w)[OZ_MAX_SLICES]) {
#pragma unroll
  for (int t = 0; t < OZ_MAX_SLICES; ++t) {
    if (t < s) {
      v *= 128.0;
      const int q = __double2int_rz(v);
      v -= static_cast<double>(q);
      w[t] |= static_cast<unsigned long long>(static_cast<unsigned char>(static_cast<signed char>(q))) << (8 * e);
    }
  }
}

__device__ __forceinline__ double oz_block_max(double m, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < 8; ++w) r = fmax(r, red[w]);
  __syncthreads();
  return r;
}

// operand stored [R, K] (K-major): one CTA per row
__global__ void __launch_bounds__(256) oz_slice_kmajor_kernel(const double* __restrict__ X, long long ld, int R, int K,
                                                              signed char* __restrict__ planes, long long plane_stride,
                                                              long long Kp, double* __restrict__ scale, int s) {
  __shared__ double red[8];
  const long long row = blockIdx.x;
  const double* x = X + row * ld;
  double m = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) m = fmax(m, fabs(x[k]));
  m = oz_block_max(m, red);
  int e = 0;
  if (m > 0.0 && m < 1.0e300) frexp(m, &e);  // m = f 2^e, f in [0.5, 1): |x| 2^-e < 1
  if (threadIdx.x == 0) scale[row] = ldexp(1.0, e);
  const double inv = ldexp(1.0, -e);
  for (int k8 = threadIdx.x * 8; k8 < K; k8 += 256 * 8) {
    unsigned long long w[OZ_MAX_SLICES] = {};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (k8 + j < K) oz_digits(x[k8 + j] * inv, s, j, w);
#pragma unroll
    for (int t = 0; t < OZ_MAX_SLICES; ++t)
      if (t < s) *reinterpret_cast<unsigned long long*>(planes + t * plane_stride + row * Kp + k8) = w[t];
  }
}

// operand stored [K, R] (MN-major): one CTA per 32 operand rows (= 32 consecutive storage columns), transposed through smem
__global__ void __launch_bounds__(256) oz_slice_mnmajor_kernel(const double* __restrict__ X, long long ld, int R, int K,
                                                               signed char* __restrict__ planes, long long plane_stride,
                                                               long long Kp, double* __restrict__ scale, int s) {
  __shared__ double tile[32][65];
  __shared__ double cmax[8][32];
  __shared__ double sinv[32];
  const int r0 = blockIdx.x * 32;
  const int c = threadIdx.x & 31, kl = threadIdx.x >> 5;
  double m = 0.0;
  if (r0 + c < R)
    for (int k = kl; k < K; k += 8) m = fmax(m, fabs(X[static_cast<long long>(k) * ld + r0 + c]));
  cmax[kl][c] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double mm = cmax[0][c];
    for (int j = 1; j < 8; ++j) mm = fmax(mm, cmax[j][c]);
    int e = 0;
    if (mm > 0.0 && mm < 1.0e300) frexp(mm, &e);
    if (r0 + c < R) scale[r0 + c] = ldexp(1.0, e);
    sinv[c] = ldexp(1.0, -e);
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, kseg = threadIdx.x & 7;
  for (int k0 = 0; k0 < K; k0 += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = kl + 8 * j;
      double v = 0.0;
      if (k0 + kk < K && r0 + c < R) v = X[static_cast<long long>(k0 + kk) * ld + r0 + c];
      tile[c][kk] = v;
    }
    __syncthreads();
    if (r0 + r < R) {
      unsigned long long w[OZ_MAX_SLICES] = {};
      const double inv = sinv[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) oz_digits(tile[r][8 * kseg + j] * inv, s, j, w);  // k beyond K holds zeros
      const long long off = static_cast<long long>(r0 + r) * Kp + k0 + 8 * kseg;
      if (k0 + 8 * kseg < Kp) {
#pragma unroll
        for (int t = 0; t < OZ_MAX_SLICES; ++t)
          if (t < s) *reinterpret_cast<unsigned long long*>(planes + t * plane_stride + off) = w[t];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ int8 GEMM

struct OzArgs {
  int M, N;
  int nkb;            // 128-byte k-blocks per operand plane
  int g;              // digit group: pairs (t, g - t), t = t_lo .. t_hi (1-based digits)
  int t_lo, t_hi;
  int rows_a, rows_b; // padded rows per plane
  double gscale;      // 2^(-7 g)
  int accum_in;       // add to Acc (else start from zero)
  int final;          // apply scales / alpha / beta / diag and write C (else write Acc)
  double* acc;
  long long ldacc;
  const double* sa;
  const double* sb;
  double alpha, beta, diag_add;
  const double* Cin;
  long long ldcin;
  double* C;
  long long ldc;
  int tri, mirror;
  int tiles_m, tiles_n;
};

__host__ __device__ constexpr uint32_t make_idesc_i8() {
  // c_format S32 (2) [4,6); a_format INT8 (1) [7,10); b_format INT8 (1) [10,13); both K-major; N>>3 [17,23); M>>4 [24,29)
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(OZ_BN >> 3) << 17) | (static_cast<uint32_t>(OZ_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
gemm_oz_kernel(const OzArgs args, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  constexpr int STAGES = OZ_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t bar_base = ptx::smem_u32(smem);
  const uint32_t stage_base = bar_base + BAR_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 8 * (2 * STAGES + 1) + 8);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  int tm, tn;
  if (args.tri) {
    const int t = blockIdx.x;
    int i = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    tm = i;
    tn = t - i * (i + 1) / 2;
  } else {
    constexpr int GROUP = 8;
    const int t = blockIdx.x;
    const int per_group = GROUP * args.tiles_n;
    const int g = t / per_group;
    const int first_m = g * GROUP;
    const int rows = min(GROUP, args.tiles_m - first_m);
    const int r = t - g * per_group;
    tm = first_m + r % rows;
    tn = r / rows;
  }
  const int m0 = tm * OZ_BM, n0 = tn * OZ_BN;
  const int npairs = args.t_hi - args.t_lo + 1;
  const int total = npairs * args.nkb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // producer and MMA warps stay converged, one elected lane issues (ptx::elect_one: descriptors in uniform registers)
  if (warp == 0) {
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      const int pair = it / args.nkb, kb = it - pair * args.nkb;
      const int t = args.t_lo + pair, u = args.g - t;  // 1-based digit indices
      const uint32_t sA = stage_base + s * OZ_STAGE_BYTES;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar(s), OZ_STAGE_BYTES);
        ptx::tma_load_2d(sA, &tmA, full_bar(s), kb * OZ_BK, (t - 1) * args.rows_a + m0);
        ptx::tma_load_2d(sA + OZ_TILE_BYTES, &tmB, full_bar(s), kb * OZ_BK, (u - 1) * args.rows_b + n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_i8();
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after_sync();
      const uint32_t sA = stage_base + s * OZ_STAGE_BYTES;
      const uint32_t sB = sA + OZ_TILE_BYTES;
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < OZ_BK / OZ_UMMA_K; ++kk) {
          const uint64_t da = make_smem_desc(sA + kk * OZ_UMMA_K, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sB + kk * OZ_UMMA_K, 16, 1024, 2);
          umma_i8(tmem_base, da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar(s));
        if (it + 1 == total) ptx::umma_commit(acc_bar);
      }
      __syncwarp();
    }
  } else {
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after_sync();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = m0 + q * 32 + lane;
    const bool diag_tile = args.tri && (tm == tn);
    const double sam = (args.final && m < args.M) ? args.alpha * args.sa[m] : 0.0;
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int c0 = half * 64 + chunk * 32;
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      const int nbase = n0 + c0;
      if (m < args.M && nbase < args.N) {
        double* arow = args.acc + static_cast<long long>(m) * args.ldacc + nbase;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n >= args.N || (diag_tile && n > m)) continue;
          double v = static_cast<double>(static_cast<int>(r[j])) * args.gscale;
          if (args.accum_in) v += arow[j];
          if (!args.final) {
            arow[j] = v;
          } else {
            double o = sam * args.sb[n] * v;
            if (args.beta != 0.0) o += args.beta * args.Cin[static_cast<long long>(m) * args.ldcin + n];
            if (m == n) o += args.diag_add;
            args.C[static_cast<long long>(m) * args.ldc + n] = o;
            if (args.mirror && n != m) args.C[static_cast<long long>(n) * args.ldc + m] = o;
          }
        }
      }
    }
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

int make_tmap_u8(CUtensorMap* out, const signed char* ptr, long long rows, long long cols_bytes) {
  static PFN_tmapEncodeTiled enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return GSMVI_EDRIVER;
    enc = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols_bytes), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols_bytes)};
  cuuint32_t box[2] = {OZ_BK, OZ_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GSMVI_OK : GSMVI_EDRIVER;
}

}  // namespace

size_t oz_workspace_bytes(int M, int N, int K, int slices) {
  const long long Kp = rup(K, OZ_BK), ra = rup(M, OZ_BM), rb = rup(N, OZ_BN);
  long long b = 0;
  b += rup(static_cast<long long>(slices) * ra * Kp, 1024);
  b += rup(static_cast<long long>(slices) * rb * Kp, 1024);
  b += rup(static_cast<long long>(M) * rup(N, 2) * 8, 1024);
  b += rup((ra + rb) * 8, 1024);
  return static_cast<size_t>(b);
}

int launch_dgemm_oz(cudaStream_t stream, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                    long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o, void* ws, int slices) {
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C || !ws || slices < 2 || slices > OZ_MAX_SLICES) return GSMVI_EINVAL;
  if (K > 8192) return GSMVI_EINVAL;  // int32 accumulator: 8 pairs x K x 127^2 < 2^31
  if (o.tri && M != N) return GSMVI_EINVAL;
  if (o.beta != 0.0 && !o.Cin) return GSMVI_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return GSMVI_EALIGN;
  const long long Kp = rup(K, OZ_BK), ra = rup(M, OZ_BM), rb = rup(N, OZ_BN);
  signed char* pa = static_cast<signed char*>(ws);
  signed char* pb = pa + rup(static_cast<long long>(slices) * ra * Kp, 1024);
  double* acc = reinterpret_cast<double*>(pb + rup(static_cast<long long>(slices) * rb * Kp, 1024));
  const long long ldacc = rup(N, 2);
  double* sa = acc + rup(static_cast<long long>(M) * ldacc * 8, 1024) / 8;
  double* sb = sa + ra;
  cudaError_t e;
  if (ra != M || Kp != K) {
    e = cudaMemsetAsync(pa, 0, static_cast<size_t>(slices) * ra * Kp, stream);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (rb != N || Kp != K) {
    e = cudaMemsetAsync(pb, 0, static_cast<size_t>(slices) * rb * Kp, stream);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const bool same = (A == B && lda == ldb && a_mn == b_mn && M == N);  // Gram products: slice once
  if (a_mn) oz_slice_mnmajor_kernel<<<(M + 31) / 32, 256, 0, stream>>>(A, lda, M, K, pa, ra * Kp, Kp, sa, slices);
  else oz_slice_kmajor_kernel<<<M, 256, 0, stream>>>(A, lda, M, K, pa, ra * Kp, Kp, sa, slices);
  if (same) {
    pb = pa;
    sb = sa;
  } else if (b_mn) {
    oz_slice_mnmajor_kernel<<<(N + 31) / 32, 256, 0, stream>>>(B, ldb, N, K, pb, rb * Kp, Kp, sb, slices);
  } else {
    oz_slice_kmajor_kernel<<<N, 256, 0, stream>>>(B, ldb, N, K, pb, rb * Kp, Kp, sb, slices);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return static_cast<int>(e);

  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_u8(&ta, pa, static_cast<long long>(slices) * ra, Kp)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_u8(&tb, pb, static_cast<long long>(slices) * (same ? ra : rb), Kp)) != GSMVI_OK) return rc;
  static PerDeviceOnce attr_set;
  if (!attr_set.get()) {
    e = cudaFuncSetAttribute(gemm_oz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  OzArgs a;
  a.M = M; a.N = N;
  a.nkb = static_cast<int>(Kp / OZ_BK);
  a.rows_a = static_cast<int>(ra);
  a.rows_b = static_cast<int>(same ? ra : rb);
  a.acc = acc; a.ldacc = ldacc;
  a.sa = sa; a.sb = sb;
  a.alpha = o.alpha; a.beta = o.beta; a.diag_add = o.diag_add;
  a.Cin = o.Cin; a.ldcin = o.ldcin;
  a.C = C; a.ldc = ldc;
  a.tri = o.tri ? 1 : 0;
  a.mirror = o.mirror ? 1 : 0;
  a.tiles_m = (M + OZ_BM - 1) / OZ_BM;
  a.tiles_n = (N + OZ_BN - 1) / OZ_BN;
  const int grid = o.tri ? a.tiles_m * (a.tiles_m + 1) / 2 : a.tiles_m * a.tiles_n;
  // least significant digit group first, so the fp64 running sum loses nothing that matters
  for (int g = slices + 1; g >= 2; --g) {
    a.g = g;
    a.t_lo = g - slices > 1 ? g - slices : 1;
    a.t_hi = g - 1 < slices ? g - 1 : slices;
    a.gscale = ldexp(1.0, -7 * g);
    a.accum_in = (g != slices + 1) ? 1 : 0;
    a.final = (g == 2) ? 1 : 0;
    gemm_oz_kernel<<<grid, OZ_THREADS, OZ_SMEM_BYTES, stream>>>(a, ta, tb);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
