// Host side of the scaled 3xFP16 GEMM (h3_gemm.cuh): fp16 tensor maps, dispatch, and the operand split kernels.
// The contractions it serves: sampler x = mu + z L^T (gsmvi/gsm.py:119), dense-Gaussian score (examples/
// example_gsm_numpy.py:24-29), W = G Sigma and the batch-mean covariance update of gsm_update (gsmvi/gsm.py:11-27, 53-54).
#include "dev_once.cuh"
#include "h3_gemm.cuh"
#include "h3x2_gemm.cuh"
#include <cstdlib>

namespace gsmvi {

static int make_tmap_h(CUtensorMap* out, const __half* ptr, long long rows, long long cols, long long ld, int box_cols,
                       int box_rows) {
  static PFN_tmapEncodeTiled enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return GSMVI_EDRIVER;
    enc = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  if (rows <= 0 || cols <= 0) return GSMVI_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & 7) != 0 || ld < cols) return GSMVI_EALIGN;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(__half)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GSMVI_OK : GSMVI_EDRIVER;
}

template <bool A_MN, bool B_MN>
static int launch_h3_one(cudaStream_t stream, const H3Args& args, const CUtensorMap& tah, const CUtensorMap& tbh,
                         const CUtensorMap& tal, const CUtensorMap& tbl, dim3 grid, bool pdl) {
  static PerDeviceOnce attr_set;
  auto kern = gemm_h3_kernel<A_MN, B_MN>;
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(H3_THREADS);
    cfg.dynamicSmemBytes = H3_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args, tah, tbh, tal, tbl);
    return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
  }
  kern<<<grid, H3_THREADS, H3_SMEM_BYTES, stream>>>(args, tah, tbh, tal, tbl);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int h3_make_tmap(CUtensorMap* out, const __half* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  return make_tmap_h(out, ptr, rows, cols, ld, box_cols, box_rows);
}

int h3_prepare(int M, int N, int K, const HView& A, const HView& B, float* C, long long ldc, const H3Opts& o, H3Args* out,
               CUtensorMap* maps, dim3* grid_out) {
  if (M <= 0 || N <= 0 || K <= 0 || (!C && !o.push_base) || !A.hi || !A.lo || !B.hi || !B.lo || !A.scale || !B.scale) return GSMVI_EINVAL;
  if (o.tri && M != N) return GSMVI_EINVAL;
  if (o.beta != 0.0f && !o.Cin) return GSMVI_EINVAL;
  if (o.splits < 1 || (o.splits > 1 && (o.beta != 0.0f || o.bias_n || o.mirror))) return GSMVI_EINVAL;
  H3Args& a = *out;
  a.M = M; a.N = N; a.K = K;
  a.alpha = o.alpha; a.beta = o.beta;
  a.Cin = o.Cin; a.ldcin = o.ldcin;
  a.C = C; a.ldc = ldc;
  a.bias_n = o.bias_n;
  a.scale_a = A.scale; a.scale_b = B.scale;
  a.absmax_out = o.absmax_out;
  a.tri = o.tri ? 1 : 0;
  a.mirror = o.mirror ? 1 : 0;
  a.krange = o.krange;
  a.tiles_m = (M + H3_BM - 1) / H3_BM;
  a.tiles_n = (N + H3_BN - 1) / H3_BN;
  a.splits = o.splits;
  a.split_stride = o.split_stride;
  a.push_base = o.push_base;
  a.push_stage_off = o.push_stage_off;
  a.push_cnt_off = o.push_cnt_off;
  a.push_rank = o.push_rank;
  a.push_world = o.push_world;
  a.push_tpo = o.push_tpo;
  a.split_hi = o.split_hi;
  a.split_lo = o.split_lo;
  a.split_ld = o.split_ld;
  a.split_scale = o.split_scale;
  if (o.split_hi && (!o.split_lo || !o.split_scale || (o.split_ld & 3) != 0 || o.tri || o.splits != 1 || o.push_base)) return GSMVI_EINVAL;
  if (o.push_base && (!o.tri || o.splits != 1 || o.mirror || o.beta != 0.0f || o.bias_n)) return GSMVI_EINVAL;
  const int tiles = o.tri ? a.tiles_m * (a.tiles_m + 1) / 2 : a.tiles_m * a.tiles_n;
  *grid_out = dim3(tiles, o.splits);

  int rc;
  // K-major: box = 64 K-elements (128 B) x 128 rows.  MN-major: box = 64 MN-elements (128 B) x 64 K-rows, two per tile.
  const int abc = 64, abr = o.a_mn ? H3_BK : H3_BM, bbr = o.b_mn ? H3_BK : H3_BN;
  if ((rc = make_tmap_h(&maps[0], A.hi, A.rows, A.cols, A.ld, abc, abr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&maps[1], B.hi, B.rows, B.cols, B.ld, abc, bbr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&maps[2], A.lo, A.rows, A.cols, A.ld, abc, abr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&maps[3], B.lo, B.rows, B.cols, B.ld, abc, bbr)) != GSMVI_OK) return rc;
  return GSMVI_OK;
}

// ---- persistent 2-CTA kernel (h3x2_gemm.cuh): used for every launch it can express, GSMVI_H3X2=0 turns it off

template <bool A_MN, bool B_MN, bool PUSH>
static int launch_h3x2_one(cudaStream_t stream, const H3Args& args, const CUtensorMap& tah, const CUtensorMap& tbh,
                           const CUtensorMap& tal, const CUtensorMap& tbl, int tiles_mp, int n_st, bool pdl) {
  static PerDeviceOnce attr_set;
  static PerDeviceInt max_pairs;
  constexpr int SMEM = PUSH ? X2_PUSH_SMEM_BYTES : X2_SMEM_BYTES;
  auto kern = gemm_h3x2_kernel<A_MN, B_MN, PUSH>;
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    // how many pairs fit at once (one CTA per SM, two SMs of one TPC per pair): the persistent grid
    int dev = 0, sms = 0, clusters = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(2 * (sms / 2));
    q.blockDim = dim3(X2_THREADS);
    q.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&clusters, kern, &q) != cudaSuccess || clusters <= 0) {
      cudaGetLastError();
      clusters = sms / 2;
    }
    max_pairs.ref() = clusters < sms / 2 ? clusters : sms / 2;
    attr_set.set();
  }
  const int pairs = n_st < max_pairs.ref() ? n_st : max_pairs.ref();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(X2_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;  // the cluster shape is a compile-time attribute of the kernel (__cluster_dims__)
  int chunk_kb = X2_CHUNK_KB;
  if (const char* cs = getenv("GSMVI_X2_CHUNK_KB")) {  // accuracy / speed probe: additions per TMEM accumulator = 4 * chunk_kb
    const int v = atoi(cs);
    if (v >= 1 && v <= 64) chunk_kb = v;
  }
  int probe = 0;  // GSMVI_X2_PROBE: bottleneck probes of tools/probe_gemm_h3x2.py (bit 0: no TMA loads, bit 1: hi*hi MMAs only)
  if (const char* ps = getenv("GSMVI_X2_PROBE")) probe = atoi(ps) & 3;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args, tah, tbh, tal, tbl, tiles_mp, n_st, chunk_kb, probe);
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

static int g_h3x2 = -1;  // -1: read GSMVI_H3X2 on first use
static bool h3x2_enabled() {
  if (g_h3x2 < 0) {
    const char* s = getenv("GSMVI_H3X2");
    g_h3x2 = (s && s[0] == '0') ? 0 : 1;
  }
  return g_h3x2 == 1;
}
int h3_pair_kernel(int enable) {
  const int prev = h3x2_enabled() ? 1 : 0;
  if (enable == 0 || enable == 1) g_h3x2 = enable;
  return prev;
}

static bool h3x2_eligible(int M, const H3Opts& o) {
  if (o.push_base && !(o.tri && o.a_mn && o.b_mn)) return false;  // push mode: the covariance update's MN-major form only
  return h3x2_enabled() && M > H3_BM && o.splits == 1 && (o.krange & ~KR_B_LOWER) == 0;
}

static int launch_gemm_h3x2(cudaStream_t stream, int M, int N, int K, const HView& A, const HView& B, float* C, long long ldc,
                            const H3Opts& o) {
  H3Args a;
  CUtensorMap tm[4];
  dim3 grid;
  int rc = h3_prepare(M, N, K, A, B, C, ldc, o, &a, tm, &grid);  // argument block + validation (its maps are rebuilt below)
  if (rc != GSMVI_OK) return rc;
  // per CTA and k-block: 128 rows of A, 64 rows of B.  K-major: box = 64 K-elements x rows.  MN-major: 64 MN-elements x 64 K-rows.
  const int abr = o.a_mn ? H3_BK : 128, bbr = 64;
  if ((rc = make_tmap_h(&tm[0], A.hi, A.rows, A.cols, A.ld, 64, abr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&tm[1], B.hi, B.rows, B.cols, B.ld, 64, bbr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&tm[2], A.lo, A.rows, A.cols, A.ld, 64, abr)) != GSMVI_OK) return rc;
  if ((rc = make_tmap_h(&tm[3], B.lo, B.rows, B.cols, B.ld, 64, bbr)) != GSMVI_OK) return rc;
  const int tiles_mp = (a.tiles_m + 1) / 2;
  const int n_st = o.tri ? tiles_mp * (tiles_mp + 1) : tiles_mp * a.tiles_n;
  if (o.push_base) return launch_h3x2_one<true, true, true>(stream, a, tm[0], tm[1], tm[2], tm[3], tiles_mp, n_st, o.pdl);
  if (!o.a_mn && !o.b_mn) return launch_h3x2_one<false, false, false>(stream, a, tm[0], tm[1], tm[2], tm[3], tiles_mp, n_st, o.pdl);
  if (o.a_mn && !o.b_mn) return launch_h3x2_one<true, false, false>(stream, a, tm[0], tm[1], tm[2], tm[3], tiles_mp, n_st, o.pdl);
  if (!o.a_mn && o.b_mn) return launch_h3x2_one<false, true, false>(stream, a, tm[0], tm[1], tm[2], tm[3], tiles_mp, n_st, o.pdl);
  return launch_h3x2_one<true, true, false>(stream, a, tm[0], tm[1], tm[2], tm[3], tiles_mp, n_st, o.pdl);
}

int launch_gemm_h3(cudaStream_t stream, int M, int N, int K, const HView& A, const HView& B, float* C, long long ldc,
                   const H3Opts& o) {
  if (h3x2_eligible(M, o)) return launch_gemm_h3x2(stream, M, N, K, A, B, C, ldc, o);
  H3Args a;
  CUtensorMap tm[4];  // A_hi, B_hi, A_lo, B_lo
  dim3 grid;
  const int rc = h3_prepare(M, N, K, A, B, C, ldc, o, &a, tm, &grid);
  if (rc != GSMVI_OK) return rc;
  if (!o.a_mn && !o.b_mn) return launch_h3_one<false, false>(stream, a, tm[0], tm[1], tm[2], tm[3], grid, o.pdl);
  if (o.a_mn && !o.b_mn) return launch_h3_one<true, false>(stream, a, tm[0], tm[1], tm[2], tm[3], grid, o.pdl);
  if (!o.a_mn && o.b_mn) return launch_h3_one<false, true>(stream, a, tm[0], tm[1], tm[2], tm[3], grid, o.pdl);
  return launch_h3_one<true, true>(stream, a, tm[0], tm[1], tm[2], tm[3], grid, o.pdl);
}

// ------------------------------------------------------------------------------------------------ operand split

__global__ void __launch_bounds__(256) h3_absmax_kernel(const float* __restrict__ A, long long lda, int rows, int cols,
                                                        unsigned* __restrict__ out) {
  unsigned m = 0u;
  const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  for (long long i = blockIdx.y; i < rows; i += gridDim.y) {
    const float* row = A + i * lda;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4; j < cols; j += gridDim.x * blockDim.x * 4) {
      if (vec && j + 3 < cols) {
        const float4 v = *reinterpret_cast<const float4*>(row + j);
        m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
      } else {
        for (int t = 0; t < 4 && j + t < cols; ++t) m = max(m, __float_as_uint(fabsf(row[j + t])));
      }
    }
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(out, m);
}

// (six CTAs per SM = at most 42 registers: the split of the new covariance runs beside the Cholesky, whose CTAs leave 11.7k)
__global__ void __launch_bounds__(256, 6) h3_split_kernel(const float* __restrict__ A, long long lda, int rows, int cols,
                                                       const unsigned* __restrict__ absmax, int sqrt_mode,
                                                       float* __restrict__ scale_out, __half* __restrict__ Hi,
                                                       __half* __restrict__ Lo, long long ldo) {
  const float s = h3_scale_from_absmax(*absmax, sqrt_mode);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = s;
  const bool vec = ((lda & 3) == 0) && ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                   (((reinterpret_cast<uintptr_t>(Hi) | reinterpret_cast<uintptr_t>(Lo)) & 7) == 0);
  // four rows per trip (blockIdx.y walks row quads): four independent 16-byte loads in flight per thread
  for (long long i0 = 4LL * blockIdx.y; i0 < rows; i0 += 4LL * gridDim.y) {
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4; j < cols; j += gridDim.x * blockDim.x * 4) {
      if (vec && j + 3 < cols) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long i = (i0 + q < rows) ? i0 + q : rows - 1;
          v[q] = *reinterpret_cast<const float4*>(A + i * lda + j);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long i = i0 + q;
          if (i >= rows) break;
          __half h[4], l[4];
          h3_split1(v[q].x, s, h[0], l[0]);
          h3_split1(v[q].y, s, h[1], l[1]);
          h3_split1(v[q].z, s, h[2], l[2]);
          h3_split1(v[q].w, s, h[3], l[3]);
          *reinterpret_cast<uint2*>(Hi + i * ldo + j) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(Lo + i * ldo + j) = *reinterpret_cast<const uint2*>(l);
        }
      } else {
        for (int q = 0; q < 4 && i0 + q < rows; ++q)
          for (int t = 0; t < 4 && j + t < cols; ++t)
            h3_split1(A[(i0 + q) * lda + j + t], s, Hi[(i0 + q) * ldo + j + t], Lo[(i0 + q) * ldo + j + t]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) h3_bound_scales_kernel(const float* __restrict__ mu, int D,
                                                              const unsigned* __restrict__ sigma_absmax,
                                                              const unsigned* __restrict__ zmax_bits, float zmax_const,
                                                              float pnorm, float cmax, float* __restrict__ scale_x,
                                                              float* __restrict__ scale_g) {
  __shared__ unsigned smax[8];
  unsigned m = 0u;
  for (int i = threadIdx.x; i < D; i += 256) m = max(m, __float_as_uint(fabsf(mu[i])));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = max(m, smax[w]);
    const float zmax = zmax_bits ? __uint_as_float(*zmax_bits) : zmax_const;
    const float xb = __uint_as_float(m) + zmax * sqrtf(static_cast<float>(D)) * sqrtf(__uint_as_float(*sigma_absmax));
    *scale_x = h3_scale_from_absmax(__float_as_uint(xb), 0);
    if (scale_g) *scale_g = h3_scale_from_absmax(__float_as_uint(xb * pnorm + cmax), 0);
  }
}

int h3_bound_scales(cudaStream_t stream, const float* mu, int D, const unsigned* sigma_absmax, const unsigned* zmax_bits,
                    float zmax_const, float pnorm, float cmax, float* scale_x, float* scale_g) {
  if (!mu || D <= 0 || !sigma_absmax || !scale_x) return GSMVI_EINVAL;
  h3_bound_scales_kernel<<<1, 256, 0, stream>>>(mu, D, sigma_absmax, zmax_bits, zmax_const, pnorm, cmax, scale_x, scale_g);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

static inline dim3 rowwise_grid(int rows, int cols) {
  const int gx = (cols / 4 + 255) / 256 > 0 ? (cols / 4 + 255) / 256 : 1;
  int gy = rows;
  if (gy > 65535) gy = 65535;
  return dim3(gx, gy);
}

int h3_absmax(cudaStream_t stream, const float* A, long long lda, int rows, int cols, unsigned* out) {
  if (!A || !out || rows <= 0 || cols <= 0) return GSMVI_EINVAL;
  // a few rows per CTA keeps the atomic count low
  dim3 g = rowwise_grid(rows, cols);
  g.y = (rows + 7) / 8;
  if (g.y > 65535) g.y = 65535;
  h3_absmax_kernel<<<g, 256, 0, stream>>>(A, lda, rows, cols, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int h3_split(cudaStream_t stream, const float* A, long long lda, int rows, int cols, const unsigned* absmax, int sqrt_mode,
             float* scale_out, __half* Hi, __half* Lo, long long ldo) {
  if (!A || !absmax || !scale_out || !Hi || !Lo || rows <= 0 || cols <= 0) return GSMVI_EINVAL;
  dim3 g = rowwise_grid(rows, cols);
  g.y = (rows + 3) / 4 > 65535 ? 65535 : (rows + 3) / 4;
  h3_split_kernel<<<g, 256, 0, stream>>>(A, lda, rows, cols, absmax, sqrt_mode, scale_out, Hi, Lo, ldo);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
