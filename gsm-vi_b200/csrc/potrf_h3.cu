// Left-looking blocked Cholesky  Sigma = L L^T  (fp32 result + the fp16 (hi, lo) split of L) with a device-side
// "is positive definite" flag - the factorisation behind the PD check (gsmvi/gsm.py:136-150, bam.py:219-233: host
// np.linalg.cholesky) and the sampler (gsmvi/gsm.py:119, bam.py:193: an SVD inside np.random.multivariate_normal).
//
// Per 128-column panel k (columns j0 = 128 k ...), two launches:
//   (1) the panel's accumulated update  P = L[j0:, 0:j0] L[j0:j0+128, 0:j0]^T  on the scaled 3xFP16 tensor-core engine
//       (h3_gemm.cuh), split-K over up to 8 CTAs per tile; each split writes its raw partial (deterministic: the panel
//       kernel adds the partials in a fixed order - no atomics, so replicated ranks factor bit-identically);
//   (2) the panel kernel: CTA 0 forms A11 - sum P and factors the 128 x 128 diagonal block 32 columns at a time,
//       publishing each finished block-row through a release epoch; the other CTAs (32 rows of the panel each) form
//       A21 - sum P while CTA 0 works, then solve X L11^T = A21 block-column by block-column as the epochs arrive, so
//       only the last 32-column stage trails the diagonal factorisation.  Every CTA writes its part of L as fp32 and
//       as the fp16 pair, which is what the next panels' update GEMMs (and the sampler) load by TMA.
// Left-looking means the trailing matrix is never rewritten: total update traffic is ~D^3/(3*128) operand bytes read
// once instead of a D^2 read-modify-write per panel, and each update is one long-K GEMM instead of a K=128 sliver.
#include "potrf.cuh"
#include "chol_block.cuh"
#include "h3_gemm.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace gsmvi {

namespace {

constexpr int NB = 128;
constexpr int RPC = 32;           // rows of the panel per TRSM CTA
constexpr int TPR = 256 / RPC;    // threads per row in the block-update phase
constexpr int DS = NB + 4;        // smem leading dimension (rows 16-byte aligned, quarter-warps on distinct banks)
constexpr int DT = 36;            // leading dimension of the transposed 32 x 32 diagonal block
constexpr int MAX_SPLITS = 8;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// acc[r][c] += sum_k Arows[r * lda_][k] * Brows[c * ldb_][k]   (both row-major over k; kbeg, kend multiples of 4)
__device__ __forceinline__ void tile4x4_mac(const float* __restrict__ Arows, int lda_, const float* __restrict__ Brows,
                                            int ldb_, int kbeg, int kend, float (&acc)[4][4]) {
  for (int k = kbeg; k < kend; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(Arows + r * lda_ + k);
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(Brows + c * ldb_ + k);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        acc[r][c] += a[r].x * b[c].x + a[r].y * b[c].y + a[r].z * b[c].z + a[r].w * b[c].w;
  }
}

// Forward substitution of one row against a 32 x 32 lower-triangular block, right-looking so that the dependent chain
// per column is one multiply and one FMA:  x_j = v_j / l_jj;  v_k -= x_j l_kj (k > j).  dT holds the block TRANSPOSED
// (dT[j * DT + k] = l_kj) so the column below the pivot is read as broadcast float4s.
__device__ __forceinline__ void row_solve32(float (&v)[32], const float* __restrict__ dT, const float* __restrict__ dinv) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = v[j] * dinv[j];
    v[j] = x;
    const float* col = dT + j * DT;
#pragma unroll
    for (int k = (j + 1) & ~3; k < 32; k += 4) {
      const float4 l = *reinterpret_cast<const float4*>(col + k);
      if (k + 0 > j) v[k + 0] -= x * l.x;
      if (k + 1 > j) v[k + 1] -= x * l.y;
      if (k + 2 > j) v[k + 2] -= x * l.z;
      if (k + 3 > j) v[k + 3] -= x * l.w;
    }
  }
}

// one 16-byte group of a finished row of L: fp32 store + the fp16 pair
__device__ __forceinline__ void store_l4(float* __restrict__ Lrow, __half* __restrict__ hrow, __half* __restrict__ lrow,
                                         int j, const float4 t, float s) {
  *reinterpret_cast<float4*>(Lrow + j) = t;
  __half h[4], l[4];
  h3_split1(t.x, s, h[0], l[0]);
  h3_split1(t.y, s, h[1], l[1]);
  h3_split1(t.z, s, h[2], l[2]);
  h3_split1(t.w, s, h[3], l[3]);
  *reinterpret_cast<uint2*>(hrow + j) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(lrow + j) = *reinterpret_cast<const uint2*>(l);
}

// scale of the fp16 split of L from max |Sigma_ii| (|L_ij| <= sqrt(Sigma_ii)); also resets the epoch word.
__global__ void __launch_bounds__(256) potrf_prepare_kernel(const float* __restrict__ A, long long lda, int n,
                                                            float* __restrict__ scale_l, unsigned* __restrict__ ready,
                                                            int* __restrict__ flag) {
  __shared__ unsigned smax[8];
  unsigned m = 0u;
  for (int i = threadIdx.x; i < n; i += 256) m = max(m, __float_as_uint(fabsf(A[static_cast<long long>(i) * lda + i])));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = max(m, smax[w]);
    *scale_l = h3_scale_from_absmax(m, 1);
    *ready = 0u;
    *flag = 0;
  }
}

// zero the strict upper triangle of L (fp32 and the fp16 pair) outside the diagonal blocks, which the panel kernel writes
__global__ void potrf_zero_upper_kernel(float* __restrict__ L, long long ldl, __half* __restrict__ Lhi,
                                        __half* __restrict__ Llo, long long ldh, int n) {
  const int i = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int jstart = (i / NB + 1) * NB;
  if (j < jstart || j >= n) return;
  for (int t = 0; t < 4 && j + t < n; ++t) {
    L[static_cast<long long>(i) * ldl + j + t] = 0.0f;
    Lhi[static_cast<long long>(i) * ldh + j + t] = __float2half_rn(0.0f);
    Llo[static_cast<long long>(i) * ldh + j + t] = __float2half_rn(0.0f);
  }
}

// optional phase timing of one panel (GSMVI_POTRF_TIMING=1): clock64 stamps of CTA 0 / CTA 1, read back by the host
__device__ long long g_pt3[32];
#define PT3(i) do { if (a.timing && threadIdx.x == 0) g_pt3[i] = clock64(); } while (0)

struct PanelArgs {
  const float* A;
  long long lda;
  float* L;
  long long ldl;
  __half* Lhi;
  __half* Llo;
  long long ldh;
  const float* scale_l;
  int n, j0, nb;
  const float* partials;  // [splits][n - j0][128] raw partial products of the update GEMM (row 0 = row j0 of the matrix)
  int splits;
  long long split_stride;
  int* flag;
  unsigned* ready;
  unsigned epoch_base;  // this panel publishes epoch_base + 1 .. epoch_base + 7
  int timing;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// spin (thread 0) until the panel's epoch word reaches `target`, then release the whole CTA
__device__ __forceinline__ void wait_epoch(const unsigned* ready, unsigned target, int j0) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_u32(ready) - target) < 0) {
      __nanosleep(20);
      if (clock64() - t0 > 4000000000LL) { printf("gsmvi: potrf panel watchdog (j0=%d epoch %u)\n", j0, target); __trap(); }
    }
  }
  __syncthreads();
}

// Epochs of one panel (relative to epoch_base): 2p+1 = diagonal block p factored and stored, 2p+2 = the rows below it in
// block-column p solved and stored.  A TRSM CTA needs 2J for the block-update of its stage J and 2J+1 for the in-block
// substitution, so only that last substitution trails CTA 0.
__global__ void __launch_bounds__(256, 2) potrf_panel_h3_kernel(const PanelArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s = sm;                   // [NB][DS]   diagonal block (CTA 0) / published block-rows of L11 (TRSM CTAs)
  float* at = sm + NB * DS;        // [RPC][DS]  TRSM CTAs: their rows of the panel
  float* dT = at + RPC * DS;       // [32][DT]   current 32 x 32 diagonal block, transposed
  __shared__ float dinv[NB];
  __shared__ int bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = a.j0, nb = a.nb, n = a.n;
  const float sl = *a.scale_l;
  const bool full = (nb == NB);
  if (tid == 0) bad = 0;

  if (blockIdx.x == 0) {
    // ------------------------------------------------------------------ diagonal block
    PT3(0);
    const float* a11 = a.A + static_cast<long long>(j0) * a.lda + j0;
    if (full) {
      // A11 - sum of the update partials, lower-triangle 16-byte groups only: 2112 groups, compactly enumerated (row i has
      // i/4 + 1 groups), eight per thread with all loads of a split issued before any is consumed
      constexpr int NG = 2112;
      float4 v[9];
      int gi[9], gj[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const int g = tid + e * 256;
        // row i = 4 b + r holds groups [2 b (b+1) + r (b+1), ...): invert with b = floor((sqrt(1 + 2 g) - 1) / 2) then fix up
        int b = static_cast<int>((sqrtf(1.0f + 2.0f * g) - 1.0f) * 0.5f);
        while (2 * (b + 1) * (b + 2) <= g) ++b;
        while (2 * b * (b + 1) > g) --b;
        const int rem = g - 2 * b * (b + 1);
        gi[e] = 4 * b + rem / (b + 1);
        gj[e] = 4 * (rem % (b + 1));
        v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < NG) v[e] = *reinterpret_cast<const float4*>(a11 + static_cast<long long>(gi[e]) * a.lda + gj[e]);
      }
      for (int sp = 0; sp < a.splits; ++sp) {
        const float* pb = a.partials + sp * a.split_stride;
        float4 pv[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) {
          pv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tid + e * 256 < NG) pv[e] = *reinterpret_cast<const float4*>(pb + gi[e] * NB + gj[e]);
        }
#pragma unroll
        for (int e = 0; e < 9; ++e) { v[e].x -= pv[e].x; v[e].y -= pv[e].y; v[e].z -= pv[e].z; v[e].w -= pv[e].w; }
      }
#pragma unroll
      for (int e = 0; e < 9; ++e)
        if (tid + e * 256 < NG) {
          float4 t = v[e];
          if (gj[e] + 1 > gi[e]) t.y = 0.f;
          if (gj[e] + 2 > gi[e]) t.z = 0.f;
          if (gj[e] + 3 > gi[e]) t.w = 0.f;
          *reinterpret_cast<float4*>(s + gi[e] * DS + gj[e]) = t;
        }
    } else {
      for (int q = tid; q < NB * 32; q += 256) {
        const int i = q >> 5, j4 = (q & 31) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < nb && j4 <= i) {
          for (int t = 0; t < 4; ++t)
            if (j4 + t < nb) v[t] = a11[static_cast<long long>(i) * a.lda + j4 + t];
          for (int sp = 0; sp < a.splits; ++sp) {
            const float4 p = *reinterpret_cast<const float4*>(a.partials + sp * a.split_stride + static_cast<long long>(i) * NB + j4);
            v[0] -= p.x; v[1] -= p.y; v[2] -= p.z; v[3] -= p.w;
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (j4 + t > i) v[t] = 0.f;
        if (i >= nb && j4 <= i && i < j4 + 4) v[i - j4] = 1.0f;  // identity padding of a ragged last panel
        *reinterpret_cast<float4*>(s + i * DS + j4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    __syncthreads();
    PT3(1);
    float* l11 = a.L + static_cast<long long>(j0) * a.ldl + j0;
    __half* h11 = a.Lhi + static_cast<long long>(j0) * a.ldh + j0;
    __half* o11 = a.Llo + static_cast<long long>(j0) * a.ldh + j0;
    // rows [i_lo, i_hi) x 16-byte column groups [g_lo, g_hi) of the block in smem -> L (fp32 + fp16 pair), by the 128 threads
    // of warps 4..7; groups right of the diagonal block are written as zeros (upper triangle)
    auto store_rect = [&](int i_lo, int i_hi, int g_lo, int g_hi, int zero_from_col) {
      const int gw = g_hi - g_lo;
      for (int q = tid - 128; q < (i_hi - i_lo) * gw; q += 128) {
        const int i = i_lo + q / gw, j4 = 4 * (g_lo + q % gw);
        if (i >= nb) continue;
        float4 t = (j4 >= zero_from_col) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(s + i * DS + j4);
        if (full) {
          store_l4(l11 + static_cast<long long>(i) * a.ldl, h11 + static_cast<long long>(i) * a.ldh,
                   o11 + static_cast<long long>(i) * a.ldh, j4, t, sl);
        } else {
          const float tv[4] = {t.x, t.y, t.z, t.w};
          for (int u = 0; u < 4; ++u)
            if (j4 + u < nb) {
              l11[static_cast<long long>(i) * a.ldl + j4 + u] = tv[u];
              h3_split1(tv[u], sl, h11[static_cast<long long>(i) * a.ldh + j4 + u], o11[static_cast<long long>(i) * a.ldh + j4 + u]);
            }
        }
      }
    };
    for (int p = 0; p < NB / 32; ++p) {
      const int c0 = 32 * p;
      // ---- (1) 32 x 32 diagonal block: warp 0, row `lane` in registers; also leaves the block transposed in dT
      if (warp == 0) {
        float row[32];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 t = *reinterpret_cast<const float4*>(s + (c0 + lane) * DS + c0 + k);
          row[k] = t.x; row[k + 1] = t.y; row[k + 2] = t.z; row[k + 3] = t.w;
        }
        int isbad = 0;
        chol32_b8(row, lane, dinv + c0, isbad);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = (k <= lane) ? row[k] : 0.0f;
          row[k] = v;
          dT[k * DT + lane] = v;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(s + (c0 + lane) * DS + c0 + k) = make_float4(row[k], row[k + 1], row[k + 2], row[k + 3]);
        if (isbad && lane == 0) bad = 1;
      }
      __syncthreads();
      PT3(2 + 4 * p);
      // ---- (2) warps 1..3: rows below, x L_pp^T = a, one thread per row (right-looking substitution);
      //      warps 4..7: store the finished diagonal block (and the zeros to its right) and publish it
      if (warp >= 4) {
        store_rect(c0, c0 + 32, c0 / 4, NB / 4, c0 + 32);
        named_bar_sync(1, 128);
        if (tid == 255) {
          __threadfence();
          st_release_u32(a.ready, a.epoch_base + 2 * p + 1);
        }
      } else if (tid >= c0 + 32) {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 t = *reinterpret_cast<const float4*>(s + tid * DS + c0 + k);
          v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
        }
        row_solve32(v, dT, dinv + c0);
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(s + tid * DS + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
      PT3(3 + 4 * p);
      if (p == NB / 32 - 1) break;
      __syncthreads();
      PT3(4 + 4 * p);
      // ---- (3) warps 4..7 first store the solved rows of block-column p and publish them; then everyone applies the
      //      trailing update S[i][k] -= sum_c P[i][c] P[k][c] on the lower triangle of the (NB-c0-32)^2 block in
      //      interleaved 4x4 register tiles (neighbouring lanes read neighbouring rows: conflict-free float4 loads).
      //      Tile (ti, tj) holds rows ti + r*mt and columns tj + c*mt: entries with c > r are above the diagonal (never
      //      formed), c < r below it, c == r below or on it iff tj <= ti.
      if (warp >= 4) {
        store_rect(c0 + 32, NB, c0 / 4, c0 / 4 + 8, NB);
        named_bar_sync(1, 128);
        if (tid == 255) {
          __threadfence();
          st_release_u32(a.ready, a.epoch_base + 2 * p + 2);
        }
      }
      const int m0 = c0 + 32, mt = (NB - m0) / 4;  // mt = 24, 16, 8
      for (int t = tid; t < mt * mt; t += 256) {
        const int ti = t % mt, tj = t / mt;
        const float* Ar = s + (m0 + ti) * DS;
        const float* Br = s + (m0 + tj) * DS;
        float acc[4][4] = {};
#pragma unroll 2
        for (int k = c0; k < c0 + 32; k += 4) {
          float4 av[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) av[r] = *reinterpret_cast<const float4*>(Ar + r * mt * DS + k);
#pragma unroll
          for (int c = 0; c < 4; ++c) bv[c] = *reinterpret_cast<const float4*>(Br + c * mt * DS + k);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c)
              acc[r][c] += av[r].x * bv[c].x + av[r].y * bv[c].y + av[r].z * bv[c].z + av[r].w * bv[c].w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c)
            if (c < r || tj <= ti) s[(m0 + ti + r * mt) * DS + m0 + tj + c * mt] -= acc[r][c];
      }
      __syncthreads();
      PT3(5 + 4 * p);
    }
    __syncthreads();
    if (tid == 0 && bad) atomicOr(a.flag, 1);
    PT3(18);
    return;
  }

  // -------------------------------------------------------------------- TRSM rows (nb == NB whenever rows below exist)
  const int r0 = j0 + NB + (blockIdx.x - 1) * RPC;
  const int rows = min(RPC, n - r0);
  if (blockIdx.x == 1) PT3(20);
  {
    // this CTA's rows of A21 - sum of the update partials (four 16-byte groups per thread), while CTA 0 factors
    const float* a21 = a.A + static_cast<long long>(r0) * a.lda + j0;
    float4 v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
      v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < rows) v[e] = *reinterpret_cast<const float4*>(a21 + static_cast<long long>(i) * a.lda + j4);
    }
    for (int sp = 0; sp < a.splits; sp += 2) {
      float4 pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int q = tid + (e & 3) * 256, i = q >> 5, j4 = (q & 31) * 4;
        pv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows && sp + (e >> 2) < a.splits)
          pv[e] = *reinterpret_cast<const float4*>(a.partials + (sp + (e >> 2)) * a.split_stride +
                                                    static_cast<long long>(r0 - j0 + i) * NB + j4);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e & 3].x -= pv[e].x; v[e & 3].y -= pv[e].y; v[e & 3].z -= pv[e].z; v[e & 3].w -= pv[e].w;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
      *reinterpret_cast<float4*>(at + i * DS + j4) = v[e];
    }
  }
  if (blockIdx.x == 1) PT3(21);
  const float* l11 = a.L + static_cast<long long>(j0) * a.ldl + j0;
  constexpr int CPT = 32 / TPR;  // columns per thread in the block-update phase
  const int row = tid % RPC, part = tid / RPC;
  float* myrow = at + row * DS;
#pragma unroll 1
  for (int J = 0; J < NB / 32; ++J) {
    if (J > 0) {
      // block-update with the finished block-columns I < J of block-row J (final once step J-1's rows are published)
      wait_epoch(a.ready, a.epoch_base + 2 * J, j0);
      const int ncol4 = 8 * J;
      for (int q = tid; q < 32 * ncol4; q += 256) {
        const int i = 32 * J + q / ncol4, j4 = (q % ncol4) * 4;
        *reinterpret_cast<float4*>(s + i * DS + j4) = __ldcg(reinterpret_cast<const float4*>(l11 + static_cast<long long>(i) * a.ldl + j4));
      }
      __syncthreads();
      // v -= X_I L11[J][I]^T: thread (row, part) owns CPT of the 32 columns of block J
      float v[CPT];
      {
        const float4 t = *reinterpret_cast<const float4*>(myrow + 32 * J + CPT * part);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      }
#pragma unroll 1
      for (int I = 0; I < J; ++I) {
        float4 xp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xp[k] = *reinterpret_cast<const float4*>(myrow + 32 * I + 4 * k);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float4* lr = reinterpret_cast<const float4*>(s + (32 * J + CPT * part + j) * DS + 32 * I);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 l = lr[k];
            a0 += xp[k].x * l.x;
            a1 += xp[k].y * l.y;
            a2 += xp[k].z * l.z;
            a3 += xp[k].w * l.w;
          }
          v[j] -= (a0 + a1) + (a2 + a3);
        }
      }
      *reinterpret_cast<float4*>(myrow + 32 * J + CPT * part) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (blockIdx.x == 1 && J == 3) PT3(22);
    // diagonal block J (transposed) once CTA 0 has factored it
    wait_epoch(a.ready, a.epoch_base + 2 * J + 1, j0);
    {
      const int i = 32 * J + (tid >> 3), jj = (tid & 7) * 4;
      const float4 t = __ldcg(reinterpret_cast<const float4*>(l11 + static_cast<long long>(i) * a.ldl + 32 * J + jj));
      const int k = i - 32 * J;
      dT[(jj + 0) * DT + k] = t.x;
      dT[(jj + 1) * DT + k] = t.y;
      dT[(jj + 2) * DT + k] = t.z;
      dT[(jj + 3) * DT + k] = t.w;
      if (jj <= k && k < jj + 4) dinv[k] = 1.0f / (k == jj ? t.x : k == jj + 1 ? t.y : k == jj + 2 ? t.z : t.w);
    }
    __syncthreads();
    if (tid < RPC) {  // in-block forward substitution, one thread per row
      float v[32];
      float* r = at + tid * DS + 32 * J;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 t = *reinterpret_cast<const float4*>(r + k);
        v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
      }
      row_solve32(v, dT, dinv);
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(r + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
    __syncthreads();
  }
  if (blockIdx.x == 1) PT3(23);
  float* l21 = a.L + static_cast<long long>(r0) * a.ldl + j0;
  __half* h21 = a.Lhi + static_cast<long long>(r0) * a.ldh + j0;
  __half* o21 = a.Llo + static_cast<long long>(r0) * a.ldh + j0;
  for (int q = tid; q < RPC * 32; q += 256) {
    const int i = q >> 5, j4 = (q & 31) * 4;
    if (i < rows)
      store_l4(l21 + static_cast<long long>(i) * a.ldl, h21 + static_cast<long long>(i) * a.ldh,
               o21 + static_cast<long long>(i) * a.ldh, j4, *reinterpret_cast<const float4*>(at + i * DS + j4), sl);
  }
  if (blockIdx.x == 1) PT3(24);
}

constexpr int PANEL_SMEM = (NB * DS + RPC * DS + 32 * DT) * static_cast<int>(sizeof(float));

}  // namespace

size_t potrf_h3_workspace_bytes(int n) {
  // 256 bytes of scalars (epoch word) + split-K partials: at most MAX_SPLITS x rows x 128 with splits*tiles <= ~148+8
  const long long rows = n > 0 ? n : 1;
  long long worst = 0;
  for (long long j0 = NB; j0 < rows; j0 += NB) {
    const long long M = rows - j0, tiles = (M + NB - 1) / NB;
    long long S = 148 / tiles;
    if (S < 1) S = 1;
    if (S > MAX_SPLITS) S = MAX_SPLITS;
    if (S > j0 / H3_BK) S = j0 / H3_BK;
    if (S * M > worst) worst = S * M;
  }
  return 256 + static_cast<size_t>(worst) * NB * sizeof(float);
}

int potrf_h3(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, const gsmvi_h3_operand& Lh, int n,
             int* flag, void* workspace, int zero_upper) {
  if (n <= 0 || !A || !L || !flag || !workspace || !Lh.hi || !Lh.lo || !Lh.scale) return GSMVI_EINVAL;
  if ((lda & 3) != 0 || (ldl & 3) != 0 || (Lh.ld & 7) != 0 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(L)) & 15) != 0)
    return GSMVI_EALIGN;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(potrf_panel_h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  unsigned* ready = static_cast<unsigned*>(workspace);
  float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  __half* Lhi = static_cast<__half*>(Lh.hi);
  __half* Llo = static_cast<__half*>(Lh.lo);
  potrf_prepare_kernel<<<1, 256, 0, stream>>>(A, lda, n, Lh.scale, ready, flag);
  if (zero_upper && n > NB)
    potrf_zero_upper_kernel<<<dim3((n / 4 + 255) / 256, n), 256, 0, stream>>>(L, ldl, Lhi, Llo, Lh.ld, n);
  unsigned epoch = 0;
  const bool timing = getenv("GSMVI_POTRF_TIMING") != nullptr;
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = min(NB, n - j0);
    const int M = n - j0, rest = M - nb;
    PanelArgs pa;
    pa.A = A; pa.lda = lda; pa.L = L; pa.ldl = ldl; pa.Lhi = Lhi; pa.Llo = Llo; pa.ldh = Lh.ld; pa.scale_l = Lh.scale;
    pa.n = n; pa.j0 = j0; pa.nb = nb; pa.partials = partials; pa.splits = 0; pa.split_stride = static_cast<long long>(M) * NB;
    pa.flag = flag; pa.ready = ready; pa.epoch_base = epoch;
    pa.timing = (timing && j0 == NB * 8) ? 1 : 0;
    epoch += 8;
    if (j0 > 0) {
      const int tiles = (M + NB - 1) / NB;
      int S = 148 / tiles;
      if (S < 1) S = 1;
      if (S > MAX_SPLITS) S = MAX_SPLITS;
      if (S > j0 / H3_BK) S = j0 / H3_BK;
      pa.splits = S;
      HView va{Lhi + static_cast<long long>(j0) * Lh.ld, Llo + static_cast<long long>(j0) * Lh.ld, M, j0, Lh.ld, Lh.scale};
      HView vb{Lhi + static_cast<long long>(j0) * Lh.ld, Llo + static_cast<long long>(j0) * Lh.ld, nb, j0, Lh.ld, Lh.scale};
      H3Opts o;
      o.splits = S;
      o.split_stride = pa.split_stride;
      int rc = launch_gemm_h3(stream, M, nb, j0, va, vb, partials, NB, o);
      if (rc != GSMVI_OK) return rc;
    }
    potrf_panel_h3_kernel<<<1 + (rest + RPC - 1) / RPC, 256, PANEL_SMEM, stream>>>(pa);
  }
  if (timing && n > NB * 9) {
    long long h[32];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_pt3, sizeof(h));
    fprintf(stderr, "[potrf_h3 panel 8] CTA0: load %lld |", h[1] - h[0]);
    for (int p = 0; p < 4; ++p)
      fprintf(stderr, " p%d chol32 %lld trsm(t0) %lld sync %lld upd %lld |", p, h[2 + 4 * p] - (p ? h[1 + 4 * p] : h[1]),
              h[3 + 4 * p] - h[2 + 4 * p], p < 3 ? h[4 + 4 * p] - h[3 + 4 * p] : 0LL, p < 3 ? h[5 + 4 * p] - h[4 + 4 * p] : 0LL);
    fprintf(stderr, " total %lld || CTA1: prefetch %lld, stage-3 start at %lld, solve end %lld, store %lld (since CTA0 start)\n",
            h[18] - h[0], h[21] - h[20], h[22] - h[0], h[23] - h[0], h[24] - h[0]);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
