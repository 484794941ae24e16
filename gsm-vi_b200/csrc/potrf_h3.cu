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
// Look-ahead form (default for 512 <= n <= ~17k, GSMVI_POTRF_LOOKAHEAD=0 disables): ONE launch per panel.  The fused
// kernel's first CTAs run the panel program (2); its remaining CTAs run the (row tile, split) work items of the NEXT
// panel's update (1) restricted to the block-columns before the current panel - final at launch time, so neither role
// waits for the other - and the last of them that holds a partial of the next diagonal tile reduces all planes of that
// tile (fixed order) into a buffer the next launch's CTA 0 starts from.  The next panel subtracts the current panel's own
// K = 128 term itself: CTA 0 on its tensor core (TMA of the fp16 pair, 24 tcgen05 MMAs), the row owners in fp32 on the
// CUDA cores.  Inside CTA 0 the four-warp team that factors sits on the warps the scheduler serves first, the warps that
// solve the rows below (eight columns behind the team), publish and feed the trailing update sit underneath.
// Every launch runs each role once from a cold instruction cache, so what is compiled into the default path matters:
// polling loops and watchdogs are out of line, the default (LEAN) instantiation does not contain the alternative forms.
// See DESIGN.md section 3.2 and profiles/r02_potrf_h3_round2b.txt.
#include "dev_once.cuh"
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
// byte offset (from the 1 KiB-aligned operand tiles) of CTA 0's late-term tiles: behind the two tf32 operand tiles and
// the fp32 panel arrays, rounded up to 1 KiB
constexpr int LATE_TILE_OFF = 2 * 12288 + (((NB * DS + 2 * RPC * DS + 32 * DT) * 4 + 1023) / 1024) * 1024;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Barrier over the 256 threads that run the panel program.  Not __syncthreads(): in the fused kernel the CTA has 320
// threads (the GEMM role's shape) and the last two warps of a panel-role CTA have exited.
__device__ __forceinline__ void cta_sync() { named_bar_sync(0, 256); }

// fp32 -> tf32, round to nearest (ties away), on the integer ALU: the same result as cvt.rna.tf32.f32 for the finite values
// that occur here (measured on B200: cvt.rna.tf32 has a ~29-cycle latency, an IADD + LOP3 pair 8).
__device__ __forceinline__ float tf32_rna_alu(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// Cholesky of the 32 x 32 block at `blk` (shared memory, leading dimension DS, lower triangle valid, zeros above) by
// a team of four warps (128 threads; `warp` = 0..3 within the team), eight columns at a time:
//   warp 0: reads the 8 x 8 pivot block (broadcast loads), factors it in registers - software-pipelined so that the next
//           pivot's rsqrt is issued as soon as its entry is final and the rest of the column update fills its latency -
//           solves every row's eight entries against it and writes them back (also transposed into dT, 1/l_jj into dinv);
//   all four warps: rank-8 update of the remaining columns, a quarter of the columns each, rows below the diagonal only.
// A single warp retires ~1 instruction per 2 cycles (nothing hides its latencies), so the 24-column rank-8 update done
// by one warp with shuffles cost more than the eight dependent pivots; split over four warps through shared memory it
// is ~4x shorter.  A non-positive / non-finite pivot poisons its column; *bad is set if any diagonal entry is not finite > 0.
// follow_threads > 0: after group q is stored, warp 0 arrives on named barrier 12 + q, where the warps that solve the rows
// below the block wait for it (row_follow; follow_threads = 32 + 32 x their number) - they never hold the team up.
__device__ __forceinline__ void chol32_coop(float* __restrict__ blk, float* __restrict__ dinv, int* __restrict__ bad,
                                            const int warp, const int lane, const int follow_threads) {
  float mydiag = 1.0f;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int cq = 8 * q;
    if (warp == 0) {
      float a[8], d[8][8];
      {
        const float4 t0 = *reinterpret_cast<const float4*>(blk + lane * DS + cq);
        const float4 t1 = *reinterpret_cast<const float4*>(blk + lane * DS + cq + 4);
        a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 t0 = *reinterpret_cast<const float4*>(blk + (cq + r) * DS + cq);
        const float4 t1 = *reinterpret_cast<const float4*>(blk + (cq + r) * DS + cq + 4);
        d[r][0] = t0.x; d[r][1] = t0.y; d[r][2] = t0.z; d[r][3] = t0.w; d[r][4] = t1.x; d[r][5] = t1.y; d[r][6] = t1.z; d[r][7] = t1.w;
      }
      float rinv[8];
      // The chain from pivot to pivot runs on the raw rsqrt.approx value (relative error 2^-22.9): next diagonal entry =
      // d - (l * ra)^2, then straight into the next rsqrt.  The Newton step that brings 1 / l_jj to full fp32 accuracy is
      // taken off that chain (three dependent operations per pivot, 128 pivots per panel); everything that is stored -
      // the column of L, 1 / l_jj - uses the refined value.
      float ra;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(d[0][0]));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float djj = d[j][j];
        float rn = 0.0f;
        if (j < 7) {  // the entries the next pivot depends on first, and its rsqrt right behind them
          const float t = d[j + 1][j] * ra;
          const float dn = d[j + 1][j + 1] - t * t;
          d[j + 1][j + 1] = dn;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(dn));
        }
        const float r = ra * fmaf(-0.5f * djj, ra * ra, 1.5f);
        rinv[j] = r;
        if (j < 7) d[j + 1][j] *= r;
#pragma unroll
        for (int i = j + 2; i < 8; ++i) d[i][j] *= r;
#pragma unroll
        for (int i = j + 2; i < 8; ++i)
#pragma unroll
          for (int k = j + 1; k <= i; ++k) d[i][k] -= d[i][j] * d[k][j];
        ra = rn;
      }
      // this row against the pivot block, right-looking: x_j = a_j / l_jj, a_k -= x_j l_kj
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = a[j] * rinv[j];
        a[j] = x;
#pragma unroll
        for (int k = j + 1; k < 8; ++k) a[k] -= x * d[k][j];
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (cq + j <= lane) ? a[j] : 0.0f;
        if (lane == cq + j) {
          dinv[cq + j] = rinv[j];
          mydiag = a[j];
        }
      }
      *reinterpret_cast<float4*>(blk + lane * DS + cq) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(blk + lane * DS + cq + 4) = make_float4(o[4], o[5], o[6], o[7]);
      if (follow_threads > 0) asm volatile("bar.arrive %0, %1;" ::"r"(12 + q), "r"(follow_threads) : "memory");
    }
    if (q == 3) break;
    named_bar_sync(10, 128);
    {
      // rank-8 update of columns cq+8 .. 31: thread (row i = lane, column group = warp) takes columns cq+8+warp, +4, ...
      const float4 x0 = *reinterpret_cast<const float4*>(blk + lane * DS + cq);
      const float4 x1 = *reinterpret_cast<const float4*>(blk + lane * DS + cq + 4);
      // all of a thread's (at most six) columns at once: loads first, so that one shared-memory latency is exposed, not six
      float4 l0[6], l1[6];
      float cur[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const int k = cq + 8 + warp + 4 * m;
        if (k < 32 && k <= lane) {  // on or below the diagonal
          l0[m] = *reinterpret_cast<const float4*>(blk + k * DS + cq);
          l1[m] = *reinterpret_cast<const float4*>(blk + k * DS + cq + 4);
          cur[m] = blk[lane * DS + k];
        }
      }
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const int k = cq + 8 + warp + 4 * m;
        if (k < 32 && k <= lane) {
          float acc0 = cur[m], acc1 = 0.0f;
          acc0 -= x0.x * l0[m].x; acc1 -= x0.y * l0[m].y; acc0 -= x0.z * l0[m].z; acc1 -= x0.w * l0[m].w;
          acc0 -= x1.x * l1[m].x; acc1 -= x1.y * l1[m].y; acc0 -= x1.z * l1[m].z; acc1 -= x1.w * l1[m].w;
          blk[lane * DS + k] = acc0 + acc1;
        }
      }
      named_bar_sync(10, 128);
    }
  }
  if (warp == 0 && __any_sync(0xffffffffu, !(mydiag > 0.0f) || !(mydiag < 3.0e38f)) && lane == 0) *bad = 1;
}

// The rows below a 32 x 32 diagonal block, one thread per row (myrow = the row's 32 entries in the block's columns), solved
// eight columns behind chol32_coop: as soon as group q of the block is stored (named barrier 12 + q), x_j = v_j / l_jj
// against pivot block q, then v_k -= sum_j x_j l_kj for the row's remaining columns.  Run by warps that share their
// scheduler with (and yield to) the team: the triangular solve of the rows below hides behind the team's next eight pivots
// instead of being a phase of its own (as a separate one-thread-per-row substitution after the factorisation it cost
// 1.4-1.8k cycles per stage: ~700 instructions per thread at one issue every ~1.8 cycles).
__device__ __forceinline__ void row_follow(const float* __restrict__ blk, const float* __restrict__ dinv,
                                           float* __restrict__ myrow, const int follow_threads) {
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int cq = 8 * q;
    {
      named_bar_sync(12 + q, follow_threads);
      // this row against pivot block q (rows cq .. cq+7 of blk, final since the first barrier; zeros above its diagonal)
      float x[8];
      {
        const float4 t0 = *reinterpret_cast<const float4*>(myrow + cq);
        const float4 t1 = *reinterpret_cast<const float4*>(myrow + cq + 4);
        x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
      }
      float l[8][8], di[8];
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        const float4 t0 = *reinterpret_cast<const float4*>(blk + (cq + r) * DS + cq);
        l[r][0] = t0.x; l[r][1] = t0.y; l[r][2] = t0.z; l[r][3] = t0.w;
        if (r > 4) {
          const float4 t1 = *reinterpret_cast<const float4*>(blk + (cq + r) * DS + cq + 4);
          l[r][4] = t1.x; l[r][5] = t1.y; l[r][6] = t1.z; l[r][7] = t1.w;
        }
      }
      {
        const float4 t0 = *reinterpret_cast<const float4*>(dinv + cq);
        const float4 t1 = *reinterpret_cast<const float4*>(dinv + cq + 4);
        di[0] = t0.x; di[1] = t0.y; di[2] = t0.z; di[3] = t0.w; di[4] = t1.x; di[5] = t1.y; di[6] = t1.z; di[7] = t1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] *= di[j];
#pragma unroll
        for (int k = j + 1; k < 8; ++k) x[k] -= x[j] * l[k][j];
      }
      *reinterpret_cast<float4*>(myrow + cq) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(myrow + cq + 4) = make_float4(x[4], x[5], x[6], x[7]);
      // ... and the warp's 32 rows' remaining columns: V[32 x (24 - 8q)] -= X[32 x 8] L[cq+8.., cq..cq+7]^T on the warp-level
      // tensor path (mma.sync m16n8k8 tf32, hi*hi + lo*hi + hi*lo; fragments straight from shared memory - the row stride of
      // 132 words makes the (groupID, threadID_in_group) pattern conflict-free).  As one-thread-per-row FMAs this update
      // (192 FFMAs + 54 broadcast LDS.128 per thread at q = 0) made a column group of the rows slower than the team's eight
      // pivots, and the rows fell behind the factorisation they are meant to hide under.
      if (q < 3) {
        __syncwarp();
        const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
        float* rowbase = myrow - lane * DS;  // row 0 of this warp's 32 rows, first column of the block
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* xa = rowbase + (16 * mt) * DS + cq;
          const float av[4] = {xa[g * DS + t], xa[(g + 8) * DS + t], xa[g * DS + t + 4], xa[(g + 8) * DS + t + 4]};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h = tf32_rna_alu(av[e]);
            ah[mt][e] = __float_as_uint(h);
            al[mt][e] = __float_as_uint(tf32_rna_alu(av[e] - h));
          }
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          if (nt < 3 - q) {
            const float* lb = blk + (cq + 8 + 8 * nt + g) * DS + cq;  // row cq+8+8nt+g of the block: B[k][n] = l_{n,k}
            const float b0 = lb[t], b1 = lb[t + 4];
            const float b0h = tf32_rna_alu(b0), b1h = tf32_rna_alu(b1);
            const uint32_t bh[2] = {__float_as_uint(b0h), __float_as_uint(b1h)};
            const uint32_t bl[2] = {__float_as_uint(tf32_rna_alu(b0 - b0h)), __float_as_uint(tf32_rna_alu(b1 - b1h))};
            float acc[2][4], cor[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mt][e] = cor[mt][e] = 0.0f;
            ptx::mma_tf32_16x8x8(acc[0], ah[0], bh);
            ptx::mma_tf32_16x8x8(acc[1], ah[1], bh);
            ptx::mma_tf32_16x8x8(cor[0], al[0], bh);
            ptx::mma_tf32_16x8x8(cor[1], al[1], bh);
            ptx::mma_tf32_16x8x8(cor[0], ah[0], bl);
            ptx::mma_tf32_16x8x8(cor[1], ah[1], bl);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int hrow = 0; hrow < 2; ++hrow) {
                float2* dst = reinterpret_cast<float2*>(rowbase + (16 * mt + g + 8 * hrow) * DS + cq + 8 + 8 * nt + 2 * t);
                float2 o = *dst;
                o.x -= acc[mt][2 * hrow] + cor[mt][2 * hrow];
                o.y -= acc[mt][2 * hrow + 1] + cor[mt][2 * hrow + 1];
                *dst = o;
              }
          }
        }
        __syncwarp();
      }
    }
  }
}

// Forward substitution of one row against a 32 x 32 lower-triangular block, right-looking so that the dependent chain
// per column is one multiply and one FMA:  x_j = v_j / l_jj;  v_k -= x_j l_kj (k > j).  dT holds the block TRANSPOSED
// (dT[j * DT + k] = l_kj) so the column below the pivot is read as broadcast float4s.  Fully unrolled (a rolled variant
// with rotated registers measured 40% slower); the solved row replaces row_io[0..31].
// Inlined on purpose: every CTA runs this code once per launch from a cold instruction cache, and fall-through code is
// prefetched while the first call of an out-of-line copy measured 11-12k cycles (one full miss per 128-byte line).
template <int COPY>
__device__ __forceinline__ void row_solve32(float* __restrict__ row_io, const float* __restrict__ dT,
                                         const float* __restrict__ dinv) {
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row_io + k);
    v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
  }
  float* out_row = row_io;
  // software-pipelined: the pivot column j+1 (and its 1/l) is loaded while column j is applied, so the in-order warp
  // never waits on a shared-memory load inside the dependent chain
  float4 lc[8];
  float di = dinv[0];
#pragma unroll
  for (int g = 0; g < 8; ++g) lc[g] = *reinterpret_cast<const float4*>(dT + 4 * g);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float4 ln[8];
    float dn = 0.0f;
    if (j + 1 < 32) {
      dn = dinv[j + 1];
#pragma unroll
      for (int g = (j + 2) / 4; g < 8; ++g) ln[g] = *reinterpret_cast<const float4*>(dT + (j + 1) * DT + 4 * g);
    }
    const float x = v[j] * di;
    v[j] = x;
#pragma unroll
    for (int g = (j + 1) / 4; g < 8; ++g) {
      const int k = 4 * g;
      if (k + 0 > j) v[k + 0] -= x * lc[g].x;
      if (k + 1 > j) v[k + 1] -= x * lc[g].y;
      if (k + 2 > j) v[k + 2] -= x * lc[g].z;
      if (k + 3 > j) v[k + 3] -= x * lc[g].w;
    }
    di = dn;
#pragma unroll
    for (int g = (j + 2) / 4; g < 8; ++g) lc[g] = ln[g];
  }
#pragma unroll
  for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(out_row + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
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
    ready[0] = 0u;  // panel epochs: diagonal blocks (odd values, written by CTA 0's publisher warp)
    ready[1] = 0u;  // helper CTA count
    ready[2] = 0u;  // panel epochs: solved row slabs (even values, written by CTA 0's helper warps)
    ready[3] = 0u;  // count of update-GEMM CTAs that have stored a partial of the next diagonal tile
    ready[4] = 0u;  // row blocks stored / ready[5]: hosted update-GEMM CTAs finished (chained launches wait on these)
    ready[5] = 0u;
    ready[6] = 0u;  // row blocks 0..3 of each panel stored (what the next panel's CTA 0 waits for when chained)
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
__device__ long long g_pt3[64];
#define PT3(i) do { if (TIMING && threadIdx.x == 0 && a.j0 == 8 * NB) g_pt3[i] = clock64(); } while (0)
// per-panel wall-clock stamps (%globaltimer, ns) of CTA 0: [panel][0] = first instruction after the wait for the previous
// launch, [1] = diagonal block in shared memory (start phase done), [2] = last stage done
__device__ unsigned long long g_pt_panel[64][3];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PTP(i) do { if (TIMING && threadIdx.x == 128 && a.j0 / NB < 64) g_pt_panel[a.j0 / NB][i] = global_ns(); } while (0)
// CTA 0's stamps come from the first thread of the critical team (physical warp 4)
#define PT3C(i) do { if (TIMING && threadIdx.x == 128 && a.j0 == 8 * NB) g_pt3[i] = clock64(); } while (0)

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
  // helpers: the first `helpers` TRSM CTAs (0 or 16) each reduce 8 rows of A11 - sum P into d0 [128][128] before their own
  // rows and bump *helper_count; CTA 0 starts from d0 once the count reaches helper_target (16 CTAs' worth of loads in
  // flight instead of one's)
  float* d0;
  unsigned* helper_count;
  unsigned helper_target;
  int helpers;
  int trsm_ctas;  // CTAs 1 .. trsm_ctas own the rows below the diagonal block (the fused kernel's grid is larger)
  // look-ahead: `partials` holds the update with the panels before the previous one only (computed by GEMM-role CTAs of
  // the previous launch, concurrently with that panel); the previous panel's K = 128 contribution
  // L[rows, j0-128 : j0) L[j0 : j0+128, j0-128 : j0)^T is subtracted here, by the CTAs that own the rows
  int late;
  // the diagonal tile's share of that term on CTA 0's tensor core instead of the helpers' CUDA cores: the 128 x 128 block
  // L[j0 : j0+128, j0-128 : j0) arrives by TMA as the fp16 (hi, lo) pair that is already in memory, 24 tcgen05 MMAs
  // (kind::f16, hi*hi over two accumulators + the hi*lo / lo*hi correction) run while the helpers reduce A11 - sum P
  int late_mma;
  // with late_mma: `base` (when not null) = A11 - sum P of this panel's diagonal tile, [128][128], complete at launch; and
  // the update GEMM this launch hosts leaves the same for the NEXT panel in nb_out: the CTAs of its row tile 0 count
  // themselves in *nb_count after storing their partial, the one that brings it to nb_target adds the nb_splits partials
  // (in split order: deterministic) to the next diagonal tile of A.  No helper CTAs, no waiting at the start of a panel.
  const float* base;
  float* nb_out;
  const float* nb_A;        // A[nj0][nj0]
  const float* nb_partials; // the hosted GEMM's output planes (row 0 = row nj0)
  long long nb_stride;
  int nb_splits;
  unsigned* nb_count;
  unsigned nb_target;
  // chained launches (GSMVI_POTRF_CHAIN=1, panel k >= 1 of the look-ahead form): CTA 0 - the critical path - does not wait
  // for the previous grid to drain (griddepcontrol.wait = completion + flush of the whole grid) but for the two things it
  // reads: the first four row blocks of the previous panel stored (done[2], counted by their owners: the K = 128 term's
  // operand) and the reduced diagonal tile written (done[1], counted by the update GEMM's last CTA of row tile 0).  Both
  // counters only grow; the targets are cumulative.  Every other CTA keeps griddepcontrol.wait.  CTA 0 of launch k+1 can
  // only start once every CTA of launch k has run griddepcontrol.launch_dependents, i.e. is resident: nothing it waits
  // for can be without an SM.
  int chained;
  int early_scale;  // *scale_l may be read before the wait for the previous launch (every fused launch after the first:
                    // the scale was written before the first one's wait returned, and launches start in order)
  int track;  // count finished row blocks / GEMM CTAs in done[] (only needed when launches are chained)
  unsigned* done;
  unsigned gemm_target, first4_target;
};

// Chained launches only (GSMVI_POTRF_CHAIN=1): thread 0 of the CTA polls the previous launch's completion counters.  Kept out
// of line: the default path must not carry this code (the panel kernels run from a cold instruction cache every launch, and
// the same source with this loop inlined measured 0.83 ms per factorisation against 0.78 ms).
__device__ __noinline__ void poll_previous_launch(const unsigned* w0, unsigned t0w, const unsigned* w1, unsigned t1w, int j0) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_u32(w0) - t0w) < 0 || static_cast<int>(ld_acquire_u32(w1) - t1w) < 0) {
    __nanosleep(20);
    if (clock64() - t0 > 60000000000LL) { printf("gsmvi: potrf chained-launch watchdog (j0=%d block %d)\n", j0, blockIdx.x); __trap(); }
  }
}

// Start of a CTA's dependence on the previous launch (all 256 threads of the panel program).
__device__ __forceinline__ void wait_previous_launch(const PanelArgs& a) {
  if (!a.chained || blockIdx.x != 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    return;
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) poll_previous_launch(a.done + 2, a.first4_target, a.done + 1, a.gemm_target, a.j0);
  cta_sync();
  // what follows reads the previous launch's results through L2 (ld.cg) and through TMA (async proxy)
  asm volatile("fence.proxy.async;" ::: "memory");
}

// The waits of the panel program keep their polling loops and watchdogs OUT OF LINE: every launch runs this code once
// from a cold instruction cache, and code that is merely skipped still costs fetches (see poll_previous_launch).
__device__ __noinline__ void poll_epoch(const unsigned* ready, unsigned target, int j0) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_u32(ready) - target) < 0) {
    __nanosleep(20);
    if (clock64() - t0 > 60000000000LL) { printf("gsmvi: potrf panel watchdog (j0=%d epoch %u)\n", j0, target); __trap(); }
  }
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { ptx::mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (!ptx::mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// spin (thread 0) until the panel's epoch word reaches `target`, then release the whole CTA
__device__ __forceinline__ void wait_epoch(const unsigned* ready, unsigned target, int j0) {
  if (threadIdx.x == 0 && static_cast<int>(ld_acquire_u32(ready) - target) < 0) poll_epoch(ready, target, j0);
  cta_sync();
}

// Epochs of one panel (relative to epoch_base): 2p+1 = diagonal block p factored and stored (word ready[0]), 2p+2 = the rows
// below it in block-column p solved and stored (word ready[2]: two writers, so two words - each stays monotone).  A TRSM CTA needs 2J for the block-update of its stage J and 2J+1 for the in-block
// substitution, so only that last substitution trails CTA 0.
// FULL: nb == 128 (every panel but a ragged last one, which has no rows below it and runs as a single CTA).
// The straight-line parts are kept small on purpose: each CTA runs this code once per launch with a cold instruction
// cache, and an earlier fully unrolled version (10.8k SASS instructions) spent more time fetching than computing.
// LEAN: the look-ahead form with the late term on CTA 0's tensor core (the default): no helper CTAs and no split-K partials
// for CTA 0 - the code of those paths is not compiled into the kernel at all.
template <bool FULL, bool TIMING, bool LEAN>
__device__ __forceinline__ void panel_body(const PanelArgs& a, uint8_t* sm_raw, const CUtensorMap* tmLhi,
                                           const CUtensorMap* tmLlo) {
  // [1 KiB-aligned: tf32 hi | lo operand tiles of CTA 0's tensor-core update, 96 x 128 B each] then the fp32 arrays.  The
  // MMA (M = 128) reads 32 rows past each 96-row tile: those land in the lo tile / in s (ignored accumulator rows).
  const uint32_t pt_addr = (ptx::smem_u32(sm_raw) + 1023u) & ~1023u;
  float* sm = reinterpret_cast<float*>(sm_raw + (pt_addr - ptx::smem_u32(sm_raw)) + 2 * 12288);
  float* s = sm;                   // [NB][DS]   diagonal block (CTA 0) / published block-rows of L11 (TRSM CTAs)
  float* at = sm + NB * DS;        // [RPC][DS]  TRSM CTAs: their rows of the panel
  float* dT = at + RPC * DS;       // [32][DT]   current 32 x 32 diagonal block, transposed
  float* la = dT + 32 * DT;        // [RPC][DS]  TRSM CTAs, look-ahead: their rows of the previous panel's block-column
  __shared__ __align__(16) float dinv[32];  // read as float4 by row_follow
  __shared__ int bad;
  __shared__ __align__(8) unsigned long long mma_bar;
  __shared__ __align__(8) unsigned long long late_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = a.j0, nb = a.nb, n = a.n;
  float sl = 0.0f;  // scale of the fp16 split of L (written by the prepare kernel: read after the wait)
  if (tid == 0) bad = 0;
  // programmatic dependent launch: the launch itself overlapped the tail of the previous kernel in the stream; everything
  // that reads its results comes after griddepcontrol.wait (CTA 0 first does the set-up that needs none of them)
  if (a.early_scale) sl = *a.scale_l;  // ~700 cycles of L2 latency that would otherwise follow the wait, every panel
  if (blockIdx.x != 0) {
    wait_previous_launch(a);
    if (!a.early_scale) sl = *a.scale_l;
  }

  if (blockIdx.x == 0) {
    // ------------------------------------------------------------------ diagonal block
    const uint32_t bar_addr = ptx::smem_u32(&mma_bar);
    const bool late_mma = FULL && tmLhi != nullptr && a.late_mma != 0;  // (the two-launch kernel passes no tensor maps)
    const uint32_t tmem_cols = late_mma ? 256u : 128u;
    if (warp == 0) {  // tensor-core trailing update below: 128 TMEM columns and one mbarrier (late term: 128 columns more)
      ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), tmem_cols);
      ptx::tmem_relinquish();
      if (lane == 0) {
        ptx::mbar_init(bar_addr, 1);
        ptx::mbar_init(ptx::smem_u32(&late_bar), 1);
        ptx::fence_mbar_init();
        if (late_mma) {
          ptx::prefetch_tmap(tmLhi);
          ptx::prefetch_tmap(tmLlo);
        }
      }
      __syncwarp();
    }
    wait_previous_launch(a);
    if (!a.early_scale) sl = *a.scale_l;
    PT3C(0);
    PTP(0);
    if (warp == 0) {
      if (late_mma) {
        // the previous panel's block of these 128 rows, fp16 pair: four 128 x 64 boxes (hi | lo, two K halves), SWIZZLE_128B
        const uint32_t lt_addr = pt_addr + LATE_TILE_OFF, lb = ptx::smem_u32(&late_bar);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(lb, 4 * H3_TILE_BYTES);
          ptx::tma_load_2d(lt_addr, tmLhi, lb, j0 - NB, j0);
          ptx::tma_load_2d(lt_addr + H3_TILE_BYTES, tmLhi, lb, j0 - NB + H3_BK, j0);
          ptx::tma_load_2d(lt_addr + 2 * H3_TILE_BYTES, tmLlo, lb, j0 - NB, j0);
          ptx::tma_load_2d(lt_addr + 3 * H3_TILE_BYTES, tmLlo, lb, j0 - NB + H3_BK, j0);
        }
        __syncwarp();
      }
    }
    // late term on the tensor core (see PanelArgs::late_mma): issued by physical warp 0 once the TMA boxes have landed -
    // call it after this thread's global loads have been issued, so that they are in flight during the wait
    auto issue_late_mma = [&]() {
      if (late_mma && __shfl_sync(0xffffffffu, warp, 0) == 0) {
        // main accumulator at TMEM column 0, the correction products at 128; the drain below forms
        // (main + 2^-11 corr) / scale^2
        ptx::tc_fence_before_sync();
        __syncwarp();
        ptx::tc_fence_after_sync();
        const uint32_t tb = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(&tmem_slot), 0);
        const uint32_t lt_addr = pt_addr + LATE_TILE_OFF;
        mbar_wait_lean(ptx::smem_u32(&late_bar), 0);
        ptx::tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_f16(false, false);
        if (ptx::elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < H3_BK / H3_UMMA_K; ++kk) {
              const uint64_t dh = make_smem_desc(lt_addr + h * H3_TILE_BYTES + kk * (H3_UMMA_K * 2), 16, 1024, 2);
              const uint64_t dl = make_smem_desc(lt_addr + (2 + h) * H3_TILE_BYTES + kk * (H3_UMMA_K * 2), 16, 1024, 2);
              ptx::umma_f16(tb + 128, dl, dh, idesc, (h > 0 || kk > 0) ? 1u : 0u);
              ptx::umma_f16(tb + 128, dh, dl, idesc, 1u);
              ptx::umma_f16(tb, dh, dh, idesc, (h > 0 || kk > 0) ? 1u : 0u);
            }
          ptx::umma_commit(bar_addr);
        }
        __syncwarp();
      }
    };
    const float* a11 = a.A + static_cast<long long>(j0) * a.lda + j0;
    PT3C(50);
    if (FULL) {
      // 16-byte groups of the lower triangle: thread = (column group tid & 31, rows (tid >> 5) + 8 e); the groups above the
      // diagonal are skipped (an earlier compact enumeration of the 2112 groups cost nine sqrtf-based index computations
      // per thread: 2.5k cycles before the first load was issued)
      constexpr int PER = 16;
      float4 v[PER];
      int off_s[PER];  // (row << 8) | first column; -1: no group
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        const int gi = (tid >> 5) + 8 * e, gj = 4 * (tid & 31);
        off_s[e] = (gj <= gi) ? ((gi << 8) | gj) : -1;
      }
      if (LEAN || a.base != nullptr) {
        // A11 - sum P was formed by the previous launch (the last of the update GEMM's CTAs that hold a partial of this tile);
        // LEAN: the first two panels, which have no such term, read A11 itself through the same code
        const float* src = a.base != nullptr ? a.base : a11;
        const long long ld = a.base != nullptr ? NB : a.lda;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off_s[e] >= 0) v[e] = __ldcg(reinterpret_cast<const float4*>(src + (off_s[e] >> 8) * ld + (off_s[e] & 255)));
        }
        PT3C(51);
        issue_late_mma();
        PT3C(52);
      } else if (a.helpers > 0) {
        issue_late_mma();
        // the helper CTAs have formed A11 - sum P in d0: wait for all of them, then one round of loads
        if (tid == 0) {
          const long long t0 = clock64();
          while (static_cast<int>(ld_acquire_u32(a.helper_count) - a.helper_target) < 0) {
            if (clock64() - t0 > 60000000000LL) { printf("gsmvi: potrf helper watchdog (j0=%d)\n", j0); __trap(); }
          }
        }
        cta_sync();
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off_s[e] >= 0) v[e] = __ldcg(reinterpret_cast<const float4*>(a.d0 + (off_s[e] >> 8) * NB + (off_s[e] & 255)));
        }
      } else {
#pragma unroll
        for (int e = 0; e < PER; ++e) {
          v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off_s[e] >= 0) v[e] = *reinterpret_cast<const float4*>(a11 + static_cast<long long>(off_s[e] >> 8) * a.lda + (off_s[e] & 255));
        }
        issue_late_mma();
#pragma unroll 1
        for (int sp = 0; sp < a.splits; ++sp) {
          const float* pb = a.partials + sp * a.split_stride;
          float4 pv[PER];
#pragma unroll
          for (int e = 0; e < PER; ++e) {
            pv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off_s[e] >= 0) pv[e] = __ldcg(reinterpret_cast<const float4*>(pb + (off_s[e] >> 8) * NB + (off_s[e] & 255)));
          }
#pragma unroll
          for (int e = 0; e < PER; ++e) { v[e].x -= pv[e].x; v[e].y -= pv[e].y; v[e].z -= pv[e].z; v[e].w -= pv[e].w; }
        }
      }
#pragma unroll
      for (int e = 0; e < PER; ++e)
        if (off_s[e] >= 0) {
          const int gi = off_s[e] >> 8, gj = off_s[e] & 255;
          float4 t = v[e];
          if (gj + 1 > gi) t.y = 0.f;
          if (gj + 2 > gi) t.z = 0.f;
          if (gj + 3 > gi) t.w = 0.f;
          *reinterpret_cast<float4*>(s + gi * DS + gj) = t;
        }
    } else {
#pragma unroll 1
      for (int q = tid; q < NB * 32; q += 256) {
        const int i = q >> 5, j4 = (q & 31) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = j4 + t;
          if (i < nb && j <= i) {
            float x = a11[static_cast<long long>(i) * a.lda + j];
            for (int sp = 0; sp < a.splits; ++sp) x -= a.partials[sp * a.split_stride + static_cast<long long>(i) * NB + j];
            v[t] = x;
          } else if (i >= nb && j == i) {
            v[t] = 1.0f;  // identity padding of the ragged last panel
          }
        }
        *reinterpret_cast<float4*>(s + i * DS + j4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    ptx::tc_fence_before_sync();
    cta_sync();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    PT3C(53);
    if (late_mma) {
      // s -= L_prev L_prev^T (lower triangle): warp w drains TMEM lane quadrant w % 4, every other 32-column chunk
      mbar_wait_lean(bar_addr, 0);
      ptx::tc_fence_after_sync();
      PT3C(54);
      const int qd = warp & 3, r = 32 * qd + lane;
      const float c1 = 1.0f / (sl * sl), c2 = c1 * (1.0f / H3_LO_SCALE);
#pragma unroll 1
      for (int col0 = 32 * (warp >> 2); col0 <= 32 * qd; col0 += 64) {
        uint32_t mn[32], cr[32];
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(32 * qd) << 16) + col0;
        ptx::tmem_ld_32x32(ta, mn);
        ptx::tmem_ld_32x32(ta + 128, cr);
        ptx::tmem_ld_wait();
        float* dst = s + r * DS + col0;
#pragma unroll
        for (int u = 0; u < 32; u += 4) {
          if (col0 + u <= r) {
            float4 o = *reinterpret_cast<float4*>(dst + u);
            float w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = fmaf(__uint_as_float(cr[u + e]), c2, __uint_as_float(mn[u + e]) * c1);
            o.x -= w[0];
            if (col0 + u + 1 <= r) o.y -= w[1];
            if (col0 + u + 2 <= r) o.z -= w[2];
            if (col0 + u + 3 <= r) o.w -= w[3];
            *reinterpret_cast<float4*>(dst + u) = o;
          }
        }
      }
      ptx::tc_fence_before_sync();
      cta_sync();
      ptx::tc_fence_after_sync();
    }
    const uint32_t late_par = late_mma ? 1u : 0u;  // the late-term commit was phase 0 of the MMA barrier
    PT3C(1);
    PTP(1);
    float* l11 = a.L + static_cast<long long>(j0) * a.ldl + j0;
    __half* h11 = a.Lhi + static_cast<long long>(j0) * a.ldh + j0;
    __half* o11 = a.Llo + static_cast<long long>(j0) * a.ldh + j0;
    // one 16-byte group (row i, columns j4 .. j4+3 of the block) -> L (fp32 + fp16 pair)
    auto store_group = [&](int i, int j4, const float4 t) {
      if (FULL) {
        store_l4(l11 + static_cast<long long>(i) * a.ldl, h11 + static_cast<long long>(i) * a.ldh,
                 o11 + static_cast<long long>(i) * a.ldh, j4, t, sl);
      } else if (i < nb) {
        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j4 + u < nb) {
            l11[static_cast<long long>(i) * a.ldl + j4 + u] = tv[u];
            h3_split1(tv[u], sl, h11[static_cast<long long>(i) * a.ldh + j4 + u], o11[static_cast<long long>(i) * a.ldh + j4 + u]);
          }
      }
    };
    // Warp roles from here on.  The warp scheduler of an SM sub-partition serves its resident warps highest warp id first,
    // and the pairs (w, w + 4) share a sub-partition: the critical team - the 32 x 32 factorisations, the row solves and the
    // next-diagonal-block update - therefore runs on the physical warps 4..7 (roles 0..3), the work that only has to keep
    // up on the physical warps 0..3 underneath it: role 4 = publisher (finished parts of L11 -> global memory, epochs), roles
    // 5..7 = helpers (the rows below the diagonal block, solved eight columns behind the team; tensor-core trailing update).
    // The publisher sits under the team's lead warp, whose dependent chain leaves most issue slots free.  (With the
    // publisher on physical warp 7 its ~60-instruction store iterations out-prioritised the team's fourth warp, and every
    // team barrier waited for that warp: the stage cost ~9.2k cycles whatever the update between two factorisations was.)
    // Barrier 2 = the seven computing warps (224 threads); barrier 3 + p = "stage p done" (+ the publisher).
    constexpr int NCOMP = 224;
    const int role = (__shfl_sync(0xffffffffu, warp, 0) + 4) & 7;
    const int rtid = (role << 5) | lane;
    if (role == 4) {
      {
        // while the first block is being factored: zeros in the 32 x 32 blocks above the block diagonal of L11
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int q = lane; q < 6 * 256; q += 32) {
          const int b = q >> 8, e = q & 255;                    // block b = (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
          const int bi = b < 3 ? 0 : (b < 5 ? 1 : 2), bj = b < 3 ? b + 1 : (b < 5 ? b - 1 : 3);
          store_group(32 * bi + (e >> 3), 32 * bj + 4 * (e & 7), z4);
        }
      }
#pragma unroll 1
      for (int p = 0; p < NB / 32; ++p) {
        const int c0 = 32 * p;
        named_bar_sync(3 + p, 256);  // diagonal block p factored
        // the 32 x 32 block (zeros above its diagonal come from s): lane = (row mod 4, group)
#pragma unroll 2
        for (int i = c0 + (lane >> 3); i < c0 + 32; i += 4) {
          const int j4 = c0 + 4 * (lane & 7);
          store_group(i, j4, *reinterpret_cast<const float4*>(s + i * DS + j4));
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release_u32(a.ready, a.epoch_base + 2 * p + 1);
        }
        // the solved rows below it in block-column p (final at the same barrier)
        if (p == NB / 32 - 1) break;
#pragma unroll 2
        for (int i = c0 + 32 + (lane >> 3); i < NB; i += 4) {
          const int j4 = c0 + 4 * (lane & 7);
          store_group(i, j4, *reinterpret_cast<const float4*>(s + i * DS + j4));
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release_u32(a.ready + 2, a.epoch_base + 2 * p + 2);
        }
      }
    } else {
      // Stage p (block-column c0 = 32 p) - the chain from one 32 x 32 Cholesky to the next is kept as short as the
      // arithmetic allows: after block p is factored, the rows below are solved (one thread per row), then the team
      // applies the solved rows of block-row p+1 to the NEXT diagonal block only (warp-level tensor products, 32 x 32 x 32)
      // and goes straight on to factor it, while the helpers publish the solved rows, put the rest of the trailing update
      // on the tensor core (tf32 hi/lo split of the solved slab -> swizzled operand tiles -> 12 MMAs into TMEM) and drain
      // it into s during that next factorisation.  The drained rows (below block-row p+1) are first read by the next
      // stage's row solves, after the barrier that follows the factorisation.
#pragma unroll 1
      for (int p = 0; p < NB / 32; ++p) {
        const int c0 = 32 * p;
        // ---- (1) 32 x 32 diagonal block and the rows below it by the team | roles 5, 6: drain of the
        //      previous stage's tensor update (accumulator quadrant 0 is the block the fast path already updated)
        const int nrw = NB / 32 - 1 - p;  // warps that solve the rows below this block: roles 5..7 (p = 0), 5..6, 5
        if (role < 4) {
          chol32_coop(s + c0 * DS + c0, dinv, &bad, role, lane, nrw > 0 ? 32 + 32 * nrw : 0);
        } else {
          // block-row b of the tile: the one whose rows sit in this warp's quadrant (role - 4 = physical warp id % 4) of the
          // previous stage's accumulator, so that the warp that drains a row is the one that goes on to solve it
          const int b = p + role - 4;
          if (b < NB / 32) {
            if (p >= 1) {
              const int qd = role - 4;
              mbar_wait_lean(bar_addr, (static_cast<uint32_t>(p - 1) + late_par) & 1u);
              ptx::tc_fence_after_sync();
              const int r = 32 * qd + lane;  // accumulator row = row c0 + r of the block
#pragma unroll 1
              for (int col0 = 0; col0 <= 32 * qd + 16; col0 += 16) {  // chunks of 16 columns up to the diagonal
                uint32_t t[16];
                ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(32 * qd) << 16) + col0, t);
                ptx::tmem_ld_wait();
                float* dst = s + (c0 + r) * DS + c0 + col0;
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                  if (col0 + u <= r) {
                    float4 o = *reinterpret_cast<float4*>(dst + u);
                    o.x -= __uint_as_float(t[u]);
                    if (col0 + u + 1 <= r) o.y -= __uint_as_float(t[u + 1]);
                    if (col0 + u + 2 <= r) o.z -= __uint_as_float(t[u + 2]);
                    if (col0 + u + 3 <= r) o.w -= __uint_as_float(t[u + 3]);
                    *reinterpret_cast<float4*>(dst + u) = o;
                  }
                }
              }
              ptx::tc_fence_before_sync();
            }
            row_follow(s + c0 * DS + c0, dinv, s + (32 * b + lane) * DS + c0, 32 + 32 * nrw);
          }
        }
        // ---- (2) the rows below were solved inside chol32_coop, eight columns behind the factorisation
        named_bar_sync(2, NCOMP);
        asm volatile("bar.arrive %0, %1;" ::"r"(3 + p), "r"(256) : "memory");
        PT3C(2 + 4 * p);
        if (p == NB / 32 - 1) break;
        PT3C(3 + 4 * p);
        PT3C(4 + 4 * p);
        const int m0 = c0 + 32, R = NB - m0;  // R = 96, 64, 32 rows (and columns) left
        if (role < 4) {
          // ---- (3a) next diagonal block: S[m0+i][m0+k] -= sum_c X[i][c] X[k][c], X = the solved rows m0 .. m0+31 of
          //      block-column p, on the warp-level tensor path (mma.sync m16n8k8 tf32, hi*hi + lo*hi + hi*lo): fragments are
          //      read straight from s (row stride 132 words: the (groupID, threadID_in_group) pattern is conflict-free),
          //      no staging, no barrier.
          //      Role 0: 16 x 16 quadrant (0,0); role 2: (1,1); roles 1, 3: the two 16 x 8 halves of (1,0).
          const int g = lane >> 2, t = lane & 3;
          const int qi = role == 0 ? 0 : 1, qj = role == 2 ? 1 : 0;
          const int nt_lo = role == 3 ? 1 : 0, nt_hi = role == 1 ? 1 : 2;
          const float* xa = s + (m0 + 16 * qi) * DS + c0;
          const float* xb = s + (m0 + 16 * qj) * DS + c0;
          // hi*hi and the two correction products go to separate accumulators: four independent MMA chains per warp (a
          // dependent mma.sync costs ~20 cycles, an independent one ~10) and the small terms are summed among themselves
          float acc[2][4], cor[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = cor[nt][e] = 0.0f;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float av[4], bv[2][2];
            av[0] = xa[g * DS + 8 * kk + t];
            av[1] = xa[(g + 8) * DS + 8 * kk + t];
            av[2] = xa[g * DS + 8 * kk + t + 4];
            av[3] = xa[(g + 8) * DS + 8 * kk + t + 4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              bv[nt][0] = xb[(8 * nt + g) * DS + 8 * kk + t];
              bv[nt][1] = xb[(8 * nt + g) * DS + 8 * kk + t + 4];
            }
            uint32_t ah[4], al[4], bh[2][2], bl[2][2];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float h = tf32_rna_alu(av[e]);
              ah[e] = __float_as_uint(h);
              al[e] = __float_as_uint(tf32_rna_alu(av[e] - h));
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float h = tf32_rna_alu(bv[nt][e]);
                bh[nt][e] = __float_as_uint(h);
                bl[nt][e] = __float_as_uint(tf32_rna_alu(bv[nt][e] - h));
              }
            if (nt_lo == 0) ptx::mma_tf32_16x8x8(acc[0], ah, bh[0]);
            if (nt_hi == 2) ptx::mma_tf32_16x8x8(acc[1], ah, bh[1]);
            if (nt_lo == 0) ptx::mma_tf32_16x8x8(cor[0], al, bh[0]);
            if (nt_hi == 2) ptx::mma_tf32_16x8x8(cor[1], al, bh[1]);
            if (nt_lo == 0) ptx::mma_tf32_16x8x8(cor[0], ah, bl[0]);
            if (nt_hi == 2) ptx::mma_tf32_16x8x8(cor[1], ah, bl[1]);
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] += cor[nt][e];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (nt >= nt_lo && nt < nt_hi) {
#pragma unroll
              for (int hrow = 0; hrow < 2; ++hrow) {
                const int i = 16 * qi + g + 8 * hrow, k = 16 * qj + 8 * nt + 2 * t;  // entry (i, k), (i, k + 1) of the block
                if (k <= i) {
                  float2* dst = reinterpret_cast<float2*>(s + (m0 + i) * DS + m0 + k);
                  float2 o = *dst;
                  o.x -= acc[nt][2 * hrow];
                  if (k + 1 <= i) o.y -= acc[nt][2 * hrow + 1];
                  *dst = o;
                }
              }
            }
          }
          named_bar_sync(10, 128);
        } else {
          // ---- (3b) helpers: while more than the next diagonal block is left, put the trailing update of the R x R block on
          //      the tensor core: S[i][k] -= sum_c P[i][c] P[k][c] with P = the solved slab (rows m0.., columns c0..c0+31), as
          //      three tf32 MMAs (hi*hi + lo*hi + hi*lo: fp32-grade products, K = 32) into TMEM - the helpers split P into the
          //      swizzled K-major operand tiles, one thread issues, the drain is step (1) of the next stage.
          if (R > 32) {
#pragma unroll 1
            for (int q = rtid - 160; q < R * 8; q += 96) {
              const int r = q >> 3, ch = q & 7;
              const float4 v = *reinterpret_cast<const float4*>(s + (m0 + r) * DS + c0 + 4 * ch);
              float4 hi, lo;
              hi.x = tf32_rna_alu(v.x); lo.x = tf32_rna_alu(v.x - hi.x);
              hi.y = tf32_rna_alu(v.y); lo.y = tf32_rna_alu(v.y - hi.y);
              hi.z = tf32_rna_alu(v.z); lo.z = tf32_rna_alu(v.z - hi.z);
              hi.w = tf32_rna_alu(v.w); lo.w = tf32_rna_alu(v.w - hi.w);
              const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4);  // SWIZZLE_128B
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pt_addr + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pt_addr + 12288 + off), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
            }
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before_sync();
            named_bar_sync(11, 96);
            if (role == 5) {
              // the warp stays converged and one elected lane issues: addresses and descriptors are then warp-uniform to the
              // compiler (uniform registers) instead of going through an ELECT + R2UR loop per MMA (ptx::elect_one)
              ptx::tc_fence_after_sync();
              // c_format F32, a/b TF32, both K-major, N = R, M = 128 (rows beyond R hold stale data: their outputs are ignored)
              const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(R >> 3) << 17) | (8u << 24);
              const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
              if (ptx::elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t dh = make_smem_desc(pt_addr + kk * 32, 16, 1024, 2);
                  const uint64_t dl = make_smem_desc(pt_addr + 12288 + kk * 32, 16, 1024, 2);
                  ptx::umma_tf32(tb, dh, dh, idesc, kk > 0 ? 1u : 0u);
                  ptx::umma_tf32(tb, dl, dh, idesc, 1u);
                  ptx::umma_tf32(tb, dh, dl, idesc, 1u);
                }
                ptx::umma_commit(bar_addr);
              }
              __syncwarp();
            }
          }
        }
        PT3C(5 + 4 * p);
      }
    }
    cta_sync();
    if (warp == 0) {
      ptx::tc_fence_after_sync();
      ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
    if (tid == 0 && bad) atomicOr(a.flag, 1);
    PT3C(18);
    PTP(2);
    return;
  }

  if (!FULL) return;
  // -------------------------------------------------------------------- TRSM rows (only full panels have rows below)
  if (blockIdx.x == 1) PT3(20);
  bool lp_staged = false;  // s holds the previous panel's rows L[j0 .. j0+127][j0-128 .. j0) (look-ahead operand)
  // the previous panel's block-column of the 128 rows of this panel's diagonal block -> s (64 KiB, L2-resident)
  auto stage_lp = [&]() {
    const float* lp = a.L + static_cast<long long>(j0) * a.ldl + (j0 - NB);
    float4 t[16];  // all sixteen loads of a thread in flight: one L2 round trip instead of four
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int q = tid + e * 256;
      t[e] = __ldcg(reinterpret_cast<const float4*>(lp + static_cast<long long>(q >> 5) * a.ldl + (q & 31) * 4));
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int q = tid + e * 256;
      *reinterpret_cast<float4*>(s + (q >> 5) * DS + (q & 31) * 4) = t[e];
    }
  };
  if (!LEAN && static_cast<int>(blockIdx.x) <= a.helpers) {
    // helper: rows 8 (blockIdx.x - 1) .. +7 of the diagonal block, A11 - sum P (- the look-ahead term): thread = one
    // column x four rows, so that a warp reads 32 consecutive rows of s (conflict-free float4) and broadcasts its own rows
    const int col = tid & 127, ih = 8 * (blockIdx.x - 1) + 4 * (tid >> 7);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // A11 and its look-ahead partials first: their latency hides behind the staging and the FMA loop
    float hv[4], hp[4][MAX_SPLITS];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ih + r;
      hv[r] = (col <= i) ? a.A[static_cast<long long>(j0 + i) * a.lda + j0 + col] : 0.0f;
#pragma unroll
      for (int sp = 0; sp < MAX_SPLITS; ++sp)
        hp[r][sp] = (col <= i && sp < a.splits) ? __ldcg(a.partials + sp * a.split_stride + static_cast<long long>(i) * NB + col) : 0.0f;
    }
    if (a.late && !a.late_mma) {
      stage_lp();
      lp_staged = true;
      cta_sync();
      if (blockIdx.x == 1) PT3(30);
      if (col <= ih + 3) {
#pragma unroll 4
        for (int k = 0; k < NB; k += 4) {
          const float4 b = *reinterpret_cast<const float4*>(s + col * DS + k);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(s + (ih + r) * DS + k);
            acc[r] = fmaf(x.x, b.x, fmaf(x.y, b.y, fmaf(x.z, b.z, fmaf(x.w, b.w, acc[r]))));
          }
        }
      }
    }
    if (blockIdx.x == 1) PT3(31);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ih + r;
      if (col <= i) {
        float v = hv[r];
#pragma unroll
        for (int sp = 0; sp < MAX_SPLITS; ++sp) v -= hp[r][sp];
        a.d0[i * NB + col] = v - acc[r];
      }
    }
    cta_sync();
    if (tid == 0) {
      __threadfence();
      atomicAdd(a.helper_count, 1u);
    }
    if (blockIdx.x == 1) PT3(32);
  }
  if (blockIdx.x == 17) PT3(40);
  // one CTA per SM (the spin-waits need every CTA resident): a CTA takes row blocks blockIdx.x - 1, + gridDim.x - 1, ...
  // of 32 rows; from its second block on the epochs it waits for have already been published
  const int nblocks = (n - j0 - NB + RPC - 1) / RPC;
#pragma unroll 1
  for (int rb = blockIdx.x - 1; rb < nblocks; rb += a.trsm_ctas) {
  const int r0 = j0 + NB + rb * RPC;
  const int rows = min(RPC, n - r0);
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
#pragma unroll 1
    for (int sp = 0; sp < a.splits; sp += 2) {
      float4 pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int q = tid + (e & 3) * 256, i = q >> 5, j4 = (q & 31) * 4;
        pv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows && sp + (e >> 2) < a.splits)
          pv[e] = __ldcg(reinterpret_cast<const float4*>(a.partials + (sp + (e >> 2)) * a.split_stride +
                                                         static_cast<long long>(r0 - j0 + i) * NB + j4));
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
  if (blockIdx.x == 17 && rb == 16) PT3(41);
  if (a.late) {
    // look-ahead term: at -= L[r0 .., j0-128 : j0) L[j0 : j0+128, j0-128 : j0)^T  (32 x 128, K = 128, fp32 FMAs; thread =
    // one column x 16 rows: a warp reads 32 consecutive rows of s and broadcasts the rows of la)
    if (!lp_staged) stage_lp();
    {
      const float* lr = a.L + static_cast<long long>(r0) * a.ldl + (j0 - NB);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = tid + e * 256, i = q >> 5, j4 = (q & 31) * 4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows) t = __ldcg(reinterpret_cast<const float4*>(lr + static_cast<long long>(i) * a.ldl + j4));
        *reinterpret_cast<float4*>(la + i * DS + j4) = t;
      }
    }
    cta_sync();
    // thread = 4 rows (the warp's: broadcast reads) x 4 columns (lane, lane + 32, ...: conflict-free reads): eight 16-byte
    // loads per 64 FMAs.  (One column x 16 rows per thread needed seventeen, and the shared-memory pipe, not the FMA pipe,
    // set the pace of this loop: 544 LDS.128 per thread in eight warps.)
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    const float* xrow = la + (4 * warp) * DS;
    const float* bcol = s + lane * DS;
#pragma unroll 2
    for (int k = 0; k < NB; k += 4) {
      float4 x[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(xrow + r * DS + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(bcol + (32 * c) * DS + k);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          acc[r][c] = fmaf(x[r].x, b[c].x, fmaf(x[r].y, b[c].y, fmaf(x[r].z, b[c].z, fmaf(x[r].w, b[c].w, acc[r][c]))));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) at[(4 * warp + r) * DS + lane + 32 * c] -= acc[r][c];
    lp_staged = false;  // the stages below overwrite s
  }
  if (blockIdx.x == 1) PT3(21);
  if (blockIdx.x == 17 && rb == 16) PT3(42);
  const float* l11 = a.L + static_cast<long long>(j0) * a.ldl + j0;
  constexpr int CPT = 32 / TPR;  // columns per thread in the block-update phase
  const int row = tid % RPC, part = tid / RPC;
  float* myrow = at + row * DS;
#pragma unroll 1
  for (int J = 0; J < NB / 32; ++J) {
    if (J > 0) {
      // block-update with the finished block-columns I < J of block-row J (final once step J-1's rows are published)
      wait_epoch(a.ready + 2, a.epoch_base + 2 * J, j0);
      const int ncol4 = 8 * J;
#pragma unroll 1
      for (int q = tid; q < 32 * ncol4; q += 256) {
        const int i = 32 * J + q / ncol4, j4 = (q % ncol4) * 4;
        *reinterpret_cast<float4*>(s + i * DS + j4) = __ldcg(reinterpret_cast<const float4*>(l11 + static_cast<long long>(i) * a.ldl + j4));
      }
      cta_sync();
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
    cta_sync();
    if (tid < RPC) row_solve32<1>(at + tid * DS + 32 * J, dT, dinv);  // in-block forward substitution, one thread per row
    cta_sync();
    {
      // these 32 columns are final: store them now (one 16-byte group per thread), so that only the last stage's quarter
      // of the panel is left to write after CTA 0's last epoch
      const int i = tid >> 3, j4 = 32 * J + (tid & 7) * 4;
      if (i < rows)
        store_l4(a.L + static_cast<long long>(r0 + i) * a.ldl + j0, a.Lhi + static_cast<long long>(r0 + i) * a.ldh + j0,
                 a.Llo + static_cast<long long>(r0 + i) * a.ldh + j0, j4, *reinterpret_cast<const float4*>(at + i * DS + j4), sl);
    }
  }
  if (blockIdx.x == 1) PT3(23);
  cta_sync();
  if (tid == 0 && a.track && rb < 4) {  // one of the four row blocks the next panel's CTA 0 reads is stored
    __threadfence();
    atomicAdd(a.done + 2, 1u);
  }
  }  // row blocks
  if (blockIdx.x == 1) PT3(24);
  if (blockIdx.x == 17) PT3(43);
}

// fp32 arrays + the two 96 x 128 B swizzled tf32 operand tiles of CTA 0's tensor-core update (1 KiB alignment slack)
constexpr int PANEL_SMEM = (NB * DS + 2 * RPC * DS + 32 * DT) * static_cast<int>(sizeof(float)) + 2 * 12288 + 1024;
// late-term operand tiles of CTA 0 (fused kernel only): 4 x 16 KiB behind the panel arrays, 1 KiB aligned like pt_addr
constexpr int LATE_SMEM = 1024 + LATE_TILE_OFF + 4 * H3_TILE_BYTES;
constexpr int FUSED_SMEM0 = PANEL_SMEM > H3_SMEM_BYTES ? PANEL_SMEM : H3_SMEM_BYTES;
constexpr int FUSED_SMEM = FUSED_SMEM0 > LATE_SMEM ? FUSED_SMEM0 : LATE_SMEM;
static_assert(FUSED_SMEM <= 227 * 1024, "fused Cholesky kernel: shared memory");

template <bool FULL, bool TIMING>
__global__ void __launch_bounds__(256, 1) potrf_panel_h3_kernel(const PanelArgs a) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  panel_body<FULL, TIMING, false>(a, sm_raw, nullptr, nullptr);
}

// Look-ahead launch of panel k: CTAs [0, panel_ctas) run the panel program above (CTA 0 = diagonal block, then the row
// owners), CTAs [panel_ctas, gridDim.x) each run one (tile, split) work item of the NEXT panel's update GEMM restricted to
// the block-columns before panel k (final since the previous launch) - the tensor-core work that used to sit between two
// panel kernels now fills the SMs the panel leaves idle, and nothing in one role waits for the other.  One CTA per SM
// (the GEMM role's shared-memory footprint), grid <= number of SMs.
template <bool TIMING, bool LEAN>
__global__ void __launch_bounds__(H3_THREADS, 1)
potrf_fused_h3_kernel(const PanelArgs a, const H3Args g, const __grid_constant__ CUtensorMap tmAhi,
                      const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmAlo,
                      const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmLhi,
                      const __grid_constant__ CUtensorMap tmLlo, const int panel_ctas, const int gemm_tiles,
                      const int gemm_items) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  if (static_cast<int>(blockIdx.x) < panel_ctas) {
    if (threadIdx.x < 256) panel_body<true, TIMING, LEAN>(a, sm_raw, &tmLhi, &tmLlo);
    return;
  }
  // work items (row tile, split) of the hosted update, round-robin over the GEMM-role CTAs (one item each except when the
  // row owners need the SMs: every 32-row block of the panel gets a CTA of its own first)
  const int gemm_ctas = static_cast<int>(gridDim.x) - panel_ctas;
#pragma unroll 1
  for (int w = blockIdx.x - panel_ctas; w < gemm_items; w += gemm_ctas) {
  gemm_h3_body<false, false>(g, tmAhi, tmBhi, tmAlo, tmBlo, sm_raw, w % gemm_tiles, w / gemm_tiles, w + gemm_ctas >= gemm_items);
  if (a.nb_out != nullptr && w % gemm_tiles == 0) {
    // row tile 0 of the hosted update is the next panel's diagonal tile.  gemm_h3_body ended with __syncthreads(): this
    // CTA's partial plane is stored; the last of the tile's CTAs reduces all planes (fixed order) into nb_out.
    __shared__ unsigned is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(a.nb_count, 1u) + 1u == a.nb_target) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
#pragma unroll 1
      for (int q = threadIdx.x; q < NB * (NB / 4); q += H3_THREADS) {
        const int i = q >> 5, j4 = (q & 31) * 4;
        if (j4 > i) continue;  // lower triangle (16-byte groups that touch it)
        float4 v = *reinterpret_cast<const float4*>(a.nb_A + static_cast<long long>(i) * a.lda + j4);
        for (int sp = 0; sp < a.nb_splits; ++sp) {
          const float4 pv = __ldcg(reinterpret_cast<const float4*>(a.nb_partials + sp * a.nb_stride + static_cast<long long>(i) * NB + j4));
          v.x -= pv.x; v.y -= pv.y; v.z -= pv.z; v.w -= pv.w;
        }
        *reinterpret_cast<float4*>(a.nb_out + i * NB + j4) = v;
      }
      if (a.track) {  // the reduced tile is written: what the next panel's CTA 0 waits for when launches are chained
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(a.done + 1, 1u);
        }
      }
    }
  }
  __syncthreads();
  }
}

template <typename Kern, typename... Args>
cudaError_t launch_maybe_pdl(Kern kern, int grid, int block, int smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] == '0' ? 0 : 1;
}

// split-K factor of an update GEMM with `tiles` row tiles and kb k-blocks when `ctas` CTAs are available
int pick_splits(int tiles, int kb, int ctas) {
  int S = ctas / tiles;
  if (S < 1) S = 1;
  if (S > MAX_SPLITS) S = MAX_SPLITS;
  if (S > kb) S = kb;
  return S;
}

}  // namespace

size_t potrf_h3_workspace_bytes(int n) {
  // 256 bytes of scalars (epoch words, counters) + two reduced diagonal blocks [128][128] + two buffers of
  // split-K partials (the look-ahead GEMM of panel k+1 writes one while panel k reads the other): at most
  // MAX_SPLITS x rows x 128 each, with splits * tiles bounded by the SM count
  const long long rows = n > 0 ? n : 1;
  long long worst = 0;
  for (long long j0 = NB; j0 < rows; j0 += NB) {
    const long long M = rows - j0, tiles = (M + NB - 1) / NB;
    const long long S = pick_splits(static_cast<int>(tiles), static_cast<int>(j0 / H3_BK), 148);
    if (S * M > worst) worst = S * M;
  }
  return 256 + 2 * static_cast<size_t>(NB) * NB * sizeof(float) + 2 * static_cast<size_t>(worst) * NB * sizeof(float);
}

// The launch plan of a factorisation (one row of eight ints per panel), recorded instead of launched when `plan` is set:
// the grid arithmetic below is what keeps every spin-wait's partner resident, so it is testable without a GPU.
struct PlanSink {
  int sms;       // number of SMs to plan for
  int* rows;     // [max_rows][8]: j0, fused, panel CTAs (1 + row owners), GEMM CTAs, GEMM row tiles, GEMM splits,
                 //                partial planes this panel reads, helpers
  int max_rows, count;
};

static int potrf_h3_impl(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, const gsmvi_h3_operand& Lh,
                         int n, int* flag, void* workspace, int zero_upper, PlanSink* plan) {
  if (!plan) {
    if (n <= 0 || !A || !L || !flag || !workspace || !Lh.hi || !Lh.lo || !Lh.scale) return GSMVI_EINVAL;
    if ((lda & 3) != 0 || (ldl & 3) != 0 || (Lh.ld & 7) != 0 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(L)) & 15) != 0)
      return GSMVI_EALIGN;
  } else if (n <= 0) {
    return GSMVI_EINVAL;
  }
  static PerDeviceOnce attr_set;
  if (!plan && !attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(potrf_panel_h3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_panel_h3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_panel_h3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_fused_h3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_fused_h3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_fused_h3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(potrf_fused_h3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set.set();
  }
  long long worst = 0;
  for (long long j0 = NB; j0 < n; j0 += NB) {
    const long long M = n - j0, tiles = (M + NB - 1) / NB;
    const long long S = pick_splits(static_cast<int>(tiles), static_cast<int>(j0 / H3_BK), 148);
    if (S * M > worst) worst = S * M;
  }
  unsigned* ready = static_cast<unsigned*>(workspace);
  float* d0 = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  float* dbase[2] = {d0, d0 + NB * NB};  // reduced diagonal tiles, ping-pong between panels (helpers use the first)
  float* pbuf[2] = {d0 + 2 * NB * NB, d0 + 2 * NB * NB + worst * NB};
  unsigned helper_target = 0;
  __half* Lhi = static_cast<__half*>(Lh.hi);
  __half* Llo = static_cast<__half*>(Lh.lo);
  if (!plan) {
    potrf_prepare_kernel<<<1, 256, 0, stream>>>(A, lda, n, Lh.scale, ready, flag);
    if (zero_upper && n > NB)
      potrf_zero_upper_kernel<<<dim3((n / 4 + 255) / 256, n), 256, 0, stream>>>(L, ldl, Lhi, Llo, Lh.ld, n);
  }
  unsigned epoch = 0;
  static PerDeviceInt dev_ctas_pd;
  int& dev_ctas = dev_ctas_pd.ref();
  if (dev_ctas == 0 && !plan) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dev_ctas = sms > 17 ? sms : 148;  // one CTA per SM; helpers need 16 TRSM CTAs
  }
  const int max_ctas = plan ? (plan->sms > 17 ? plan->sms : 148) : dev_ctas;
  auto record = [&](int j0_, int fused_, int panel_ctas_, int gemm_ctas_, int gemm_tiles_, int gemm_splits_, int splits_in_, int helpers_) {
    if (plan->count < plan->max_rows) {
      int* r = plan->rows + 8 * plan->count;
      r[0] = j0_; r[1] = fused_; r[2] = panel_ctas_; r[3] = gemm_ctas_; r[4] = gemm_tiles_; r[5] = gemm_splits_; r[6] = splits_in_; r[7] = helpers_;
    }
    ++plan->count;
  };
  const bool timing = getenv("GSMVI_POTRF_TIMING") != nullptr;
  static int pdl_env = -1, look_env = -1, late_env = -1, chain_env = -1;
  // chained launches: see PanelArgs::chained and profiles/r02_potrf_h3_round2b.txt for what they measured
  if (chain_env < 0) chain_env = env_flag("GSMVI_POTRF_CHAIN", 0);
  if (late_env < 0) late_env = env_flag("GSMVI_POTRF_LATE_MMA", 1);
  if (pdl_env < 0) pdl_env = env_flag("GSMVI_POTRF_PDL", 1);
  if (look_env < 0) look_env = env_flag("GSMVI_POTRF_LOOKAHEAD", 1);
  const bool pdl = pdl_env == 1;
  // Look-ahead (GSMVI_POTRF_LOOKAHEAD=0 disables): every full panel k >= 1 is one fused launch whose spare CTAs run the
  // update GEMM of panel k+1 over the block-columns before panel k; panel k+1 then adds the K = 128 term of panel k itself.
  // (one GEMM work item per spare CTA: the next panel's row tiles must fit beside CTA 0 and the 16 helpers)
  const bool look = look_env == 1 && n >= 4 * NB && max_ctas >= 64 && (n + NB - 1) / NB - 2 <= max_ctas - 17;
  // whole-matrix views of the fp16 pair for CTA 0's late-term loads (box 64 x 128; coordinates are passed at run time)
  CUtensorMap tmL[2] = {};
  const bool late_mma = look && late_env == 1;
  if (late_mma && !plan) {
    int rc = h3_make_tmap(&tmL[0], Lhi, n, n, Lh.ld, H3_BK, NB);
    if (rc == GSMVI_OK) rc = h3_make_tmap(&tmL[1], Llo, n, n, Lh.ld, H3_BK, NB);
    if (rc != GSMVI_OK) return rc;
  }
  const float* next_base = nullptr;  // reduced diagonal tile the NEXT panel will find (late_mma)
  unsigned base_target = 0;
  unsigned first4_done = 0;  // row blocks 0..3 of all panels so far
  unsigned gemm_done = 0;    // reduced diagonal tiles written by all launches so far (chained CTA 0s wait for theirs)
  bool prev_fused = false;
  int next_splits = 0;  // split count of the look-ahead partials the NEXT panel will find in pbuf[(k + 1) & 1]
  for (int j0 = 0, k = 0; j0 < n; j0 += NB, ++k) {
    const int nb = min(NB, n - j0);
    const int M = n - j0, rest = M - nb;
    const int nblocks = (rest + RPC - 1) / RPC;
    PanelArgs pa;
    pa.A = A; pa.lda = lda; pa.L = L; pa.ldl = ldl; pa.Lhi = Lhi; pa.Llo = Llo; pa.ldh = Lh.ld; pa.scale_l = Lh.scale;
    pa.n = n; pa.j0 = j0; pa.nb = nb; pa.partials = pbuf[k & 1]; pa.splits = 0; pa.split_stride = static_cast<long long>(M) * NB;
    pa.flag = flag; pa.ready = ready; pa.epoch_base = epoch;
    pa.d0 = d0; pa.helper_count = ready + 1; pa.helpers = 0; pa.helper_target = 0; pa.late = 0; pa.late_mma = 0;
    pa.base = nullptr; pa.nb_out = nullptr; pa.nb_A = nullptr; pa.nb_partials = nullptr; pa.nb_stride = 0; pa.nb_splits = 0;
    pa.nb_count = ready + 3; pa.nb_target = 0;
    pa.chained = 0; pa.early_scale = 0; pa.track = (pdl && chain_env == 1) ? 1 : 0; pa.done = ready + 4; pa.gemm_target = gemm_done; pa.first4_target = first4_done;
    epoch += 8;
    const bool fused = look && nb == NB;
    if (fused) {
      // ---- this panel: look-ahead partials (from the previous launch, if any) + the late term
      if (k >= 1) {
        pa.splits = next_splits;
        pa.late = 1;
        if (late_mma) {
          pa.late_mma = 1;
          pa.base = next_base;
        } else {
          pa.helpers = 16;
          helper_target += 16;
          pa.helper_target = helper_target;
        }
      }
      // ---- hosted work: update of panel k+1 (if it is a full panel) with block-columns [0, j0)
      const int nj0 = j0 + NB;
      const bool host_next = k >= 1 && nj0 < n && n - nj0 >= NB;
      int G = 0, gtiles = 1, S = 0, gitems = 0;
      H3Args ga = {};
      CUtensorMap tm[4] = {};
      if (host_next) {
        const int Mn = n - nj0;
        gtiles = (Mn + NB - 1) / NB;
        S = pick_splits(gtiles, j0 / H3_BK, max_ctas - 1 - (nblocks > 16 ? nblocks : 16));
        while (S > 1 && 1 + 16 + gtiles * S > max_ctas) --S;
        {  // never more planes than potrf_h3_workspace_bytes sized the partial buffers for (a no-op up to 148 SMs)
          const int s_ws = pick_splits(gtiles, j0 / H3_BK, 148);
          if (S > s_ws) S = s_ws;
        }
        G = gtiles * S;
        gitems = G;
        {  // every row block of this panel gets its own CTA first; the update's items share what is left
          static int gcap_env = -1;
          if (gcap_env < 0) gcap_env = env_flag("GSMVI_POTRF_GCAP", 1);
          // (only while no GEMM CTA has to take more than two items: at larger n, where the row blocks alone outnumber the
          // SMs, the items keep one CTA each and the row owners loop instead - the update would become the long pole)
          const int room = max_ctas - 1 - (nblocks > 0 ? nblocks : 1);
          if (gcap_env == 1 && G > room && 2 * room >= G) G = room;
        }
        HView va{Lhi + static_cast<long long>(nj0) * Lh.ld, Llo + static_cast<long long>(nj0) * Lh.ld, Mn, j0, Lh.ld, Lh.scale};
        HView vb{Lhi + static_cast<long long>(nj0) * Lh.ld, Llo + static_cast<long long>(nj0) * Lh.ld, NB, j0, Lh.ld, Lh.scale};
        H3Opts o;
        o.splits = S;
        o.split_stride = static_cast<long long>(Mn) * NB;
        dim3 ggrid;
        if (!plan) {
          int rc = h3_prepare(Mn, NB, j0, va, vb, pbuf[(k + 1) & 1], NB, o, &ga, tm, &ggrid);
          if (rc != GSMVI_OK) return rc;
        }
      }
      next_splits = S;
      next_base = nullptr;
      if (host_next && late_mma) {
        pa.nb_out = dbase[(k + 1) & 1];
        pa.nb_A = A + static_cast<long long>(nj0) * lda + nj0;
        pa.nb_partials = pbuf[(k + 1) & 1];
        pa.nb_stride = static_cast<long long>(n - nj0) * NB;
        pa.nb_splits = S;
        base_target += static_cast<unsigned>(S);
        pa.nb_target = base_target;
        next_base = pa.nb_out;
      }
      int T = nblocks;
      if (pa.helpers > 0 && T < 16) T = 16;
      if (T > max_ctas - 1 - G) T = max_ctas - 1 - G;
      if (T < (pa.helpers > 0 ? 16 : 0)) return GSMVI_EINVAL;  // cannot happen: G <= max_ctas - 17 by construction
      pa.trsm_ctas = T > 0 ? T : 1;
      // chained to the previous fused launch: wait for its row blocks and its hosted GEMM instead of the grid's completion
      pa.chained = (prev_fused && pdl && chain_env == 1) ? 1 : 0;
      pa.early_scale = prev_fused ? 1 : 0;
      first4_done += static_cast<unsigned>(nblocks < 4 ? nblocks : 4);
      gemm_done += (host_next && late_mma) ? 1u : 0u;  // one reduced diagonal tile per hosting launch
      prev_fused = true;
      const int grid = 1 + T + G;
      if (plan) {
        record(j0, 1, 1 + T, G, host_next ? gtiles : 0, S, pa.splits, pa.helpers);
        continue;
      }
      cudaError_t le;
      auto kern = late_mma ? (timing ? potrf_fused_h3_kernel<true, true> : potrf_fused_h3_kernel<false, true>)
                           : (timing ? potrf_fused_h3_kernel<true, false> : potrf_fused_h3_kernel<false, false>);
      le = launch_maybe_pdl(kern, grid, H3_THREADS, FUSED_SMEM, stream, pdl, pa, ga, tm[0], tm[1], tm[2], tm[3], tmL[0], tmL[1], 1 + T, gtiles, gitems);
      if (le != cudaSuccess) return static_cast<int>(le);
      continue;
    }
    next_splits = 0;
    next_base = nullptr;
    prev_fused = false;
    first4_done += static_cast<unsigned>(nblocks < 4 ? nblocks : 4);
    if (j0 > 0) {
      const int tiles = (M + NB - 1) / NB;
      const int S = pick_splits(tiles, j0 / H3_BK, 148);
      pa.splits = S;
      HView va{Lhi + static_cast<long long>(j0) * Lh.ld, Llo + static_cast<long long>(j0) * Lh.ld, M, j0, Lh.ld, Lh.scale};
      HView vb{Lhi + static_cast<long long>(j0) * Lh.ld, Llo + static_cast<long long>(j0) * Lh.ld, nb, j0, Lh.ld, Lh.scale};
      H3Opts o;
      o.splits = S;
      o.split_stride = pa.split_stride;
      o.pdl = pdl;
      if (!plan) {
        int rc = launch_gemm_h3(stream, M, nb, j0, va, vb, pbuf[k & 1], NB, o);
        if (rc != GSMVI_OK) return rc;
      }
    }
    int grid = 1 + nblocks;
    if (grid > max_ctas) grid = max_ctas;
    pa.trsm_ctas = grid > 1 ? grid - 1 : 1;
    if (pa.splits > 0 && grid - 1 >= 16) {
      pa.helpers = 16;
      helper_target += 16;
      pa.helper_target = helper_target;
    }
    if (plan) {
      const int tiles = (M + NB - 1) / NB;
      record(j0, 0, nb < NB ? 1 : grid, pa.splits > 0 ? tiles * pa.splits : 0, pa.splits > 0 ? tiles : 0, pa.splits, pa.splits, pa.helpers);
      continue;
    }
    cudaError_t le;
    if (nb < NB) le = launch_maybe_pdl(potrf_panel_h3_kernel<false, false>, 1, 256, PANEL_SMEM, stream, false, pa);
    else if (timing) le = launch_maybe_pdl(potrf_panel_h3_kernel<true, true>, grid, 256, PANEL_SMEM, stream, pdl, pa);
    else le = launch_maybe_pdl(potrf_panel_h3_kernel<true, false>, grid, 256, PANEL_SMEM, stream, pdl, pa);
    if (le != cudaSuccess) return static_cast<int>(le);
  }
  if (plan) return GSMVI_OK;
  if (timing && n > NB * 9) {
    long long h[64];
    cudaStreamSynchronize(stream);
    {
      static unsigned long long hp[64][3];
      cudaMemcpyFromSymbol(hp, g_pt_panel, sizeof(hp));
      const int np = (n + NB - 1) / NB < 64 ? (n + NB - 1) / NB : 64;
      fprintf(stderr, "[potrf_h3 panels] CTA 0 per panel, ns: start phase / stages / gap to the next panel's start:");
      for (int k = 0; k < np; ++k)
        fprintf(stderr, " %d:%lld/%lld/%lld", k, (long long)(hp[k][1] - hp[k][0]), (long long)(hp[k][2] - hp[k][1]),
                k + 1 < np ? (long long)(hp[k + 1][0] - hp[k][2]) : 0LL);
      fprintf(stderr, "\n");
    }
    cudaMemcpyFromSymbol(h, g_pt3, sizeof(h));
    fprintf(stderr, "[potrf_h3 panel 8] CTA0 start phase (cycles since its first stamp): setup + TMA issued %lld, base loads issued %lld, late MMAs issued %lld, "
            "base in smem %lld, late MMAs complete %lld, drained %lld\n", h[50] - h[0], h[51] - h[0], h[52] - h[0], h[53] - h[0], h[54] - h[0], h[1] - h[0]);
    fprintf(stderr, "[potrf_h3 panel 8] CTA0: load %lld |", h[1] - h[0]);
    for (int p = 0; p < 4; ++p)
      fprintf(stderr, " p%d chol32 %lld trsm(t0) %lld sync %lld upd %lld |", p, h[2 + 4 * p] - (p ? h[1 + 4 * p] : h[1]),
              h[3 + 4 * p] - h[2 + 4 * p], p < 3 ? h[4 + 4 * p] - h[3 + 4 * p] : 0LL, p < 3 ? h[5 + 4 * p] - h[4 + 4 * p] : 0LL);
    fprintf(stderr, " total %lld || CTA1: prefetch %lld, stage-3 start at %lld, solve end %lld, store %lld (since CTA0 start)\n",
            h[18] - h[0], h[21] - h[20], h[22] - h[0], h[23] - h[0], h[24] - h[0]);
    fprintf(stderr, "[potrf_h3 panel 8] since CTA0 start: helper(CTA1) start %lld staged %lld fma %lld counted %lld | row CTA 17: released %lld "
            "prefetched %lld late-term %lld end %lld\n", h[20] - h[0], h[30] - h[0], h[31] - h[0], h[32] - h[0], h[40] - h[0],
            h[41] - h[0], h[42] - h[0], h[43] - h[0]);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int potrf_h3(cudaStream_t stream, const float* A, long long lda, float* L, long long ldl, const gsmvi_h3_operand& Lh, int n,
             int* flag, void* workspace, int zero_upper) {
  return potrf_h3_impl(stream, A, lda, L, ldl, Lh, n, flag, workspace, zero_upper, nullptr);
}

int potrf_h3_plan(int n, int sms, int* rows, int max_rows) {
  if (n <= 0 || sms <= 0 || (!rows && max_rows > 0)) return GSMVI_EINVAL;
  PlanSink sink{sms, rows, max_rows, 0};
  gsmvi_h3_operand none = {};
  const int rc = potrf_h3_impl(nullptr, nullptr, 0, nullptr, 0, none, n, nullptr, nullptr, 0, &sink);
  return rc == GSMVI_OK ? sink.count : (rc > 0 ? -rc : rc);
}

}  // namespace gsmvi
