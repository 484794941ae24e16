// tcgen05 TF32 / 3xTF32 GEMM for sm_100a:   C = alpha * op(A) * op(B)^T + beta * Cin + bias
//
// This one kernel is the tensor-core engine behind every O(B D^2) / O(D^3) contraction of the GSM and BaM
// iteration: the sampler's X = mu + Z L^T, the dense-Gaussian
// score G = -(X-m)P, W = G Sigma, the signed outer-product accumulation D^T D - E^T E, the centered batch
// covariances, and the TRSM / SYRK steps of the blocked Cholesky.
//
// Design (B200):
//  * operands are plain fp32 in HBM; TMA (cp.async.bulk.tensor, 128B swizzle) stages 128x32 tiles in smem
//  * 3xTF32: kind::tf32 TRUNCATES the 13 low mantissa bits of its fp32 inputs (measured on B200,
//    profiles/r01_gemm_probe.md), so the staged fp32 tile itself is the "hi" operand.  Eight converter warps
//    write lo = tf32_rn(a - trunc(a)) next to it (element-wise, so the swizzled layout is untouched), and one
//    elected thread issues lo*hi + hi*lo (into a correction accumulator) and hi*hi (into main accumulators)
//    as three tcgen05.mma.kind::tf32 per k-slice.  TMEM accumulation also truncates (error grows linearly
//    with the number of accumulating MMAs), so hi*hi is spread round-robin over three main accumulators and
//    the small terms go to their own; the epilogue adds the four in fp32 round-to-nearest.
//    The default (precision 3) rounds hi to nearest instead and stores it too: ~8x more accurate products
//    (fp32-grade) for ~10% more shared-memory traffic.  1xTF32 mode skips the split (one accumulator).
//  * pre-split operands: ncu shows the converters' LDS/STS traffic saturating the shared-memory pipe (tensor pipe
//    ~48% active).  An operand may therefore come with a precomputed `lo` array (PRE bit): TMA then loads hi and lo
//    tiles directly and the converters skip it.  Reused operands (Sigma, P, L) carry lo = tf32_rn(a - trunc(a)) next
//    to the raw array (the tensor core truncates the raw value itself); when both operands are pre-split
//    (covariance update: the row pass writes T_hi / T_lo) the kernel runs with no conversion at all.
//  * warp-specialised: warp0 = TMA producer, warp1 = MMA issuer + TMEM owner, warps 2..9 = converters,
//    and the same eight warps drain TMEM in the epilogue (tcgen05.ld 32x32b) with a fused
//    alpha/beta/bias update, optional lower-triangle-only tiles and mirrored (symmetric) stores.
//  * operands may be K-major ([rows, K] row-major) or MN-major ([K, rows] row-major), so A^T A style
//    statistics need no transposed copy; an instruction-descriptor negate bit turns the second half of the
//    K range into a subtraction (D^T D - E^T E in one pass over [D;E]).
//  * triangular operands skip the k-blocks that are structurally zero.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsmvi_b200.h"
#include "ptx.cuh"

namespace gsmvi {

constexpr int BM = 128;        // tile rows   (UMMA M)
constexpr int BN = 128;        // tile cols   (UMMA N)
constexpr int BK = 32;         // fp32 per 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32: 32 bytes of K per instruction
constexpr int TILE_BYTES = BM * BK * 4;  // 16 KiB per operand tile
constexpr int NUM_CONV_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * NUM_CONV_WARPS;  // 320
constexpr int BAR_BYTES = 1024;

// K-range policy bits: structural zeros of triangular operands (in the op() orientation: A is M x K, B is N x K)
enum : int {
  KR_FULL = 0,
  KR_A_LOWER = 1,  // A[m,k] == 0 for k > m
  KR_B_LOWER = 2,  // B[n,k] == 0 for k > n
  KR_A_UPPER = 4,  // A[m,k] == 0 for k < m
  KR_B_UPPER = 8,  // B[n,k] == 0 for k < n
};

struct GemmArgs {
  int M, N, K;
  float alpha, beta;
  const float* Cin;  // may alias C; ignored when beta == 0
  long long ldcin;
  float* C;
  long long ldc;
  const float* bias_n;  // optional, length N, added to every row
  int tri;              // 1: only tiles with tile_m >= tile_n are computed (diagonal tiles store n <= m only)
  int mirror;           // with tri: also store C[n,m] = C[m,n]
  int krange;           // KR_* bits
  int neg_from;         // first k (multiple of BK) whose A slices are negated; >= K: never
  int tiles_m, tiles_n;
};

template <int NPASS>
struct GemmCfg {
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES * (NPASS == 3 ? 2 : 1);  // [A|B] (+ [A_lo|B_lo])
  static constexpr int STAGES = (NPASS == 3) ? 3 : 6;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + BAR_BYTES + STAGES * STAGE_BYTES;
  static constexpr int N_MAIN = (NPASS == 3) ? 3 : 1;               // round-robin hi*hi accumulators
  static constexpr int TMEM_COLS = (NPASS == 3) ? 512 : 128;        // N_MAIN (+1 correction) x BN columns
  static constexpr int CORR_COL = N_MAIN * BN;
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  // SM100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout type [61,64): 2 = SWIZZLE_128B (16 B atoms), 1 = SWIZZLE_128B_BASE32B (32 B atoms; the only
  // layout the tensor core accepts for MN-major tf32 operands).
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(bool a_mn, bool b_mn, bool a_neg) {
  // c_format F32 (1) [4,6); a_format TF32 (2) [7,10); b_format TF32 (2) [10,13); a_negate 13; a_major 15;
  // b_major 16; N>>3 [17,23); M>>4 [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_neg ? 1u : 0u) << 13) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

template <int NPASS, bool SPLIT_RN, bool A_MN, bool B_MN, int PRE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const GemmArgs args, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo) {
  constexpr bool A_PRE = (PRE & 1) != 0, B_PRE = (PRE & 2) != 0;  // operand arrives pre-split (hi via tmA/tmB, lo via tmAlo/tmBlo)
  constexpr bool NEED_CONV = (NPASS == 3) && !(A_PRE && B_PRE);
  using Cfg = GemmCfg<NPASS>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t bar_base = ptx::smem_u32(smem);
  const uint32_t stage_base = bar_base + BAR_BYTES;
  // barrier slots (8 bytes each): full[s], ready[s], empty[s], acc_full
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (3 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 8 * (3 * STAGES + 1) + 8);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int tm, tn;
  if (args.tri) {
    const int t = blockIdx.x;
    int i = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    tm = i;
    tn = t - i * (i + 1) / 2;
  } else {
    // group GROUP tile-rows together so concurrently resident CTAs share A and B tiles in L2
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
  const int m0 = tm * BM, n0 = tn * BN;

  // ---- K range (whole BK blocks)
  int k_begin = 0, k_end = args.K;
  if (args.krange & KR_A_LOWER) k_end = min(k_end, m0 + BM);
  if (args.krange & KR_B_LOWER) k_end = min(k_end, n0 + BN);
  if (args.krange & KR_A_UPPER) k_begin = max(k_begin, m0);
  if (args.krange & KR_B_UPPER) k_begin = max(k_begin, n0);
  const int kb_begin = k_begin / BK;
  const int kb_end = (k_end > k_begin) ? (k_end + BK - 1) / BK : kb_begin;
  const int num_kb = kb_end - kb_begin;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(ready_bar(s), NUM_CONV_WARPS);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer (the warp stays converged, one elected lane issues: see ptx::elect_one) =========
    for (int it = 0; it < num_kb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      const int k0 = (kb_begin + it) * BK;
      const uint32_t sA = stage_base + s * Cfg::STAGE_BYTES;
      const uint32_t sB = sA + TILE_BYTES;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar(s), (2 + (A_PRE ? 1 : 0) + (B_PRE ? 1 : 0)) * TILE_BYTES);
        if (!A_MN) {
          ptx::tma_load_2d(sA, &tmA, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) ptx::tma_load_2d(sA + c * (BK * 128), &tmA, full_bar(s), m0 + 32 * c, k0);
        }
        if (!B_MN) {
          ptx::tma_load_2d(sB, &tmB, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) ptx::tma_load_2d(sB + c * (BK * 128), &tmB, full_bar(s), n0 + 32 * c, k0);
        }
        if (A_PRE) {
          if (!A_MN) {
            ptx::tma_load_2d(sA + 2 * TILE_BYTES, &tmAlo, full_bar(s), k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c)
              ptx::tma_load_2d(sA + 2 * TILE_BYTES + c * (BK * 128), &tmAlo, full_bar(s), m0 + 32 * c, k0);
          }
        }
        if (B_PRE) {
          if (!B_MN) {
            ptx::tma_load_2d(sB + 2 * TILE_BYTES, &tmBlo, full_bar(s), k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c)
              ptx::tma_load_2d(sB + 2 * TILE_BYTES + c * (BK * 128), &tmBlo, full_bar(s), n0 + 32 * c, k0);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane) =====================
    {
      constexpr uint32_t idesc_pos = make_idesc_tf32(A_MN, B_MN, false);
      constexpr uint32_t idesc_neg = make_idesc_tf32(A_MN, B_MN, true);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), LBO unused.
      // MN-major (SWIZZLE_128B_BASE32B): 32-wide MN chunks BK*128 B apart (LBO), 4-deep K atoms 512 B apart (SBO);
      // one instruction (K=8) spans two K atoms, so the next k-slice starts 1024 B further.
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_SBO = A_MN ? 512 : 1024, A_KSTEP = A_MN ? 1024 : UMMA_K * 4;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_SBO = B_MN ? 512 : 1024, B_KSTEP = B_MN ? 1024 : UMMA_K * 4;
      constexpr uint32_t A_LT = A_MN ? 1 : 2, B_LT = B_MN ? 1 : 2;
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(NEED_CONV ? ready_bar(s) : full_bar(s), ph);
        ptx::tc_fence_after_sync();
        const uint32_t sA = stage_base + s * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + TILE_BYTES;
        const uint32_t idesc = ((kb_begin + it) * BK >= args.neg_from) ? idesc_neg : idesc_pos;
        const uint32_t t_main = tmem_base + (it % Cfg::N_MAIN) * BN;
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t da = make_smem_desc(sA + kk * A_KSTEP, A_LBO, A_SBO, A_LT);
            const uint64_t db = make_smem_desc(sB + kk * B_KSTEP, B_LBO, B_SBO, B_LT);
            const uint32_t acc_main = (it >= Cfg::N_MAIN || kk > 0) ? 1u : 0u;
            if (NPASS == 3) {
              const uint64_t da_lo = make_smem_desc(sA + 2 * TILE_BYTES + kk * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t db_lo = make_smem_desc(sB + 2 * TILE_BYTES + kk * B_KSTEP, B_LBO, B_SBO, B_LT);
              ptx::umma_tf32(tmem_base + Cfg::CORR_COL, da_lo, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
              ptx::umma_tf32(tmem_base + Cfg::CORR_COL, da, db_lo, idesc, 1u);
            }
            ptx::umma_tf32(t_main, da, db, idesc, acc_main);
          }
          ptx::umma_commit(empty_bar(s));  // smem slot reusable once these MMAs retire
          if (it + 1 == num_kb) ptx::umma_commit(acc_bar);
        }
        __syncwarp();
      }
      if (num_kb == 0 && ptx::elect_one()) ptx::mbar_arrive(acc_bar);
      __syncwarp();
    }
  } else {
    // ===================== converter warps, then epilogue =====================
    const int ct = threadIdx.x - 64;  // 0..255
    if (NEED_CONV) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(full_bar(s), ph);
        uint8_t* st = smem + BAR_BYTES + s * Cfg::STAGE_BYTES;
        // 16-byte chunks of the operands that still need splitting: [A|B], or only A / only B when the other is pre-split
        constexpr int CHUNK0 = A_PRE ? TILE_BYTES / 16 : 0;
        constexpr int CHUNKS = (A_PRE || B_PRE) ? TILE_BYTES / 16 : 2 * TILE_BYTES / 16;
#pragma unroll
        for (int j = 0; j < CHUNKS / (32 * NUM_CONV_WARPS); ++j) {
          const int c = CHUNK0 + ct + j * 32 * NUM_CONV_WARPS;
          const float4 v = *reinterpret_cast<const float4*>(st + c * 16);
          float4 lo;
          if (SPLIT_RN) {
            // precise split: hi = rn_tf32(v) overwrites the staged value, lo = rn_tf32(v - hi).  |lo| <= 2^-12 |v|,
            // the dropped lo*lo term is <= 2^-24 and unbiased: fp32-grade products.
            float4 hi;
            hi.x = ptx::to_tf32(v.x); lo.x = ptx::to_tf32(v.x - hi.x);
            hi.y = ptx::to_tf32(v.y); lo.y = ptx::to_tf32(v.y - hi.y);
            hi.z = ptx::to_tf32(v.z); lo.z = ptx::to_tf32(v.z - hi.z);
            hi.w = ptx::to_tf32(v.w); lo.w = ptx::to_tf32(v.w - hi.w);
            *reinterpret_cast<float4*>(st + c * 16) = hi;
          } else {
            // fast split: the tensor core truncates v to tf32 itself, so only lo = rn_tf32(v - trunc_tf32(v)) is
            // written.  lo in [0, 2^-10 |v|): the dropped lo*lo term is a one-sided ~2^-22 relative bias.
            lo.x = ptx::to_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
            lo.y = ptx::to_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
            lo.z = ptx::to_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
            lo.w = ptx::to_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
          }
          *reinterpret_cast<float4*>(st + 2 * TILE_BYTES + c * 16) = lo;
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ready_bar(s));
      }
    }
    // ---- epilogue: TMEM -> registers -> global
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after_sync();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;       // which 64 columns
    const int m = m0 + q * 32 + lane;
    const bool diag_tile = args.tri && (tm == tn);
    const float alpha = args.alpha, beta = args.beta;
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int c0 = half * 64 + chunk * 32;
      uint32_t r[32];
      if (num_kb > 0) {
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0;
        ptx::tmem_ld_32x32(t0, r);
        ptx::tmem_ld_wait();
        if (NPASS == 3) {
          const int n_main = min(num_kb, Cfg::N_MAIN);
          uint32_t t[32];
          for (int a = 1; a <= Cfg::N_MAIN; ++a) {  // remaining main accumulators, then the correction one
            if (a < Cfg::N_MAIN && a >= n_main) continue;
            ptx::tmem_ld_32x32(t0 + a * BN, t);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int nbase = n0 + c0;
      if (m < args.M && nbase < args.N) {
        float* crow = args.C + static_cast<long long>(m) * args.ldc + nbase;
        const float* cin = (beta != 0.0f) ? args.Cin + static_cast<long long>(m) * args.ldcin + nbase : nullptr;
        const bool vec_ok = !diag_tile && (nbase + 32 <= args.N) && ((args.ldc & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(args.C) & 15) == 0) &&
                            (cin == nullptr || (((args.ldcin & 3) == 0) && ((reinterpret_cast<uintptr_t>(args.Cin) & 15) == 0)));
        if (vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = alpha * __uint_as_float(r[j + 0]);
            o.y = alpha * __uint_as_float(r[j + 1]);
            o.z = alpha * __uint_as_float(r[j + 2]);
            o.w = alpha * __uint_as_float(r[j + 3]);
            if (cin) {
              const float4 ci = *reinterpret_cast<const float4*>(cin + j);
              o.x += beta * ci.x; o.y += beta * ci.y; o.z += beta * ci.z; o.w += beta * ci.w;
            }
            if (args.bias_n) {
              const float4 b = *reinterpret_cast<const float4*>(args.bias_n + nbase + j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            *reinterpret_cast<float4*>(crow + j) = o;
            r[j + 0] = __float_as_uint(o.x); r[j + 1] = __float_as_uint(o.y);
            r[j + 2] = __float_as_uint(o.z); r[j + 3] = __float_as_uint(o.w);
          }
          if (args.mirror) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              args.C[static_cast<long long>(nbase + j) * args.ldc + m] = __uint_as_float(r[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nbase + j;
            if (n < args.N && !(diag_tile && n > m)) {
              float o = alpha * __uint_as_float(r[j]);
              if (cin) o += beta * cin[j];
              if (args.bias_n) o += args.bias_n[n];
              crow[j] = o;
              if (args.mirror && n != m) args.C[static_cast<long long>(n) * args.ldc + m] = o;
            }
          }
        }
      }
    }
    ptx::tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side

// Operand view for the launcher: `rows` x `cols` fp32, row-major with leading dimension ld (elements).
// K-major operand: rows = M (or N), cols = K.   MN-major operand: rows = K, cols = M (or N).
struct MatView {
  const float* ptr;
  long long rows, cols, ld;
  const float* lo = nullptr;  // optional pre-split low part (same shape / ld): ptr is then the hi part (raw or pre-rounded)
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// status codes: GSMVI_OK / GSMVI_E* from the C ABI header

int make_tmap_2d(CUtensorMap* out, const MatView& v, int box_cols, int box_rows, bool atom32);

struct GemmOpts {
  int npass = 3;          // precision: 1 = TF32, 2 = 3xTF32 fast (truncation split), 3 = 3xTF32 (round-to-nearest split)
  bool a_mn = false;      // A given as [K, M]
  bool b_mn = false;      // B given as [K, N]
  float alpha = 1.0f, beta = 0.0f;
  const float* Cin = nullptr;
  long long ldcin = 0;
  const float* bias_n = nullptr;
  bool tri = false, mirror = false;
  int krange = KR_FULL;
  int neg_from = 0x7fffffff;
};

// C[M,N] = alpha * A * B^T (+ beta*Cin + bias).  Returns GSMVI_* (<0) or a positive cudaError_t.
int launch_gemm_tf32(cudaStream_t stream, int M, int N, int K, const MatView& A, const MatView& B, float* C,
                     long long ldc, const GemmOpts& o);

}  // namespace gsmvi
