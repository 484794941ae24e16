// tcgen05 scaled 3xFP16 GEMM for sm_100a:   C = alpha * op(A) * op(B)^T + beta * Cin + bias      (fp32-grade products)
//
// Same role as the 3xTF32 engine in tc_gemm.cuh at twice the tensor-pipe rate.
// A tf32 value and an fp16 value both carry an 11-bit significand, so "hi + lo" is a 22-bit split either way; what
// fp16 lacks is exponent range, and a per-tensor power-of-two scale supplies it:
//     s      = 2^(14 - e),  absmax(A) = f * 2^e  (f in [0.5, 1))            -> |A s| < 2^14, exact scaling
//     A_hi   = rn_f16(A s)
//     A_lo   = rn_f16((A s - A_hi) * 2^11)                                   -> same range as A_hi, never subnormal first
// Elements down to 2^-27 of the tensor's absmax keep the full 22 bits (hi normal above 2^-14, lo normal above 2^-25 of
// the scaled range); smaller ones degrade gracefully to an absolute error of 2^-39 absmax - far below the 2^-22
// relative error the dot product already carries from its large terms.
// Three kind::f16 MMAs per k-slice (fp16 products are exact in the fp32 accumulator):
//     corr += A_lo B_hi ;  corr += A_hi B_lo ;  main[rr] += A_hi B_hi,        C = (sum main + 2^-11 corr) / (s_a s_b)
// hi*hi is spread round-robin over three TMEM accumulators because TMEM accumulation truncates (measured, see
// tc_gemm.cuh); the epilogue adds the four in fp32 round-to-nearest.
// Operands arrive pre-split (h3_split / producer epilogues), so TMA feeds the tensor core directly: no converter
// warps, 4 bytes of shared memory per operand element instead of 8, and kind::f16 issues K=16 per instruction.
//
// Pipeline: warp 0 = TMA producer (128B swizzle, 128 x 64 fp16 tiles, hi and lo of both operands per stage, 3 stages),
// warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue (tcgen05.ld 32x32b, fused scale/alpha/beta/bias, optional
// lower-triangle tiles with mirrored stores, optional |C| max for the consumer's split, optional split-K partials).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_gemm.cuh"

namespace gsmvi {

constexpr int H3_BM = 128;
constexpr int H3_BN = 128;
constexpr int H3_BK = 64;                                // fp16 per 128-byte swizzle row
constexpr int H3_UMMA_K = 16;                            // kind::f16: 32 bytes of K per instruction
constexpr int H3_TILE_BYTES = H3_BM * H3_BK * 2;         // 16 KiB per operand part
constexpr int H3_STAGE_BYTES = 4 * H3_TILE_BYTES;        // [A_hi | B_hi | A_lo | B_lo]
constexpr int H3_STAGES = 3;
constexpr int H3_SMEM_BYTES = 1024 + BAR_BYTES + H3_STAGES * H3_STAGE_BYTES;
constexpr int H3_N_MAIN = 3;
constexpr int H3_TMEM_COLS = 512;
constexpr int H3_CORR_COL = H3_N_MAIN * H3_BN;
constexpr int H3_THREADS = 320;
constexpr float H3_LO_SCALE = 2048.0f;                   // 2^11

struct H3Args {
  int M, N, K;
  float alpha, beta;
  const float* Cin;  // may alias C; ignored when beta == 0
  long long ldcin;
  float* C;
  long long ldc;
  const float* bias_n;
  const float* scale_a;  // device scalars: the power-of-two scales the operands were stored with
  const float* scale_b;
  unsigned* absmax_out;  // optional: atomicMax of the bit patterns of |C|
  int tri, mirror, krange;
  int tiles_m, tiles_n;
  int splits;              // split-K: blockIdx.y = split index s, output goes to C + s * split_stride (raw partials)
  long long split_stride;
  // push mode (multi-GPU reduce-scatter fused into the epilogue; tri only): lower tile t belongs to rank t % push_world and is
  // written, as a dense 128 x 128 tile, straight into the owner's staging area over NVLink peer memory
  //   push_base[owner] + push_stage_off + ((push_rank * push_tpo + t / push_world) * 128 * 128)
  // followed by a system-scope release increment of the owner's per-tile arrival counter (4-byte words at push_cnt_off).
  float* const* push_base;
  long long push_stage_off, push_cnt_off;
  int push_rank, push_world, push_tpo;
  // fused split: also write C as the fp16 pair (split_hi, split_lo; leading dimension split_ld) with the GIVEN scale
  // *split_scale - for results whose magnitude is bounded a priori, so no max|C| pass and no separate split kernel
  __half* split_hi;
  __half* split_lo;
  long long split_ld;
  const float* split_scale;
};

__host__ __device__ constexpr uint32_t make_idesc_f16(bool a_mn, bool b_mn) {
  // c_format F32 (1) [4,6); a_format F16 (0) [7,10); b_format F16 (0) [10,13); a_major 15; b_major 16;
  // N>>3 [17,23); M>>4 [24,29)
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | (static_cast<uint32_t>(H3_BN >> 3) << 17) |
         (static_cast<uint32_t>(H3_BM >> 4) << 24);
}

namespace ptx {
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
}  // namespace ptx

__device__ __forceinline__ void h3_split1(float v, float s, __half& hi, __half& lo);

// The whole CTA program of one (tile, split) work item: bx = tile index, by = split index (the x / y block indices of a
// plain launch).  A device function so that the Cholesky's fused panel kernel (potrf_h3.cu) can run it in its otherwise
// idle CTAs; every one of the CTA's H3_THREADS threads must call it.  The tensor maps must live in kernel parameter
// space (__grid_constant__).
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void gemm_h3_body(const H3Args& args, const CUtensorMap& tmAhi, const CUtensorMap& tmBhi,
                                             const CUtensorMap& tmAlo, const CUtensorMap& tmBlo, uint8_t* smem_raw,
                                             const int bx, const int by, const bool last_item = true) {
  constexpr int STAGES = H3_STAGES;
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t bar_base = ptx::smem_u32(smem);
  const uint32_t stage_base = bar_base + BAR_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 8 * (2 * STAGES + 1) + 8);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int tm, tn;
  if (args.tri) {
    const int t = bx;
    int i = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    tm = i;
    tn = t - i * (i + 1) / 2;
  } else {
    constexpr int GROUP = 8;  // tile-rows per group: concurrently resident CTAs share operand tiles in L2
    const int t = bx;
    const int per_group = GROUP * args.tiles_n;
    const int g = t / per_group;
    const int first_m = g * GROUP;
    const int rows = min(GROUP, args.tiles_m - first_m);
    const int r = t - g * per_group;
    tm = first_m + r % rows;
    tn = r / rows;
    // triangular operand: K length grows with the tile index, so hand out the long tiles first (no long tail at the end)
    if (args.krange & KR_B_LOWER) tn = args.tiles_n - 1 - tn;
    if (args.krange & KR_A_LOWER) tm = args.tiles_m - 1 - tm;
  }
  const int m0 = tm * H3_BM, n0 = tn * H3_BN;

  // ---- K range (whole BK blocks), then this CTA's share of it (split-K)
  int k_begin = 0, k_end = args.K;
  if (args.krange & KR_A_LOWER) k_end = min(k_end, m0 + H3_BM);
  if (args.krange & KR_B_LOWER) k_end = min(k_end, n0 + H3_BN);
  if (args.krange & KR_A_UPPER) k_begin = max(k_begin, m0);
  if (args.krange & KR_B_UPPER) k_begin = max(k_begin, n0);
  int kb_begin = k_begin / H3_BK;
  int kb_end = (k_end > k_begin) ? (k_end + H3_BK - 1) / H3_BK : kb_begin;
  const int split = by;
  if (args.splits > 1) {
    const int total = kb_end - kb_begin;
    const int per = (total + args.splits - 1) / args.splits;
    kb_begin = min(kb_end, kb_begin + split * per);
    kb_end = min(kb_end, kb_begin + per);
  }
  const int num_kb = kb_end - kb_begin;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmAhi);
    ptx::prefetch_tmap(&tmBhi);
    ptx::prefetch_tmap(&tmAlo);
    ptx::prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), H3_TMEM_COLS);
    if (last_item) ptx::tmem_relinquish();  // (a CTA that runs the body again must keep its permit to allocate)
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream; from here
  // on its results are needed (no-ops when the kernel was launched without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (the warp stays converged, one elected lane issues) =====================
    for (int it = 0; it < num_kb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      const int k0 = (kb_begin + it) * H3_BK;
      const uint32_t sA = stage_base + s * H3_STAGE_BYTES;
      const uint32_t sB = sA + H3_TILE_BYTES;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar(s), H3_STAGE_BYTES);
        if (!A_MN) {
          ptx::tma_load_2d(sA, &tmAhi, full_bar(s), k0, m0);
          ptx::tma_load_2d(sA + 2 * H3_TILE_BYTES, &tmAlo, full_bar(s), k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < H3_BM / 64; ++c) {
            ptx::tma_load_2d(sA + c * (H3_BK * 128), &tmAhi, full_bar(s), m0 + 64 * c, k0);
            ptx::tma_load_2d(sA + 2 * H3_TILE_BYTES + c * (H3_BK * 128), &tmAlo, full_bar(s), m0 + 64 * c, k0);
          }
        }
        if (!B_MN) {
          ptx::tma_load_2d(sB, &tmBhi, full_bar(s), k0, n0);
          ptx::tma_load_2d(sB + 2 * H3_TILE_BYTES, &tmBlo, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < H3_BN / 64; ++c) {
            ptx::tma_load_2d(sB + c * (H3_BK * 128), &tmBhi, full_bar(s), n0 + 64 * c, k0);
            ptx::tma_load_2d(sB + 2 * H3_TILE_BYTES + c * (H3_BK * 128), &tmBlo, full_bar(s), n0 + 64 * c, k0);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the warp stays converged, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc_f16(A_MN, B_MN);
    // K-major (SWIZZLE_128B): rows of 64 fp16, 8-row groups 1024 B apart (SBO); a k-slice (16 fp16) is 32 B along the row.
    // MN-major (SWIZZLE_128B): 64-wide MN chunks H3_BK*128 B apart (LBO), 8-deep K groups 1024 B apart (SBO); a k-slice
    // (16 K rows) is two groups = 2048 B further.
    constexpr uint32_t A_LBO = A_MN ? H3_BK * 128 : 16, A_SBO = 1024, A_KSTEP = A_MN ? 2048 : H3_UMMA_K * 2;
    constexpr uint32_t B_LBO = B_MN ? H3_BK * 128 : 16, B_SBO = 1024, B_KSTEP = B_MN ? 2048 : H3_UMMA_K * 2;
    for (int it = 0; it < num_kb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after_sync();
      const uint32_t sA = stage_base + s * H3_STAGE_BYTES;
      const uint32_t sB = sA + H3_TILE_BYTES;
      const uint32_t t_main = tmem_base + (it % H3_N_MAIN) * H3_BN;
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < H3_BK / H3_UMMA_K; ++kk) {
          const uint64_t da = make_smem_desc(sA + kk * A_KSTEP, A_LBO, A_SBO, 2);
          const uint64_t db = make_smem_desc(sB + kk * B_KSTEP, B_LBO, B_SBO, 2);
          const uint64_t da_lo = make_smem_desc(sA + 2 * H3_TILE_BYTES + kk * A_KSTEP, A_LBO, A_SBO, 2);
          const uint64_t db_lo = make_smem_desc(sB + 2 * H3_TILE_BYTES + kk * B_KSTEP, B_LBO, B_SBO, 2);
          ptx::umma_f16(tmem_base + H3_CORR_COL, da_lo, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          ptx::umma_f16(tmem_base + H3_CORR_COL, da, db_lo, idesc, 1u);
          ptx::umma_f16(t_main, da, db, idesc, (it >= H3_N_MAIN || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar(s));
        if (it + 1 == num_kb) ptx::umma_commit(acc_bar);
      }
      __syncwarp();
    }
    if (num_kb == 0 && ptx::elect_one()) ptx::mbar_arrive(acc_bar);
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const float sa = *args.scale_a, sb = *args.scale_b;
    const float alpha = (args.alpha / sa) / sb, beta = args.beta;
    const float ssplit = args.split_hi ? *args.split_scale : 0.0f;
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after_sync();
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;  // which 64 columns
    const int m = m0 + q * 32 + lane;
    const bool diag_tile = args.tri && (tm == tn);
    float* Cout = args.C + static_cast<long long>(split) * args.split_stride;
    unsigned amax = 0u;  // bit pattern of max |o|: as unsigned, NaN > Inf > finite, so a poisoned tile is not dropped
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int c0 = half * 64 + chunk * 32;
      uint32_t r[32];
      if (num_kb > 0) {
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0;
        ptx::tmem_ld_32x32(t0, r);
        ptx::tmem_ld_wait();
        const int n_main = min(num_kb, H3_N_MAIN);
        uint32_t t[32];
        for (int a = 1; a < n_main; ++a) {
          ptx::tmem_ld_32x32(t0 + a * H3_BN, t);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
        }
        ptx::tmem_ld_32x32(t0 + H3_CORR_COL, t);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(fmaf(__uint_as_float(t[j]), 1.0f / H3_LO_SCALE, __uint_as_float(r[j])));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int nbase = n0 + c0;
      if (args.push_base != nullptr) {
        // stage the scaled tile in shared memory (the pipeline stages are idle: every MMA has retired), row stride 132
        // floats = conflict-free float4 stores; it leaves for the owner rank as 128 bulk copies of one 512-byte row each
        float* srow = reinterpret_cast<float*>(smem + BAR_BYTES) + (q * 32 + lane) * 132 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(srow + j) =
              make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]), alpha * __uint_as_float(r[j + 2]),
                          alpha * __uint_as_float(r[j + 3]));
      } else if (m < args.M && nbase < args.N) {
        float* crow = Cout + static_cast<long long>(m) * args.ldc + nbase;
        const float* cin = (beta != 0.0f) ? args.Cin + static_cast<long long>(m) * args.ldcin + nbase : nullptr;
        const bool vec_ok = !diag_tile && (nbase + 32 <= args.N) && ((args.ldc & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) &&
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
            if (args.split_hi) {
              __half h[4], l[4];
              h3_split1(o.x, ssplit, h[0], l[0]);
              h3_split1(o.y, ssplit, h[1], l[1]);
              h3_split1(o.z, ssplit, h[2], l[2]);
              h3_split1(o.w, ssplit, h[3], l[3]);
              const long long so = static_cast<long long>(m) * args.split_ld + nbase + j;
              *reinterpret_cast<uint2*>(args.split_hi + so) = *reinterpret_cast<const uint2*>(h);
              *reinterpret_cast<uint2*>(args.split_lo + so) = *reinterpret_cast<const uint2*>(l);
            }
            amax = max(max(amax, __float_as_uint(fabsf(o.x))), max(__float_as_uint(fabsf(o.y)), max(__float_as_uint(fabsf(o.z)), __float_as_uint(fabsf(o.w)))));
            r[j + 0] = __float_as_uint(o.x); r[j + 1] = __float_as_uint(o.y);
            r[j + 2] = __float_as_uint(o.z); r[j + 3] = __float_as_uint(o.w);
          }
          if (args.mirror) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              Cout[static_cast<long long>(nbase + j) * args.ldc + m] = __uint_as_float(r[j]);
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
              if (args.split_hi) {
                const long long so = static_cast<long long>(m) * args.split_ld + n;
                h3_split1(o, ssplit, args.split_hi[so], args.split_lo[so]);
              }
              amax = max(amax, __float_as_uint(fabsf(o)));
              if (args.mirror && n != m) Cout[static_cast<long long>(n) * args.ldc + m] = o;
            }
          }
        }
      }
    }
    if (args.absmax_out) {
      const unsigned bits = __reduce_max_sync(0xffffffffu, amax);
      if (lane == 0 && bits != 0u) atomicMax(args.absmax_out, bits);
    }
    if (args.push_base != nullptr) {
      // the tile is complete in shared memory: 128 threads send one row each (cp.async.bulk, shared -> peer global:
      // full-size NVLink packets instead of 16-byte scattered stores), wait for their copies, then one thread
      // publishes the tile to its owner
      const int t = bx, owner = t % args.push_world;
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int prow = threadIdx.x - 64;
      if (prow < H3_BM) {
        float* tile = args.push_base[owner] + args.push_stage_off +
                      (static_cast<long long>(args.push_rank) * args.push_tpo + t / args.push_world) * (H3_BM * H3_BN);
        const uint32_t src = bar_base + BAR_BYTES + prow * 132 * 4;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(tile + prow * H3_BN), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        unsigned* cnt = reinterpret_cast<unsigned*>(args.push_base[owner]) + args.push_cnt_off + t / args.push_world;
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence_system();
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
      }
    }
    ptx::tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, H3_TMEM_COLS);
  }
  if (warp == 0 && lane == 0) {
    // the Cholesky's fused kernel may run this body again in the same CTA (potrf_h3.cu): leave no valid mbarrier behind
    for (int s = 0; s < 2 * STAGES + 1; ++s) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar_base + 8u * s) : "memory");
  }
  __syncthreads();
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(H3_THREADS, 1)
gemm_h3_kernel(const H3Args args, const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo) {
  extern __shared__ uint8_t h3_smem_raw[];
  gemm_h3_body<A_MN, B_MN>(args, tmAhi, tmBhi, tmAlo, tmBlo, h3_smem_raw, blockIdx.x, blockIdx.y);
}

// ------------------------------------------------------------------------------------------------ host side

// Pre-split operand: `rows` x `cols` fp16 pair (hi, lo), row-major with leading dimension ld (elements, multiple of 8),
// stored with the power-of-two scale *scale (device float).  K-major: rows = M (or N), cols = K.  MN-major: rows = K.
struct HView {
  const __half* hi;
  const __half* lo;
  long long rows, cols, ld;
  const float* scale;
};

struct H3Opts {
  bool a_mn = false, b_mn = false;
  float alpha = 1.0f, beta = 0.0f;
  const float* Cin = nullptr;
  long long ldcin = 0;
  const float* bias_n = nullptr;
  bool tri = false, mirror = false;
  int krange = KR_FULL;
  unsigned* absmax_out = nullptr;
  int splits = 1;
  long long split_stride = 0;
  bool pdl = false;                   // launch with programmatic stream serialization (prologue overlaps the previous kernel)
  __half* split_hi = nullptr;         // fused split of the result with a given scale, see H3Args
  __half* split_lo = nullptr;
  long long split_ld = 0;
  const float* split_scale = nullptr;
  float* const* push_base = nullptr;  // push mode, see H3Args
  long long push_stage_off = 0, push_cnt_off = 0;
  int push_rank = 0, push_world = 1, push_tpo = 0;
};

int launch_gemm_h3(cudaStream_t stream, int M, int N, int K, const HView& A, const HView& B, float* C, long long ldc,
                   const H3Opts& o);
// Select the kernel behind launch_gemm_h3 for the launches both can express: 1 = the persistent 2-CTA kernel of
// h3x2_gemm.cuh (default; GSMVI_H3X2=0 in the environment starts with 0), 0 = the one-CTA kernel above, anything else = query.
// Returns the previous setting.
int h3_pair_kernel(int enable);
// The argument block, the four tensor maps (A_hi, B_hi, A_lo, B_lo) and the grid (tiles, splits) of that launch, for
// callers that run gemm_h3_body inside a kernel of their own (potrf_h3.cu).
// fp16 tensor map (SWIZZLE_128B) of a [rows x cols] view with leading dimension ld and the given box
int h3_make_tmap(CUtensorMap* out, const __half* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows);
int h3_prepare(int M, int N, int K, const HView& A, const HView& B, float* C, long long ldc, const H3Opts& o, H3Args* out,
               CUtensorMap* maps, dim3* grid_out);

// scale <- power of two with absmax * scale in [2^13, 2^14) (absmax given as the bit pattern of a non-negative float,
// or of its square when sqrt_mode: |L_ij| <= sqrt(max Sigma_ii)); Hi = rn_f16(A scale), Lo = rn_f16((A scale - Hi) 2^11).
int h3_split(cudaStream_t stream, const float* A, long long lda, int rows, int cols, const unsigned* absmax, int sqrt_mode,
             float* scale_out, __half* Hi, __half* Lo, long long ldo);
// *out <- max(*out, max |A|) as a float bit pattern (zero *out first).
int h3_absmax(cudaStream_t stream, const float* A, long long lda, int rows, int cols, unsigned* out);

// Scales for results bounded a priori (fused split in the producing GEMM's epilogue):
//   |x_i| <= max|mu| + zmax sqrt(D) sqrt(max Sigma_ii)   (Cauchy-Schwarz on x = mu + L z, |z_k| <= zmax),
//   |g_j| <= xbound * pnorm + cmax                        (g = -x P + c, pnorm = max_j sum_k |P_kj|).
// The bounds are loose by up to 2^8.5 / 2^15; the fp16 pair keeps an absolute precision of 2^-35 of the scaled range
// (hi and lo subnormals included), i.e. >= 2^-20 relative to the typical entry even then.
// zmax_bits: device word with the bit pattern of max|z| (or null to use zmax_const); sigma_absmax: bit pattern of max|Sigma|.
int h3_bound_scales(cudaStream_t stream, const float* mu, int D, const unsigned* sigma_absmax, const unsigned* zmax_bits,
                    float zmax_const, float pnorm, float cmax, float* scale_x, float* scale_g);

__device__ __forceinline__ float h3_scale_from_absmax(unsigned bits, int sqrt_mode) {
  float m = __uint_as_float(bits);
  if (sqrt_mode) m = sqrtf(m);
  if (!(m > 0.0f) || !(m < 3.0e38f)) return 1.0f;  // zero, NaN or Inf: keep the data as it is (NaN/Inf then propagate)
  int e;
  frexpf(m, &e);
  return ldexpf(1.0f, 14 - e);
}
__device__ __forceinline__ void h3_split1(float v, float s, __half& hi, __half& lo) {
  const float x = v * s;
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * H3_LO_SCALE);
}

}  // namespace gsmvi
