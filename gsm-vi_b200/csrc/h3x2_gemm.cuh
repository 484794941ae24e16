// Persistent 2-CTA form of the scaled 3xFP16 GEMM (h3_gemm.cuh):  C = alpha * op(A) * op(B)^T + beta * Cin + bias.
//
// Same arithmetic, operand format and epilogue options as gemm_h3_kernel; what changes is how the tile is fed and drained.
// The one-CTA kernel moves 64 KiB from L2 into shared memory per 768 tensor-pipe cycles (85 B/clk/SM at full rate: ncu shows
// 12.6-14.5 TB/s of xbar->L1 traffic, the tensor pipe waiting on it a third of the time) and drains its four TMEM
// accumulators with the tensor core idle.  Here
//   * two CTAs of one TPC form a pair (cluster 2x1x1) and compute a 256 x 128 tile with tcgen05.mma.cta_group::2: each CTA
//     loads its own 128 rows of A and HALF of the B tile (64 rows), the tensor cores read the other half from the peer's
//     shared memory: 48 KiB per CTA per k-block instead of 64, four pipeline stages instead of three;
//   * the kernel is persistent (one pair per TPC, static snake-ordered tile list), so barriers, TMEM and the TMA pipeline are
//     set up once and the producer runs ahead across tile boundaries;
//   * TMEM holds two main accumulators (hi*hi) and two correction accumulators (hi*lo + lo*hi) of 128 columns each.  The MMA
//     thread fills one main accumulator with a CHUNK of 8 k-blocks (K = 512) while the epilogue warps add the other one to
//     fp32 registers (round-to-nearest): the running sum lives in registers, so TMEM's truncating (one-sided) accumulation
//     only ever spans 32 additions - its error grows linearly with the number of additions into one accumulator, measured
//     7.8e-6 sqrt(K) for 64 additions against 85 per accumulator in the one-CTA kernel at K = 4096.  The correction terms are
//     2^-11 of the main ones, so their accumulator runs over the whole K and alternates per TILE; it is read together
//     with the tile's last chunk, and the store of a tile overlaps the next tile's first chunks.
// Serves the four batch-sized contractions of a GSM step (gsmvi/gsm.py:11-27, 53-54, 119; examples/example_gsm_numpy.py:24-29).
// Push mode (PUSH = true; the reduce-scatter of the batch-sharded covariance update, csrc/comm.cu): the epilogue stages the
// scaled tile in shared memory and sends it to its owner rank's staging slot as 128 bulk copies of one 512-byte row, then
// release-increments the owner's arrival counter - all under the next tile's MMAs.
// Not handled here (launch_gemm_h3 falls back to the one-CTA kernel): split-K partials, KR_A_* / KR_B_UPPER ranges.
#pragma once
#include "h3_gemm.cuh"

namespace gsmvi {

constexpr int X2_STAGES = 4;
constexpr int X2_A_BYTES = 128 * H3_BK * 2;                        // one part (hi or lo) of this CTA's 128 rows of A
constexpr int X2_B_BYTES = 64 * H3_BK * 2;                         // one part of this CTA's 64 rows of B
constexpr int X2_STAGE_BYTES = 2 * X2_A_BYTES + 2 * X2_B_BYTES;    // [A_hi | A_lo | B_hi | B_lo] = 48 KiB
constexpr int X2_SMEM_BYTES = 1024 + BAR_BYTES + X2_STAGES * X2_STAGE_BYTES;
// push mode (multi-GPU covariance update): three stages + a 128 x 132-float staging tile for the bulk copies to the owner rank
constexpr int X2_PUSH_STAGES = 3;
constexpr int X2_PUSH_TILE_BYTES = 128 * 132 * 4;
constexpr int X2_PUSH_SMEM_BYTES = 1024 + BAR_BYTES + X2_PUSH_STAGES * X2_STAGE_BYTES + X2_PUSH_TILE_BYTES;
constexpr int X2_CHUNK_KB = 8;                                     // default k-blocks per TMEM accumulation chunk of hi*hi
constexpr int X2_THREADS = 320;
constexpr int X2_CORR_COL = 256;                                   // TMEM columns: main0 | main1 | corr0 | corr1

__host__ __device__ constexpr uint32_t make_idesc_f16_x2(bool a_mn, bool b_mn) {
  // as make_idesc_f16 with M = 256 (both CTAs), N = 128
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | (static_cast<uint32_t>(128 >> 3) << 17) |
         (static_cast<uint32_t>(256 >> 4) << 24);
}

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_x2(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_x2(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_x2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_x2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_x2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_x2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
}  // namespace ptx

// Tile list of the pair kernel.  Plain: 256-row super-rows x 128-column tiles, eight... four super-rows per group (the pairs
// that run concurrently share operand tiles in L2); with a lower-triangular B operand the K length grows with the column, so
// the columns are handed out longest first.  tri: super-row i needs the columns j <= 2i+1 (its second CTA's diagonal tile).
struct X2Tile {
  int tmp, tn;
};
__device__ __forceinline__ X2Tile x2_tile(const H3Args& args, int tiles_mp, int pos) {
  X2Tile t;
  if (args.tri) {
    int i = static_cast<int>((sqrtf(4.0f * pos + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) <= pos) ++i;
    while (i * (i + 1) > pos) --i;
    t.tmp = i;
    t.tn = pos - i * (i + 1);
  } else if (args.krange & KR_B_LOWER) {
    t.tn = args.tiles_n - 1 - pos / tiles_mp;
    t.tmp = pos % tiles_mp;
  } else {
    constexpr int GROUP = 4;
    const int per_group = GROUP * args.tiles_n;
    const int g = pos / per_group;
    const int first = g * GROUP;
    const int rows = min(GROUP, tiles_mp - first);
    const int r = pos - g * per_group;
    t.tmp = first + r % rows;
    t.tn = r / rows;
  }
  return t;
}

// one 32-column group of a finished tile row: registers -> global (the store half of gemm_h3_body's epilogue)
__device__ __forceinline__ void x2_store32(const H3Args& args, float (&v)[32], const int m, const int nbase, const bool diag_tile,
                                           const float alpha, const float beta, const float ssplit, unsigned& amax) {
  if (!(m < args.M && nbase < args.N)) return;
  float* Cout = args.C;
  float* crow = Cout + static_cast<long long>(m) * args.ldc + nbase;
  const float* cin = (beta != 0.0f) ? args.Cin + static_cast<long long>(m) * args.ldcin + nbase : nullptr;
  const bool vec_ok = !diag_tile && (nbase + 32 <= args.N) && ((args.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) &&
                      (cin == nullptr || (((args.ldcin & 3) == 0) && ((reinterpret_cast<uintptr_t>(args.Cin) & 15) == 0)));
  if (vec_ok) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o;
      o.x = alpha * v[j + 0];
      o.y = alpha * v[j + 1];
      o.z = alpha * v[j + 2];
      o.w = alpha * v[j + 3];
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
      v[j + 0] = o.x; v[j + 1] = o.y; v[j + 2] = o.z; v[j + 3] = o.w;
    }
    if (args.mirror) {
#pragma unroll
      for (int j = 0; j < 32; ++j) Cout[static_cast<long long>(nbase + j) * args.ldc + m] = v[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = nbase + j;
      if (n < args.N && !(diag_tile && n > m)) {
        float o = alpha * v[j];
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

template <bool A_MN, bool B_MN, bool PUSH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(X2_THREADS, 1)
gemm_h3x2_kernel(const H3Args args, const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmBhi,
                 const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, const int tiles_mp,
                 const int n_st, const int chunk_kb, const int probe) {
  extern __shared__ uint8_t x2_smem_raw[];
  constexpr int STAGES = PUSH ? X2_PUSH_STAGES : X2_STAGES;
  const uint32_t raw_addr = ptx::smem_u32(x2_smem_raw);
  uint8_t* smem = x2_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t bar_base = ptx::smem_u32(smem);
  const uint32_t stage_base = bar_base + BAR_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                    // used in the leader CTA only
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };        // per CTA, signalled by multicast commits
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };    // per CTA, multicast commits
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };  // leader CTA only: 16 epilogue warps arrive
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 8 * (2 * STAGES + 4) + 8);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for a 2x1x1 cluster, and uniform to the compiler
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmAhi);
    ptx::prefetch_tmap(&tmBhi);
    ptx::prefetch_tmap(&tmAlo);
    ptx::prefetch_tmap(&tmBlo);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 16);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_x2(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    ptx::tmem_relinquish_x2();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion can reach them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // every role walks the same tile list: round i, snake order over the pairs (equalises the K lengths of a triangular operand)
  auto tile_pos = [&](int i) { return i * n_pairs + ((i & 1) ? n_pairs - 1 - pair : pair); };
  auto tile_kb = [&](const X2Tile& t) {
    int k_end = args.K;
    if (args.krange & KR_B_LOWER) k_end = min(k_end, t.tn * H3_BN + H3_BN);
    return (k_end + H3_BK - 1) / H3_BK;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; the warp stays converged, one elected lane issues) ================
    {
      uint32_t it = 0;
      for (int i = 0; i * n_pairs < n_st; ++i) {
        const int pos = tile_pos(i);
        if (pos >= n_st) continue;
        const X2Tile t = x2_tile(args, tiles_mp, pos);
        if (args.tri && t.tn >= args.tiles_n) continue;
        const int m0 = (2 * t.tmp + static_cast<int>(rank)) * 128;
        const int n0 = t.tn * H3_BN + static_cast<int>(rank) * 64;
        const int num_kb = tile_kb(t);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t fb = ptx::mapa(full_bar(s), 0);
          const int k0 = kb * H3_BK;
          const uint32_t sAh = stage_base + s * X2_STAGE_BYTES;
          const uint32_t sAl = sAh + X2_A_BYTES;
          const uint32_t sBh = sAl + X2_A_BYTES;
          const uint32_t sBl = sBh + X2_B_BYTES;
          if (ptx::elect_one()) {
            if (probe & 1) {  // bottleneck probe (results are garbage): no TMA traffic at all, the stage is declared full as is
              if (rank == 0) ptx::mbar_arrive(full_bar(s));
            } else {
              if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(s), 2 * X2_STAGE_BYTES);
              if (!A_MN) {
                ptx::tma_load_2d_x2(sAh, &tmAhi, fb, k0, m0);
                ptx::tma_load_2d_x2(sAl, &tmAlo, fb, k0, m0);
              } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  ptx::tma_load_2d_x2(sAh + c * (H3_BK * 128), &tmAhi, fb, m0 + 64 * c, k0);
                  ptx::tma_load_2d_x2(sAl + c * (H3_BK * 128), &tmAlo, fb, m0 + 64 * c, k0);
                }
              }
              if (!B_MN) {
                ptx::tma_load_2d_x2(sBh, &tmBhi, fb, k0, n0);
                ptx::tma_load_2d_x2(sBl, &tmBlo, fb, k0, n0);
              } else {
                ptx::tma_load_2d_x2(sBh, &tmBhi, fb, n0, k0);
                ptx::tma_load_2d_x2(sBl, &tmBlo, fb, n0, k0);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; the warp stays converged, one elected lane issues) =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16_x2(A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? H3_BK * 128 : 16, A_SBO = 1024, A_KSTEP = A_MN ? 2048 : H3_UMMA_K * 2;
      constexpr uint32_t B_LBO = B_MN ? H3_BK * 128 : 16, B_SBO = 1024, B_KSTEP = B_MN ? 2048 : H3_UMMA_K * 2;
      uint32_t it = 0, nchunk = 0, ntile = 0;
      for (int i = 0; i * n_pairs < n_st; ++i) {
        const int pos = tile_pos(i);
        if (pos >= n_st) continue;
        const X2Tile t = x2_tile(args, tiles_mp, pos);
        if (args.tri && t.tn >= args.tiles_n) continue;
        const int num_kb = tile_kb(t);
        // corr accumulator of this tile: the one read at the end of the tile before last, and that read preceded the
        // tempty arrival of that tile's last chunk, which the wait below has seen by the time the accumulator is reused
        const uint32_t t_corr = tmem_base + X2_CORR_COL + (ntile & 1u) * H3_BN;
        ++ntile;
        for (int kb0 = 0; kb0 < num_kb; kb0 += chunk_kb, ++nchunk) {
          const uint32_t buf = nchunk & 1u, use = nchunk >> 1;
          ptx::mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);  // both CTAs' epilogue warps have read this main accumulator
          ptx::tc_fence_after_sync();
          const uint32_t t_main = tmem_base + buf * H3_BN;
          const int kbe = min(num_kb, kb0 + chunk_kb);
          for (int kb = kb0; kb < kbe; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            ptx::mbar_wait(full_bar(s), ph);
            ptx::tc_fence_after_sync();
            const uint32_t sAh = stage_base + s * X2_STAGE_BYTES;
            const uint32_t sAl = sAh + X2_A_BYTES;
            const uint32_t sBh = sAl + X2_A_BYTES;
            const uint32_t sBl = sBh + X2_B_BYTES;
            if (ptx::elect_one()) {
              // hi*hi of the whole k-block first, then its eight correction MMAs: the accumulator the tensor core works
              // on changes twice per k-block instead of eight times
#pragma unroll
              for (int kk = 0; kk < H3_BK / H3_UMMA_K; ++kk) {
                const uint64_t da = make_smem_desc(sAh + kk * A_KSTEP, A_LBO, A_SBO, 2);
                const uint64_t db = make_smem_desc(sBh + kk * B_KSTEP, B_LBO, B_SBO, 2);
                ptx::umma_f16_x2(t_main, da, db, idesc, (kb == kb0 && kk == 0) ? 0u : 1u);
              }
              if (!(probe & 2)) {  // bottleneck probe: hi*hi only (a third of the MMAs on the same operand traffic)
#pragma unroll
                for (int kk = 0; kk < H3_BK / H3_UMMA_K; ++kk) {
                  const uint64_t da = make_smem_desc(sAh + kk * A_KSTEP, A_LBO, A_SBO, 2);
                  const uint64_t db = make_smem_desc(sBh + kk * B_KSTEP, B_LBO, B_SBO, 2);
                  const uint64_t da_lo = make_smem_desc(sAl + kk * A_KSTEP, A_LBO, A_SBO, 2);
                  const uint64_t db_lo = make_smem_desc(sBl + kk * B_KSTEP, B_LBO, B_SBO, 2);
                  ptx::umma_f16_x2(t_corr, da_lo, db, idesc, (kb == 0 && kk == 0) ? 0u : 1u);
                  ptx::umma_f16_x2(t_corr, da, db_lo, idesc, 1u);
                }
              }
              ptx::umma_commit_x2(empty_bar(s));  // the stage is free in both CTAs once these MMAs have read it
              if (kb + 1 == kbe) ptx::umma_commit_x2(tfull_bar(buf));
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): TMEM chunks -> fp32 registers -> global =====================
    const float sa = *args.scale_a, sb = *args.scale_b;
    const float alpha = (args.alpha / sa) / sb, beta = args.beta;
    const float ssplit = args.split_hi ? *args.split_scale : 0.0f;
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;  // which 64 columns
    const uint32_t tempty_leader0 = ptx::mapa(tempty_bar(0), 0);
    const uint32_t tempty_leader1 = ptx::mapa(tempty_bar(1), 0);
    unsigned amax = 0u;
    uint32_t nchunk = 0, ntile = 0;
    for (int i = 0; i * n_pairs < n_st; ++i) {
      const int pos = tile_pos(i);
      if (pos >= n_st) continue;
      const X2Tile t = x2_tile(args, tiles_mp, pos);
      if (args.tri && t.tn >= args.tiles_n) continue;
      const int tm = 2 * t.tmp + static_cast<int>(rank);
      const int num_kb = tile_kb(t);
      // not wanted: the tile lies above the diagonal (second-row-only supertile) or below the matrix (odd tile-row count)
      const bool wanted = tm < args.tiles_m && !(args.tri && t.tn > tm);
      const uint32_t lane_col = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
      const uint32_t t_corr = lane_col + X2_CORR_COL + (ntile & 1u) * H3_BN;
      ++ntile;
      float acc0[32], acc1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc0[j] = acc1[j] = 0.0f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += chunk_kb, ++nchunk) {
        const uint32_t buf = nchunk & 1u, use = nchunk >> 1;
        ptx::mbar_wait(tfull_bar(buf), use & 1u);
        ptx::tc_fence_after_sync();
        if (wanted) {
          const uint32_t t0 = lane_col + buf * H3_BN;
          uint32_t r[32], c[32];
          ptx::tmem_ld_32x32(t0, r);
          ptx::tmem_ld_32x32(t0 + 32, c);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc0[j] += __uint_as_float(r[j]);
            acc1[j] += __uint_as_float(c[j]);
          }
          if (kb0 + chunk_kb >= num_kb) {  // last chunk of the tile: every MMA of the tile has retired, add the corrections
            ptx::tmem_ld_32x32(t_corr, r);
            ptx::tmem_ld_32x32(t_corr + 32, c);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc0[j] = fmaf(__uint_as_float(r[j]), 1.0f / H3_LO_SCALE, acc0[j]);
              acc1[j] = fmaf(__uint_as_float(c[j]), 1.0f / H3_LO_SCALE, acc1[j]);
            }
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
      }
      if (wanted && PUSH) {
        // stage the scaled tile (row stride 132 floats: conflict-free float4 stores), then 128 threads send one 512-byte row
        // each to the owner rank (cp.async.bulk, shared -> peer global: full-size NVLink packets), wait for their copies,
        // and one thread publishes the tile; the staging tile is free again after the second barrier
        float* stile = reinterpret_cast<float*>(smem + BAR_BYTES + STAGES * X2_STAGE_BYTES);
        float* srow = stile + (q * 32 + lane) * 132 + half * 64;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          *reinterpret_cast<float4*>(srow + j) = make_float4(alpha * acc0[j], alpha * acc0[j + 1], alpha * acc0[j + 2], alpha * acc0[j + 3]);
          *reinterpret_cast<float4*>(srow + 32 + j) = make_float4(alpha * acc1[j], alpha * acc1[j + 1], alpha * acc1[j + 2], alpha * acc1[j + 3]);
        }
        const int tl = tm * (tm + 1) / 2 + t.tn, owner = tl % args.push_world;
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int prow = threadIdx.x - 64;
        if (prow < 128) {
          float* tile = args.push_base[owner] + args.push_stage_off +
                        (static_cast<long long>(args.push_rank) * args.push_tpo + tl / args.push_world) * (128 * 128);
          const uint32_t src = ptx::smem_u32(stile + prow * 132);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(tile + prow * 128), "r"(src) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
          unsigned* cnt = reinterpret_cast<unsigned*>(args.push_base[owner]) + args.push_cnt_off + tl / args.push_world;
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence_system();
          asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
        }
      } else if (wanted) {
        const int m = tm * 128 + q * 32 + lane;
        const int nbase = t.tn * H3_BN + half * 64;
        const bool diag_tile = args.tri && (tm == t.tn);
        x2_store32(args, acc0, m, nbase, diag_tile, alpha, beta, ssplit, amax);
        x2_store32(args, acc1, m, nbase + 32, diag_tile, alpha, beta, ssplit, amax);
      }
    }
    if (args.absmax_out) {
      const unsigned bits = __reduce_max_sync(0xffffffffu, amax);
      if (lane == 0 && bits != 0u) atomicMax(args.absmax_out, bits);
    }
    ptx::tc_fence_before_sync();
  }

  // nobody leaves while the peer may still read this CTA's shared memory, arrive on its barriers or use the paired TMEM
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_x2(tmem_base, 512);
  }
}

}  // namespace gsmvi
