// Multi-GPU exchange of the GSM batch statistics over NVLink peer memory, fused with the
// covariance-update GEMM instead of an NCCL all-reduce after it.  What is exchanged are the two batch means of
// gsm_update (gsmvi/gsm.py:53-54: mean over samples of mu_update and S_update), each rank holding B / world samples.
//
// One process per GPU; every rank owns a "comm buffer" (cudaMalloc + CUDA IPC, mapped by all peers):
//     [ staging: world x tpo dense 128x128 tiles | Sigma buffer 0 | Sigma buffer 1 | dmu: world x ld | counters ]
// Lower tile t of the D x D update belongs to rank t % world.  Per iteration, on every rank:
//   1. the covariance GEMM (h3_gemm.cuh, push mode) computes this rank's partial of EVERY lower tile (K = its batch shard)
//      and its epilogue stores each tile straight into the owner's staging slot [rank][t / world] through peer memory,
//      then bumps the owner's per-tile arrival counter (release, system scope): the reduce-scatter rides on the GEMM
//      epilogue, tile by tile, while later tiles are still in the tensor core;
//   2. comm_reduce_kernel: one CTA per owned tile waits for its `world` partials, adds them in rank order to the
//      current Sigma tile (deterministic: every rank ends up with bit-identical state), and stores the new tile and its
//      mirror into the NEXT Sigma buffer of every rank (the all-gather), then bumps every rank's "final tiles" counter;
//      one more CTA pushes this rank's mean increment to every rank;
//   3. comm_finalize_kernel waits until all tiles and all mean increments have arrived and forms the new mean.
// Counters only ever grow (targets are multiples of the step index), so nothing is reset between iterations.  A staging
// slot is rewritten by a peer only after that peer has received every final tile of the previous step, which the owner
// sends after reading the slot; the Sigma buffer a peer writes is never the one this rank is still reading.
#include "comm.cuh"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace gsmvi {

namespace {

constexpr int TILE = 128;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Watchdog of the peer waits: a rank may legitimately be late by a long time (a slow host-side score callable, a first-call
// JIT, a debugger), and a trap takes the CUDA context of every waiting rank down with it, so the limit is generous
// (GSMVI_COMM_TIMEOUT_S seconds, default 300; 0 = wait forever) and exists only to turn a dead peer into an error.
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned target, const char* what, long long timeout_cycles) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_sys(p) - target) < 0) {
    __nanosleep(100);
    if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
      printf("gsmvi: comm watchdog (%s, block %d, have %u want %u)\n", what, blockIdx.x, ld_acquire_sys(p), target);
      __trap();
    }
  }
}

struct ReduceArgs {
  float* const* base;  // device array [world] of comm-buffer bases
  gsmvi_comm_layout lay;
  int rank, world, D, cur;  // cur: index of the current Sigma buffer; the new one is 1 - cur
  unsigned step;
  const float* usum;  // this rank's sum_b u_b
  float inv_btotal;
  long long timeout_cycles;
};

constexpr int RQ = 4;              // CTAs per owned tile
constexpr int RROWS = TILE / RQ;   // rows of the tile per CTA

// blockIdx.x = li * RQ + quarter for the owned tiles, then one more CTA for the mean increment.  Only the lower triangle
// travels: each CTA stores its rows of the new tile (columns <= row on a diagonal tile) into every rank's next Sigma
// buffer; the upper triangle is mirrored locally afterwards (comm_mirror_kernel), which halves the NVLink traffic of
// the all-gather and keeps the result exactly symmetric.
__global__ void __launch_bounds__(256) comm_reduce_kernel(const ReduceArgs a) {
  const int tid = threadIdx.x;
  float* mine = a.base[a.rank];
  const int ntiles = a.lay.tiles_m * (a.lay.tiles_m + 1) / 2;
  const int li = blockIdx.x / RQ, quarter = blockIdx.x % RQ;
  const int t = li * a.world + a.rank;
  if (t >= ntiles) {
    if (blockIdx.x != gridDim.x - 1) return;
    // last CTA: this rank's mean increment to every rank's dmu[rank][:]
    for (int p = 0; p < a.world; ++p) {
      float* dst = a.base[p] + a.lay.dmu_off + static_cast<long long>(a.rank) * a.lay.lds;
      for (int j = tid; j < a.D; j += 256) dst[j] = a.usum[j] * a.inv_btotal;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        red_release_sys(reinterpret_cast<unsigned*>(a.base[p]) + a.lay.cnt_off + a.lay.tpo + 1, 1u);
    }
    return;
  }
  int tm = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  while (tm * (tm + 1) / 2 > t) --tm;
  const int tn = t - tm * (tm + 1) / 2;
  const int m0 = tm * TILE, n0 = tn * TILE;
  const bool diag = (tm == tn);
  if (tid == 0)
    spin_until(reinterpret_cast<const unsigned*>(mine) + a.lay.cnt_off + li, static_cast<unsigned>(a.world) * (a.step + 1), "partial tiles", a.timeout_cycles);
  __syncthreads();
  const float* stage = mine + a.lay.stage_off;
  const float* s0 = mine + a.lay.s_off[a.cur];
  const long long sn = a.lay.s_off[1 - a.cur];
  // RROWS x 128 elements = RROWS * 32 float4 groups, 256 threads: sum of the partials in rank order + current Sigma
  __shared__ __align__(16) float srow[RROWS][TILE + 4];
  const bool interior = (m0 + TILE <= a.D) && (n0 + TILE <= a.D);
  for (int q = tid; q < RROWS * TILE / 4; q += 256) {
    const int il = q >> 5, i = quarter * RROWS + il, j4 = (q & 31) * 4;
    const int m = m0 + i;
    if (m >= a.D || (diag && !interior && j4 > i)) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(stage + (static_cast<long long>(r) * a.lay.tpo + li) * (TILE * TILE) + i * TILE + j4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const long long off = static_cast<long long>(m) * a.lay.lds + n0 + j4;
    if (interior) {
      const float4 c = *reinterpret_cast<const float4*>(s0 + off);
      acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
      *reinterpret_cast<float4*>(&srow[il][j4]) = acc;  // staged: leaves as whole 512-byte rows below
    } else {
      const float v[4] = {acc.x, acc.y, acc.z, acc.w};
      for (int u = 0; u < 4; ++u)
        if (n0 + j4 + u < a.D) {
          const float o = v[u] + s0[off + u];
          for (int p = 0; p < a.world; ++p) a.base[p][sn + off + u] = o;
        }
    }
  }
  if (interior) {
    // one bulk copy (shared -> peer global) per (row, rank): 512 contiguous bytes each, full-size NVLink packets.  On a
    // diagonal tile the entries right of the diagonal are stale; the local mirror pass overwrites them before any use.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int il = tid & 31;
    const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(&srow[il][0]));
    const long long off = static_cast<long long>(m0 + quarter * RROWS + il) * a.lay.lds + n0;
    for (int p = tid >> 5; p < a.world; p += 8)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(a.base[p] + sn + off), "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
    for (int p = 0; p < a.world; ++p) red_release_sys(reinterpret_cast<unsigned*>(a.base[p]) + a.lay.cnt_off + a.lay.tpo, 1u);
  }
}

// Upper triangle of the new Sigma from its lower triangle (local, after every tile has arrived): one CTA per lower tile,
// transposed through shared memory; a diagonal tile mirrors inside itself.
__global__ void __launch_bounds__(256) comm_mirror_kernel(float* __restrict__ S, long long lds, int D, int tiles_m) {
  extern __shared__ float tile_raw[];
  float (*tile)[TILE + 1] = reinterpret_cast<float (*)[TILE + 1]>(tile_raw);
  const int t = blockIdx.x;
  int tm = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  while (tm * (tm + 1) / 2 > t) --tm;
  const int tn = t - tm * (tm + 1) / 2;
  const int m0 = tm * TILE, n0 = tn * TILE;
  for (int q = threadIdx.x; q < TILE * TILE / 4; q += 256) {
    const int i = q >> 5, j4 = (q & 31) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + i < D) {
      const float* row = S + static_cast<long long>(m0 + i) * lds + n0 + j4;
      for (int u = 0; u < 4; ++u)
        if (n0 + j4 + u < D) v[u] = row[u];
    }
    tile[i][j4] = v[0]; tile[i][j4 + 1] = v[1]; tile[i][j4 + 2] = v[2]; tile[i][j4 + 3] = v[3];
  }
  __syncthreads();
  const bool diag = (tm == tn);
  for (int q = threadIdx.x; q < TILE * TILE / 4; q += 256) {
    const int j = q >> 5, i4 = (q & 31) * 4;  // output row n0 + j, columns m0 + i4 ..
    if (n0 + j >= D) continue;
    float* drow = S + static_cast<long long>(n0 + j) * lds + m0 + i4;
    for (int u = 0; u < 4; ++u) {
      const int i = i4 + u;
      if (m0 + i < D && (!diag || i > j)) drow[u] = tile[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) comm_finalize_kernel(float* const* base, gsmvi_comm_layout lay, int rank, int world, int D,
                                                            unsigned step, const float* __restrict__ mu, float* __restrict__ mu_out,
                                                            long long timeout_cycles) {
  const float* mine = base[rank];
  const unsigned* cnt = reinterpret_cast<const unsigned*>(mine) + lay.cnt_off;
  const int ntiles = lay.tiles_m * (lay.tiles_m + 1) / 2;
  if (threadIdx.x == 0) {
    spin_until(cnt + lay.tpo, static_cast<unsigned>(ntiles) * RQ * (step + 1), "final tiles", timeout_cycles);
    spin_until(cnt + lay.tpo + 1, static_cast<unsigned>(world) * (step + 1), "mean increments", timeout_cycles);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += 256) {
    float acc = mu[j];
    for (int r = 0; r < world; ++r) acc += __ldcg(mine + lay.dmu_off + static_cast<long long>(r) * lay.lds + j);
    mu_out[j] = acc;
  }
}

}  // namespace

static long long comm_timeout_cycles() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("GSMVI_COMM_TIMEOUT_S");
    const double sec = e ? atof(e) : 300.0;
    v = sec <= 0.0 ? 0 : static_cast<long long>(sec * 2.0e9);  // clock64 ticks at <= 1.965 GHz: at least `sec` seconds
  }
  return v;
}

long long comm_layout(int D, int world, gsmvi_comm_layout* lay) {
  if (D <= 0 || world <= 0 || !lay) return -1;
  const long long lds = (D + 31) / 32 * 32;
  lay->tiles_m = (D + TILE - 1) / TILE;
  const int ntiles = lay->tiles_m * (lay->tiles_m + 1) / 2;
  lay->tpo = (ntiles + world - 1) / world;
  lay->lds = lds;
  long long off = 0;
  lay->stage_off = off;
  off += static_cast<long long>(world) * lay->tpo * TILE * TILE;
  lay->s_off[0] = off;
  off += static_cast<long long>(D) * lds;
  lay->s_off[1] = off;
  off += static_cast<long long>(D) * lds;
  lay->dmu_off = off;
  off += static_cast<long long>(world) * lds;
  lay->cnt_off = off;
  off += (lay->tpo + 2 + 31) / 32 * 32;
  return off * 4;
}

int comm_alloc(long long bytes, void** ptr, unsigned char* handle64) {
  if (bytes <= 0 || !ptr || !handle64) return GSMVI_EINVAL;
  cudaError_t e = cudaMalloc(ptr, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(*ptr, 0, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) return static_cast<int>(e);
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return GSMVI_OK;
}

int comm_open(const unsigned char* handle64, void** ptr) {
  if (!handle64 || !ptr) return GSMVI_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int comm_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int comm_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

int comm_reduce_broadcast(cudaStream_t stream, float* const* base, float* own_base, const gsmvi_comm_layout& lay, int rank,
                          int world, int D, int cur, unsigned step, const float* usum, float inv_btotal, const float* mu,
                          float* mu_out) {
  if (!own_base) return GSMVI_EINVAL;
  ReduceArgs a;
  a.base = base; a.lay = lay; a.rank = rank; a.world = world; a.D = D; a.cur = cur; a.step = step;
  a.usum = usum; a.inv_btotal = inv_btotal;
  a.timeout_cycles = comm_timeout_cycles();
  const int ntiles = lay.tiles_m * (lay.tiles_m + 1) / 2;
  const int mine = (ntiles - rank + world - 1) / world;  // tiles t = li * world + rank < ntiles
  constexpr int SMEM = TILE * (TILE + 1) * static_cast<int>(sizeof(float));
  {  // per call: the attribute is per device, a process may drive more than one, and setting it is cheap
    cudaError_t e = cudaFuncSetAttribute(comm_mirror_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  comm_reduce_kernel<<<mine * RQ + 1, 256, 0, stream>>>(a);
  comm_finalize_kernel<<<1, 256, 0, stream>>>(base, lay, rank, world, D, step, mu, mu_out, a.timeout_cycles);
  comm_mirror_kernel<<<ntiles, 256, SMEM, stream>>>(own_base + lay.s_off[1 - cur], lay.lds, D, lay.tiles_m);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GSMVI_OK : static_cast<int>(e);
}

}  // namespace gsmvi
