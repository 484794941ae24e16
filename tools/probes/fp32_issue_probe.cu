// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/fp32_issue_probe tools/probes/fp32_issue_probe.cu
// Single-warp issue-rate / latency probes for the fp32 pieces on the Cholesky's critical path (CTA 0 of potrf_h3):
// independent FFMA vs packed FFMA2 (fma.rn.f32x2), broadcast LDS.128, mma.sync tf32 m16n8k8, rsqrt.approx, shfl.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void k(float* out, long long* cyc, float seed) {
  __shared__ __align__(16) float sh[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = seed * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed + i + lane;
  float x = seed * 1.0001f, y = seed * 0.999f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 32; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(x), "f"(y));
  }
  long long t1 = clock64();
  u64 a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a2[i] = (static_cast<u64>(__float_as_uint(acc[2 * i])) << 32) | __float_as_uint(acc[2 * i + 1]);
  u64 x2 = (static_cast<u64>(__float_as_uint(x)) << 32) | __float_as_uint(x);
  u64 y2 = (static_cast<u64>(__float_as_uint(y)) << 32) | __float_as_uint(y);
  long long t2 = clock64();
#pragma unroll 1
  for (int it = 0; it < 32; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a2[i] = ffma2(x2, y2, a2[i]);
  }
  long long t3 = clock64();
  // broadcast LDS.128: 64 loads per iteration
  float4 s4 = make_float4(0, 0, 0, 0);
  long long t4 = clock64();
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    float4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<const float4*>(sh + 132 * (i + it) + 4 * (it & 7));
#pragma unroll
    for (int i = 0; i < 16; ++i) { s4.x += v[i].x; s4.y += v[i].y; s4.z += v[i].z; s4.w += v[i].w; }
  }
  long long t5 = clock64();
  // mma.sync tf32: dependent chain of 64, then 4 independent accumulators x 16
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
  unsigned fa[4] = {__float_as_uint(sh[lane]), __float_as_uint(sh[lane + 32]), __float_as_uint(sh[lane + 64]), __float_as_uint(sh[lane + 96])};
  unsigned fb[2] = {__float_as_uint(sh[lane + 128]), __float_as_uint(sh[lane + 160])};
  long long t6 = clock64();
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_tf32(c0, fa, fb);
  }
  long long t7 = clock64();
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
#pragma unroll
    for (int i = 0; i < 2; ++i) { mma_tf32(c0, fa, fb); mma_tf32(c1, fa, fb); mma_tf32(c2, fa, fb); mma_tf32(c3, fa, fb); }
  }
  long long t8 = clock64();
  float r = fabsf(acc[0]) + 1.0f;
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
    asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(r));
    r += 2.0f;
  }
  long long t9 = clock64();
  float sv = r;
#pragma unroll 1
  for (int it = 0; it < 64; ++it) sv = __shfl_sync(0xffffffffu, sv, (lane + 1) & 31) + 1.0f;
  long long t10 = clock64();
  unsigned cv = __float_as_uint(sv);
#pragma unroll 1
  for (int it = 0; it < 64; ++it) { asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(cv) : "f"(__uint_as_float(cv) + 1.0f)); }
  long long t11 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; cyc[3] = t7 - t6; cyc[4] = t8 - t7; cyc[5] = t9 - t8; cyc[6] = t10 - t9; cyc[7] = t11 - t10;
  }
  float o = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) o += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) o += __uint_as_float(static_cast<unsigned>(a2[i])) + __uint_as_float(static_cast<unsigned>(a2[i] >> 32));
  out[threadIdx.x] = o + s4.x + s4.y + s4.z + s4.w + c0[0] + c1[1] + c2[2] + c3[3] + c0[3] + r + sv + __uint_as_float(cv);
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1024 * 4); cudaMalloc(&c, 64);
  for (int nt : {32, 128, 256}) {
    k<<<1, nt>>>(o, c, 1.5f); k<<<1, nt>>>(o, c, 1.5f); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    printf("threads %3d (warp 0's clock): 512 indep FFMA %lld cyc (%.2f/instr) | 256 FFMA2 %lld (%.2f/instr) | 128 bcast LDS.128 + 512 FADD %lld | "
           "64 dependent mma.sync tf32 m16n8k8 %lld (%.1f each) | 64 mma over 4 accumulators %lld (%.1f each) | rsqrt.approx+FADD chain %.1f | shfl+FADD chain %.1f | cvt.rna.tf32+FADD chain %.1f\n",
           nt, h[0], h[0] / 512.0, h[1], h[1] / 256.0, h[2], h[3], h[3] / 64.0, h[4], h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 64.0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
