// Latency microbenchmarks of the operations on the dependent chain of the fp64 diagonal-block Cholesky (one warp / one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  long long t1 = clock64();
  double r = x;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) r = rsqrt(r + 2.0);
  long long t2 = clock64();
  double d = r + 3.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) d = 1.0 / (d + 2.0);
  long long t3 = clock64();
  double s = d + 2.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    double rs = (double)rsqrtf((float)s);
    rs = fma(0.5 * rs, fma(-s * rs, rs, 1.0), rs);
    rs = fma(0.5 * rs, fma(-s * rs, rs, 1.0), rs);
    s = rs + 2.0;
  }
  long long t4 = clock64();
  __shared__ double sh[256];
  double w = s;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    sh[threadIdx.x] = w;
    __syncthreads();
    w = sh[(threadIdx.x + 1) & 255] + 1.0;
  }
  long long t5 = clock64();
  float f = (float)w;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
  long long t6 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0) / 256; cyc[1] = (t2 - t1) / 64; cyc[2] = (t3 - t2) / 64; cyc[3] = (t4 - t3) / 64; cyc[4] = (t5 - t4) / 64; cyc[5] = (t6 - t5) / 256;
  }
  out[threadIdx.x] = x + r + d + s + w + f;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 64);
  for (int nt : {32, 256}) {
    k<<<1, nt>>>(o, c, 1.5); cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
    printf("threads %3d: DFMA dep %lld | rsqrt(double) %lld | 1.0/x %lld | rsqrtf seed + 2 Newton %lld | st.shared+__syncthreads+ld %lld | FFMA dep %lld cycles\n",
           nt, h[0], h[1], h[2], h[3], h[4], h[5]);
  }
  return 0;
}
