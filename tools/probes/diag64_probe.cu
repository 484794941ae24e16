#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
constexpr int NB64 = 64; constexpr int L64S = NB64 + 1;
__global__ void __launch_bounds__(256, 1) potrf64_diag_kernel(double* __restrict__ a, long long lda, int n,
   double* __restrict__ linv, int* __restrict__ flag, long long* stamps) {
  long long t0 = clock64();
  extern __shared__ __align__(16) double sm64[];
  double* s = sm64;
  double* x = sm64 + NB64 * L64S;
  __shared__ double colj[2][NB64];
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  if (tid == 0) bad = 0;
  double r[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int i = 4 * ty + ii, k = 4 * tx + kk;
      r[ii][kk] = (i < n && k <= i) ? a[static_cast<long long>(i) * lda + k] : ((i == k) ? 1.0 : 0.0);
    }
  for (int idx = tid; idx < NB64 * NB64; idx += 256) x[(idx / NB64) * L64S + idx % NB64] = 0.0;
  __syncthreads();
  long long t1 = clock64();
#pragma unroll 1
  for (int jb = 0; jb < NB64 / 4; ++jb) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * jb + jj;
      double* cj = colj[j & 1];
      if (tx == jb && ty >= jb) {  // the threads holding column j (rows >= 4 jb) publish it, unscaled
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) cj[4 * ty + ii] = r[ii][jj];
      }
      __syncthreads();
      const double p = cj[j];
      if (!(p > 0.0) || isinf(p)) {
        if (tid == 0) bad = 1;
      }
      // 1 / sqrt(p): fp32 seed (23 bits) + two Newton steps in fp64, six dependent DFMAs - the library rsqrt / division
      // are ~30-operation sequences, and this value heads the dependent chain of every one of the 64 steps
      double rs;
      if (p > 1e-30 && p < 1e30) {
        rs = static_cast<double>(rsqrtf(static_cast<float>(p)));
        rs = fma(0.5 * rs, fma(-p * rs, rs, 1.0), rs);
        rs = fma(0.5 * rs, fma(-p * rs, rs, 1.0), rs);
      } else {
        rs = rsqrt(p);  // outside the fp32 range (or not a valid pivot): the slow exact path
      }
      if (ty < jb || tx > ty) continue;  // nothing left to update in this block (uniform per thread across steps: no hazard)
      double li[4], lk[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) li[ii] = cj[4 * ty + ii] * rs;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) lk[kk] = cj[4 * tx + kk] * rs;
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int i = 4 * ty + ii, k = 4 * tx + kk;
          if (k == j) {
            if (i >= j) r[ii][kk] = (i == j) ? p * rs : li[ii];  // column j itself: l_jj = sqrt(p), l_ij = a_ij / l_jj
          } else if (k > j && i >= k) {
            r[ii][kk] -= li[ii] * lk[kk];
          }
        }
    }
  }
  long long t2 = clock64();
  // L -> shared memory
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int i = 4 * ty + ii, k = 4 * tx + kk;
      s[i * L64S + k] = (k <= i) ? r[ii][kk] : 0.0;
    }
  __syncthreads();
  // inverse X = L^-1 by forward substitution, row after row: column c of X belongs to the four adjacent lanes 4c .. 4c+3,
  // which split the inner sum over k (k = c + q, c + q + 4, ...) and combine it with two shuffles, so the dependent chain
  // per row is ~16 FMAs instead of the ~64 a thread-per-column substitution walks
  {
    // reciprocals of the diagonal once, in parallel (the pivots' rsqrt values squared would do too, but this is off the chain)
    double* dr = colj[0];
    if (tid < NB64) dr[tid] = 1.0 / s[tid * L64S + tid];
    __syncthreads();
    const int c = tid >> 2, q = tid & 3;
    for (int i = 0; i < NB64; ++i) {
      double acc0 = 0.0, acc1 = 0.0;  // two independent accumulation chains
      if (i > c) {
        int k = c + q;
        for (; k + 4 < i; k += 8) {
          acc0 += s[i * L64S + k] * x[k * L64S + c];
          acc1 += s[i * L64S + k + 4] * x[(k + 4) * L64S + c];
        }
        if (k < i) acc0 += s[i * L64S + k] * x[k * L64S + c];
      }
      double acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (q == 0 && i >= c) x[i * L64S + c] = (i == c) ? dr[c] : -acc * dr[i];
      __syncwarp();  // the four lanes of a column (always in one warp) see the new entry before the next row
    }
  }
  __syncthreads();
  long long t3 = clock64();
  for (int idx = tid; idx < NB64 * NB64; idx += 256) {
    const int i = idx / NB64, j = idx % NB64;
    if (i < n && j < n) a[static_cast<long long>(i) * lda + j] = s[i * L64S + j];
    linv[i * NB64 + j] = (i < n && j < n) ? x[i * L64S + j] : 0.0;
  }
  if (tid == 0 && bad) atomicOr(flag, 1);
  if (tid == 0) { stamps[0]=t1-t0; stamps[1]=t2-t1; stamps[2]=t3-t2; stamps[3]=clock64()-t3; }
}


int main(){
  const int n=64; double *a,*inv; int* flag; long long* st;
  cudaMalloc(&a,n*n*8); cudaMalloc(&inv,n*n*8); cudaMalloc(&flag,4); cudaMalloc(&st,64);
  static double h[n*n], L[n*n], X[n*n];
  for(int i=0;i<n;i++)for(int j=0;j<n;j++)h[i*n+j]=(i==j)?n+1.0:0.5+0.001*((i*7+j*3)%11);
  for(int i=0;i<n;i++)for(int j=0;j<i;j++)h[j*n+i]=h[i*n+j];
  const int smem=2*NB64*L64S*8;
  cudaFuncSetAttribute(potrf64_diag_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,smem);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int rep=0;rep<3;rep++){
    cudaMemcpy(a,h,n*n*8,cudaMemcpyHostToDevice);
    cudaEventRecord(e0); potrf64_diag_kernel<<<1,256,smem>>>(a,n,n,inv,flag,st); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms,e0,e1); long long s4[4]; cudaMemcpy(s4,st,32,cudaMemcpyDeviceToHost);
    printf("kernel %.1f us | cycles: load %lld factor %lld inverse %lld store %lld  err=%s\n",ms*1e3,s4[0],s4[1],s4[2],s4[3],cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpy(L,a,n*n*8,cudaMemcpyDeviceToHost); cudaMemcpy(X,inv,n*n*8,cudaMemcpyDeviceToHost);
  double e1m=0,e2m=0;
  for(int i=0;i<n;i++)for(int j=0;j<=i;j++){double s=0;for(int k=0;k<=j;k++)s+=L[i*n+k]*L[j*n+k]; e1m=fmax(e1m,fabs(s-h[i*n+j]));}
  for(int i=0;i<n;i++)for(int j=0;j<n;j++){double s=0;for(int k=0;k<n;k++)s+=X[i*n+k]*L[k*n+j]; e2m=fmax(e2m,fabs(s-(i==j)));}
  printf("max |L L^T - A| = %.3e   max |X L - I| = %.3e\n",e1m,e2m);
  return 0;
}
