"""Generate tools/probes/diag64_probe.cu: the fp64 diagonal-block kernel of csrc/bam_solve.cu with clock64 stamps around its
phases (load / factor / inverse / store), as a standalone program (nvcc -gencode arch=compute_100a,code=sm_100a -O3)."""
import os
HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(HERE, "..", "..", "gsm-vi_b200", "csrc", "bam_solve.cu")).read()
a = src.index("__global__ void __launch_bounds__(256, 1) potrf64_diag_kernel")
b = src.index("// zero the strict upper triangle (and optionally symmetrise from the lower one)")
k = src[a:b]
k = k.replace("potrf64_diag_kernel(double* __restrict__ a, long long lda, int n,\n                                                              double* __restrict__ linv, int* __restrict__ flag) {",
              "potrf64_diag_kernel(double* __restrict__ a, long long lda, int n,\n   double* __restrict__ linv, int* __restrict__ flag, long long* stamps) {\n  long long t0 = clock64();")
k = k.replace("  __syncthreads();\n#pragma unroll 1\n  for (int jb = 0;", "  __syncthreads();\n  long long t1 = clock64();\n#pragma unroll 1\n  for (int jb = 0;")
k = k.replace("  // L -> shared memory (zeros above the diagonal) for the inverse", "  long long t2 = clock64();\n  // L -> shared memory")
k = k.replace("  __syncthreads();\n  for (int idx = tid; idx < NB64 * NB64; idx += 256) {\n    const int i = idx / NB64, j = idx % NB64;\n    if (i < n && j < n) a[",
              "  __syncthreads();\n  long long t3 = clock64();\n  for (int idx = tid; idx < NB64 * NB64; idx += 256) {\n    const int i = idx / NB64, j = idx % NB64;\n    if (i < n && j < n) a[")
k = k.replace("  if (tid == 0 && bad) atomicOr(flag, 1);\n}", "  if (tid == 0 && bad) atomicOr(flag, 1);\n  if (tid == 0) { stamps[0]=t1-t0; stamps[1]=t2-t1; stamps[2]=t3-t2; stamps[3]=clock64()-t3; }\n}")
assert "stamps[0]" in k and "t3 = clock64" in k and "t2 = clock64" in k and "t1 = clock64" in k
prog = '''#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
constexpr int NB64 = 64; constexpr int L64S = NB64 + 1;
''' + k + '''
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
    printf("kernel %.1f us | cycles: load %lld factor %lld inverse %lld store %lld  err=%s\\n",ms*1e3,s4[0],s4[1],s4[2],s4[3],cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpy(L,a,n*n*8,cudaMemcpyDeviceToHost); cudaMemcpy(X,inv,n*n*8,cudaMemcpyDeviceToHost);
  double e1m=0,e2m=0;
  for(int i=0;i<n;i++)for(int j=0;j<=i;j++){double s=0;for(int k=0;k<=j;k++)s+=L[i*n+k]*L[j*n+k]; e1m=fmax(e1m,fabs(s-h[i*n+j]));}
  for(int i=0;i<n;i++)for(int j=0;j<n;j++){double s=0;for(int k=0;k<n;k++)s+=X[i*n+k]*L[k*n+j]; e2m=fmax(e2m,fabs(s-(i==j)));}
  printf("max |L L^T - A| = %.3e   max |X L - I| = %.3e\\n",e1m,e2m);
  return 0;
}
'''
open(os.path.join(HERE, "diag64_probe.cu"), "w").write(prog)
