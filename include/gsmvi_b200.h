/* gsmvi_b200.h - C ABI of libgsmvi_b200.so (work in progress).
 * All pointers are DEVICE pointers unless a name ends in _host. All functions are stream-ordered,
 * never synchronise unless documented, and return 0 on success, <0 for an invalid argument
 * (GSMVI_E*), >0 for a cudaError_t. */
#ifndef GSMVI_B200_H
#define GSMVI_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define GSMVI_ABI_VERSION 1

int gsmvi_abi_version(void);

int gsmvi_gemm_tf32(const float* A, long long a_rows, long long a_cols, long long lda, int a_mn, const float* B,
                    long long b_rows, long long b_cols, long long ldb, int b_mn, float* C, long long ldc, int M, int N,
                    int K, float alpha, float beta, const float* Cin, long long ldcin, const float* bias_n, int npass,
                    int tri, int mirror, int krange, int neg_from, void* stream);

#ifdef __cplusplus
}
#endif
#endif
