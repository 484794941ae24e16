// FP64 GEMM on the int8 tensor cores (Ozaki splitting) for the BaM solve:  C = alpha * op(A) op(B)^T + beta * Cin + diag.
//
// tcgen05 has no f64 kind and the B200's FP64 pipe tops out near 40 TFLOP/s (cuBLAS DGEMM measured 36), while
// kind::i8 runs at ~4.5 POP/s with an EXACT int32 accumulator.  Every row of op(A) (and of op(B)) is scaled by a power
// of two so that |a| < 1 and cut into s signed 7-bit digits,  a = 2^e sum_t q_t 2^(-7t),  q_t in [-127, 127]  (exact: the
// digits are peeled off an fp64 value by multiply / truncate / subtract).  Then
//     A B^T = 2^(ea_i + eb_j) sum_g 2^(-7g) sum_{t+u=g} Q_t R_u^T ,      g = 2 .. s+1  (pairs with t + u > s + 1 dropped:
// they sit below the last kept digit), and each inner sum is an int8 GEMM whose int32 result is exact for K <= 8192.
// One launch per digit group g accumulates its pairs over a concatenated K range in a single TMEM accumulator and
// adds 2^(-7g) * (int32) into an fp64 scratch (least significant group first); the last launch applies the row / column
// scales, alpha, beta and the diagonal.  s = 8 digits carry 56 bits (36 int8 GEMMs); the relative accuracy is that of
// a fixed-point dot product per (row, column) pair, 2^(-7s) K against max|a_i| max|b_j|.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dgemm.cuh"

namespace gsmvi {

// bytes of scratch for an M x N x K product with `slices` digits (planes of both operands, fp64 accumulator, scales)
size_t oz_workspace_bytes(int M, int N, int K, int slices);

// Same contract as launch_dgemm (dgemm.cuh); krange is ignored (zeros stay zeros).  ws: oz_workspace_bytes(...) bytes,
// 1 KiB aligned.  slices in [2, 8].
int launch_dgemm_oz(cudaStream_t stream, int M, int N, int K, const double* A, long long lda, bool a_mn, const double* B,
                    long long ldb, bool b_mn, double* C, long long ldc, const DgemmOpts& o, void* ws, int slices);

}  // namespace gsmvi
