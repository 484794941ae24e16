// BaM statistics + quadratic-matrix-equation solve (bam_solve.cu).
#pragma once
#include "tc_gemm.cuh"

namespace gsmvi {

size_t bam_stats_workspace_bytes(int B, int D);
size_t bam_solve_workspace_bytes(int B, int D, int lowrank);

// stage 0: unnormalised column sums of this shard's X, G into the workspace (xbar/gbar slots);
// stage 1: normalise by Btot, centre, and form C = Xc^T Xc / Btot, all in fp64; Gamma is never formed.
int bam_stats(cudaStream_t st, const float* X, long long ldx, const float* G, long long ldg, int B, int D, int Btot,
              double* ws, int stage);

// In-place fp64 Cholesky of the lower triangle (upper zeroed); *flag |= 1 if not PD.
int potrf64(cudaStream_t st, double* A, long long lda, int n, int* flag);

typedef gsmvi_bam_shard BamShard;

// shard != nullptr (world > 1, phase 2): tensor-parallel solve over peer-mapped workspaces (see gsmvi_bam_solve_sharded)
int bam_solve_full(cudaStream_t st, const double* stats_ws, int B, int D, int Btot, const float* mu0, const float* S0,
                   long long lds0, double reg, double jitter, float* mu_out, float* S_out, long long ldso, double* ws, int max_ns,
                   int* ns_iters_host, int* flag, int world, int phase, const BamShard* shard = nullptr);

int bam_solve_lowrank(cudaStream_t st, const double* stats_ws, int B, int D, int Btot, const float* mu0, const float* S0,
                      long long lds0, double reg, double jitter, float* mu_out, float* S_out, long long ldso, double* ws,
                      int max_ns, int* ns_iters_host, int* flag);

}  // namespace gsmvi
