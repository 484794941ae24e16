// Peer-memory exchange of the GSM batch statistics (comm.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/gsmvi_b200.h"

namespace gsmvi {

long long comm_layout(int D, int world, gsmvi_comm_layout* lay);
int comm_alloc(long long bytes, void** ptr, unsigned char* handle64);
int comm_open(const unsigned char* handle64, void** ptr);
int comm_close(void* peer_ptr);
int comm_free(void* ptr);
// reduce the staged partial tiles this rank owns into the next Sigma buffer of every rank, exchange the mean increments
// and form mu_out = mu + sum_r dmu_r (see comm.cu)
// own_base: this rank's buffer (the value of base[rank], as a host-visible pointer)
int comm_reduce_broadcast(cudaStream_t stream, float* const* base, float* own_base, const gsmvi_comm_layout& lay, int rank,
                          int world, int D, int cur, unsigned step, const float* usum, float inv_btotal, const float* mu,
                          float* mu_out);

}  // namespace gsmvi
