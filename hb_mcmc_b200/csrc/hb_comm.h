// hb_comm.h -- internal face of the communicator (hb_comm.cu) for the sampler (hb_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct hb_comm;

namespace hb {
int comm_rank(const hb_comm* c);
int comm_world(const hb_comm* c);
// in-place all-gather on `stream`: every rank's `count` doubles sit at d_buf + rank * count (capturable)
int comm_allgather_inplace(hb_comm* c, double* d_buf, size_t count, cudaStream_t stream);
}  // namespace hb
