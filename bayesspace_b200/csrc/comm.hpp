// comm.hpp -- the two exchanges of a spot-sharded chain, on NCCL (NVLink 5 / NVSwitch).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded in the process, or
// the system one), so single-GPU use of libbsb200.so has no NCCL dependency at all.  Both exchanges are
// enqueued on the chain's stream and are captured into its CUDA graph with the kernels around them.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace bsb {

struct Comm {
  void *nccl = nullptr;   // ncclComm_t
  int world = 1, rank = 0;
  ~Comm();
  int init(int world, int rank, const void *unique_id128);
  // in-place allgather: every rank contributes `count` doubles at buf + rank*count
  int allgather_f64(double *buf, size_t count, cudaStream_t s);
  // grouped send/recv of bytes with every peer that has a non-empty list
  int exchange_u8(const uint8_t *sendbuf, const int32_t *send_off, uint8_t *recvbuf, const int32_t *recv_off,
                  cudaStream_t s);   // offsets: world + 1 entries each (host)
  int allreduce_max_f64(double *buf, size_t count, cudaStream_t s);
};

int comm_unique_id(void *out128);

}   // namespace bsb
