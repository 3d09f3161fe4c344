// setup_kernels.hpp -- device side of the chain set-up (setup_kernels.cu).
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "chain.hpp"
#include "common.hpp"
#include "kernels.hpp"

namespace bsb {

// Host-to-device copies of pageable memory through a pool of page-locked buffers filled by host threads.
// Blocks until every copy has arrived.  One transfer at a time per process.
struct StagedCopy {
  void *dst;
  const void *src;
  size_t bytes;
};
int staged_h2d(int device, const StagedCopy *copies, int ncopies);
void staged_release();   // frees the page-locked staging buffers (bsb200_release_memory)

// Internal spot order of a chain -- of the spots [lo, hi) = virtual shards [g0, g1) a rank owns -- computed on the
// device from the caller's CSR graph: upload() copies the graph; scan() returns the sizes of the own (shard, colour)
// buckets, the largest degree and whether the colouring is proper (judged on the whole graph, so that every rank of a
// sharded chain decides alike); the host lays buckets and segments out (SpotOrder::layout); positions() sorts the own
// spots into them (stable radix sort), appends the ghosts and hands the own positions to the host when asked (halo
// plan of a sharded chain); place() writes original ids, initial labels and neighbour lists in internal order.
struct DeviceOrder {
  int64_t n = 0, lo = 0, hi = 0;
  int ncol = 0, G = 1, g0 = 0, g1 = 1;
  DevBuf<int64_t> ptr;
  DevBuf<int32_t> idx, init, colour, val, val2, inv, ghosts;
  DevBuf<uint32_t> key, key2;
  DevBuf<unsigned long long> dcnt;
  DevBuf<int32_t> flags;
  int prealloc(int64_t n, int64_t n_own);
  int upload(const int64_t *h_ptr, const int32_t *h_idx, const int32_t *h_init, const int32_t *h_colour, int64_t n, int device);
  int scan(const int32_t *h_colour_new, int ncol, int G, int g0, int g1, int64_t lo, int64_t hi, cudaStream_t s,
           std::vector<int64_t> &cnt, int *maxdeg, bool *proper);
  int positions(const std::vector<int64_t> &cnt, const std::vector<int64_t> &start, const std::vector<int32_t> &ghost_ids, int64_t n_pos,
                std::vector<int32_t> *own_inv, cudaStream_t s);
  int place(int64_t ld, int maxdeg, int32_t *d_orig, uint8_t *d_z, int32_t *d_ell, cudaStream_t s);
  void release();   // every device buffer (the upload thread must have delivered the graph)
};

}   // namespace bsb
