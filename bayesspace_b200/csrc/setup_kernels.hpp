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

// Internal spot order of an unsharded chain, computed on the device from the caller's CSR graph:
// scan() uploads the graph and returns the sizes of the (shard, colour) buckets, the largest degree and whether the
// colouring is proper; the host lays the buckets and segments out (SpotOrder::layout) and place() writes the
// original ids, the initial labels and the neighbour lists in internal order.
struct DeviceOrder {
  int64_t n = 0;
  int ncol = 0, G = 1;
  DevBuf<int64_t> ptr;
  DevBuf<int32_t> idx, init, colour, val, val2, inv;
  DevBuf<uint32_t> key, key2;
  DevBuf<unsigned long long> dcnt;
  DevBuf<int32_t> flags;
  // upload(): device copies of the caller's arrays (h_colour may be null); scan(): h_colour_new replaces the colours
  int prealloc(int64_t n);
  int upload(const int64_t *h_ptr, const int32_t *h_idx, const int32_t *h_init, const int32_t *h_colour, int64_t n, int device);
  int scan(const int32_t *h_colour_new, int ncol, int G, cudaStream_t s, std::vector<int64_t> &cnt, int *maxdeg, bool *proper);
  int place(const std::vector<int64_t> &cnt, const std::vector<int64_t> &start, int64_t ld, int maxdeg, int32_t *d_orig, uint8_t *d_z,
            int32_t *d_ell, cudaStream_t s);
};

}   // namespace bsb
