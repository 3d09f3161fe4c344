// chain.hpp -- host-side building blocks shared by the cluster and deconvolution samplers.
#pragma once
#include <memory>
#include <vector>

#include "common.hpp"
#include "kernels.hpp"
#include "util_kernels.hpp"

namespace bsb {

// Owning device buffer.
// Device memory comes from the device's default stream-ordered pool with its release threshold raised (chain.cu:
// select_device), so that a destroyed chain's buffers stay mapped for the next one: cudaFree / cudaMalloc of the 3 GB a
// 10M-spot chain holds cost between 5 and 600 ms of wall clock per bsb200_cluster() call.  bsb200_release_memory() hands
// the cached memory back to the driver.  pool_alloc returns memory valid on every stream; pool_free waits for the device.
cudaError_t pool_alloc(void **p, size_t bytes);
void pool_free(void *p);

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t count = 0;
  bool owned = true;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { if (p && owned) pool_free(p); }
  // a region of somebody else's allocation (the exported arena of the peer transport)
  void adopt(T *ptr, size_t n) {
    if (p && owned) pool_free(p);
    p = ptr; count = n; owned = false;
  }
  void release() {
    if (p && owned) pool_free(p);
    p = nullptr; count = 0;
  }
  int alloc(size_t n) {
    if (p && owned) pool_free(p);
    p = nullptr; owned = true;
    count = n;
    if (n == 0) n = 1;
    cudaError_t e = pool_alloc((void **) &p, n * sizeof(T));
    if (e != cudaSuccess) return fail(BSB200_ERR_CUDA, "device allocation of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return BSB200_OK;
  }
};

// Internal spot order and the segment / slot / group tables derived from it (chain.cu).
struct SpotOrder {
  int ncol = 0, ngroups = 1, nslots = 0;
  int64_t npos = 0;            // internal positions, holes included (every (shard, colour) range starts at a multiple of `align`)
  std::vector<int32_t> perm;   // internal position -> original spot, -1 = hole
  std::vector<int32_t> inv;    // original spot -> internal position
  std::vector<int32_t> seg_begin, seg_count, seg_slot;   // launch lists, colour-major
  std::vector<int32_t> colour_seg_start;                 // ncol + 1 offsets into the launch lists
  std::vector<int32_t> group_start;                      // ngroups + 1 offsets into the slots
  // spots [lo, hi) = virtual shards [g0, g1) of ngroups_total; defaults: everything
  // align: first position of every (shard, colour) range (TMA boxes of k_sweep_fast start at 16-byte boundaries)
  void layout(const std::vector<int64_t> &cnt, int seg_spots, int align, std::vector<int64_t> &start);
  int build(int64_t n, const int32_t *colour, int ncolours, int ngroups_total, int seg_spots, int g0 = 0, int g1 = -1,
            int64_t lo = 0, int64_t hi = 0, int align = 1);
};

int select_device(int device);
constexpr int kOrderAlign = 32;   // alignment of the ranges in the cluster samplers' internal order
int choose_groups(int64_t n);
int choose_segment(int64_t n, int ncol = 0, int d = 15);   // ncol = 0: the plain n-based rule (probes, deconvolution)
void param_kernel_init();
size_t param_smem_min(int q, int d);   // least dynamic shared memory k_param needs (one matrix per batch)

}   // namespace bsb
