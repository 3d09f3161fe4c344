// setup_kernels.cu -- set-up of a chain on the device: the staged host-to-device copies of the caller's (pageable)
// arrays and, for unsharded chains, the internal spot order and the neighbour lists in that order.
//
// Why: at 10M spots the host version of this work (chain.cu: SpotOrder::build, the ELL fill) and the driver's
// single-threaded staging of pageable memory (1.2 GB of PCs at ~6 GB/s) were a quarter of the wall clock of a
// 1000-sweep bsb200_cluster() call.  Here host threads copy chunks into a small pool of page-locked buffers that are
// sent with cudaMemcpyAsync as they fill, and the order / neighbour lists are computed from the uploaded CSR arrays by
// three kernels and one radix sort.  The result is identical to the host version (tests/test_gpu_zv_setup.py).
#include "setup_kernels.hpp"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return bsb::fail(BSB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

namespace bsb {

static inline void count_launch() { g_launches++; }

// ------------------------------------------------------------------------------------------------
// Staged copies
// ------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kStageChunk = 4u << 20;   // bytes per page-locked buffer
constexpr int kStageWorkers = 12, kStageDepth = 2;   // buffers exist for 12 workers; stage_workers() of them run

struct StagePool {   // one per process and device, allocated on first use, never freed (like the CUDA context)
  std::mutex mu;
  int device = -1;
  unsigned char *buf[kStageWorkers][kStageDepth] = {};
  cudaStream_t stream[kStageWorkers] = {};
  cudaEvent_t ev[kStageWorkers][kStageDepth] = {};
  bool ready = false;
  cudaError_t init(int dev) {
    if (ready && device == dev) return cudaSuccess;
    if (ready) release();
    for (int w = 0; w < kStageWorkers; w++) {
      cudaError_t e = cudaStreamCreateWithFlags(&stream[w], cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
      for (int k = 0; k < kStageDepth; k++) {
        if ((e = cudaMallocHost(&buf[w][k], kStageChunk)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev[w][k], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
    }
    device = dev;
    ready = true;
    return cudaSuccess;
  }
  void release() {
    for (int w = 0; w < kStageWorkers; w++) {
      for (int k = 0; k < kStageDepth; k++) {
        if (buf[w][k]) cudaFreeHost(buf[w][k]);
        if (ev[w][k]) cudaEventDestroy(ev[w][k]);
        buf[w][k] = nullptr; ev[w][k] = nullptr;
      }
      if (stream[w]) cudaStreamDestroy(stream[w]);
      stream[w] = nullptr;
    }
    ready = false;
  }
};
StagePool g_pool;
}   // namespace

int staged_h2d(int device, const StagedCopy *copies, int ncopies) {
  std::lock_guard<std::mutex> lock(g_pool.mu);   // one staged transfer at a time per process
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = g_pool.init(device);
  if (e != cudaSuccess) return fail(BSB200_ERR_CUDA, "staging buffers: %s", cudaGetErrorString(e));
  // chunk list: (copy, offset)
  struct Chunk { int c; size_t off, len; };
  std::vector<Chunk> chunks;
  for (int c = 0; c < ncopies; c++)
    for (size_t off = 0; off < copies[c].bytes; off += kStageChunk)
      chunks.push_back({c, off, std::min(kStageChunk, copies[c].bytes - off)});
  std::atomic<size_t> next{0};
  std::atomic<int> err{(int) cudaSuccess};
  auto worker = [&](int w) {
    cudaError_t ew = cudaSetDevice(device);
    int k = 0;
    bool used[kStageDepth] = {};
    while (ew == cudaSuccess) {
      const size_t i = next.fetch_add(1);
      if (i >= chunks.size()) break;
      const Chunk &ck = chunks[i];
      if (used[k]) ew = cudaEventSynchronize(g_pool.ev[w][k]);   // the buffer's previous copy has left the host
      if (ew != cudaSuccess) break;
      std::memcpy(g_pool.buf[w][k], (const unsigned char *) copies[ck.c].src + ck.off, ck.len);
      ew = cudaMemcpyAsync((unsigned char *) copies[ck.c].dst + ck.off, g_pool.buf[w][k], ck.len, cudaMemcpyHostToDevice,
                           g_pool.stream[w]);
      if (ew == cudaSuccess) ew = cudaEventRecord(g_pool.ev[w][k], g_pool.stream[w]);
      used[k] = true;
      k = (k + 1) % kStageDepth;
    }
    if (ew == cudaSuccess) ew = cudaStreamSynchronize(g_pool.stream[w]);
    if (ew != cudaSuccess) err.store((int) ew);
  };
  static const int want = []() {
    int v = 6;
    if (const char *e = std::getenv("BSB200_STAGE_WORKERS")) v = std::atoi(e);
    return v < 1 ? 1 : (v > kStageWorkers ? kStageWorkers : v);
  }();
  const int nw = (int) std::min<size_t>((size_t) want, std::max<size_t>(chunks.size(), 1));
  std::vector<std::thread> th;
  for (int w = 1; w < nw; w++) th.emplace_back(worker, w);
  worker(0);
  for (auto &t : th) t.join();
  if (err.load() != (int) cudaSuccess) return fail(BSB200_ERR_CUDA, "staged upload: %s", cudaGetErrorString((cudaError_t) err.load()));
  return BSB200_OK;
}

void staged_release() {
  std::lock_guard<std::mutex> lock(g_pool.mu);
  if (g_pool.ready) g_pool.release();
}

// ------------------------------------------------------------------------------------------------
// Spot order and neighbour lists on the device (own spots [lo, hi) of a rank, all of them for an unsharded chain)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kScanBlock = 256;
constexpr int kMaxSmemBuckets = 8192;   // 32 KB of block-local histogram

// Every spot: largest degree, colour in range and not shared with a neighbour (all ranks of a sharded chain look at the
// whole graph and so take the same decision).  Own spots [lo, hi): bucket key ((shard - g0) * ncol + colour) and identity
// value for the sort; per own bucket: size.
__global__ void __launch_bounds__(kScanBlock) k_graph_scan(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                           const int32_t *__restrict__ colour, int64_t n, int ncol, int G, int g0,
                                                           int nb, int64_t lo, int64_t hi, uint32_t *__restrict__ key,
                                                           int32_t *__restrict__ val, unsigned long long *__restrict__ cnt,
                                                           int32_t *__restrict__ flags) {
  extern __shared__ uint32_t hist[];
  const bool in_smem = nb <= kMaxSmemBuckets;   // more buckets than that (thousands of shards x colours): global atomics
  if (in_smem) {
    for (int b = threadIdx.x; b < nb; b += kScanBlock) hist[b] = 0;
    __syncthreads();
  }
  int maxdeg = 0, bad = 0;
  for (int64_t j = (int64_t) blockIdx.x * kScanBlock + threadIdx.x; j < n; j += (int64_t) gridDim.x * kScanBlock) {
    const int32_t c = colour[j];
    const int64_t p0 = ptr[j], p1 = ptr[j + 1];
    maxdeg = max(maxdeg, (int) (p1 - p0));
    if (c < 0 || c >= ncol) { bad |= 2; atomicMin(&flags[2], (int32_t) min(j, (int64_t) 0x7fffffff)); continue; }
    for (int64_t p = p0; p < p1; p++) bad |= colour[idx[p]] == c;
    if (j < lo || j >= hi) continue;
    const int g = (int) ((j * (int64_t) G) / n) - g0;
    const uint32_t b = (uint32_t) (g * ncol + c);
    key[j - lo] = b;
    val[j - lo] = (int32_t) j;
    if (in_smem) atomicAdd(&hist[b], 1u);
    else atomicAdd(&cnt[b], 1ull);
  }
  if (in_smem) {
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += kScanBlock)
      if (hist[b]) atomicAdd(&cnt[b], (unsigned long long) hist[b]);
  }
  maxdeg = __reduce_max_sync(0xffffffffu, maxdeg);
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&flags[0], maxdeg);
    if (bad) atomicOr(&flags[1], bad);
  }
}

// sorted rank t -> internal position: start of the bucket (aligned, holes between buckets) + rank inside it
__global__ void __launch_bounds__(kScanBlock) k_positions(const uint32_t *__restrict__ key, const int32_t *__restrict__ val,
                                                          int64_t n_own, const int64_t *__restrict__ start,
                                                          const int64_t *__restrict__ ustart, int32_t *__restrict__ inv) {
  for (int64_t t = (int64_t) blockIdx.x * kScanBlock + threadIdx.x; t < n_own; t += (int64_t) gridDim.x * kScanBlock) {
    const uint32_t b = key[t];
    inv[val[t]] = (int32_t) (start[b] + (t - ustart[b]));
  }
}
// ghosts (spots of other ranks read by own spots) follow the own positions, in ascending original id
__global__ void __launch_bounds__(kScanBlock) k_ghost_positions(const int32_t *__restrict__ ghosts, int64_t nghost, int64_t n_pos,
                                                                int32_t *__restrict__ inv) {
  for (int64_t g = (int64_t) blockIdx.x * kScanBlock + threadIdx.x; g < nghost; g += (int64_t) gridDim.x * kScanBlock)
    inv[ghosts[g]] = (int32_t) (n_pos + g);
}
// original id and initial label at every position (own spots, then ghosts)
__global__ void __launch_bounds__(kScanBlock) k_place(const int32_t *__restrict__ ids, int64_t count, const int32_t *__restrict__ inv,
                                                      const int32_t *__restrict__ init, int32_t *__restrict__ orig,
                                                      uint8_t *__restrict__ z) {
  for (int64_t t = (int64_t) blockIdx.x * kScanBlock + threadIdx.x; t < count; t += (int64_t) gridDim.x * kScanBlock) {
    const int32_t j = ids[t];
    const int64_t pos = inv[j];
    orig[pos] = j;
    z[pos] = (uint8_t) (init[j] - 1);
  }
}

__global__ void __launch_bounds__(kScanBlock) k_ell(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, int64_t lo,
                                                    int64_t hi, const int32_t *__restrict__ inv, int32_t *__restrict__ ell,
                                                    int64_t ld) {
  for (int64_t j = lo + (int64_t) blockIdx.x * kScanBlock + threadIdx.x; j < hi; j += (int64_t) gridDim.x * kScanBlock) {
    const int64_t pos = inv[j];
    int e = 0;
    for (int64_t p = ptr[j]; p < ptr[j + 1]; p++, e++) ell[(size_t) e * ld + pos] = inv[idx[p]];
  }
}

int grid_for(int64_t n) { return (int) std::min<int64_t>((n + kScanBlock - 1) / kScanBlock, 148 * 8); }

}   // namespace

int DeviceOrder::upload(const int64_t *h_ptr, const int32_t *h_idx, const int32_t *h_init, const int32_t *h_colour, int64_t n_,
                        int device) {
  n = n_;
  const int64_t nnz = h_ptr[n];
  int rc;
  if ((rc = ptr.alloc((size_t) n + 1)) || (rc = idx.alloc((size_t) std::max<int64_t>(nnz, 1))) || (rc = init.alloc((size_t) n)) ||
      (rc = colour.alloc((size_t) n)))
    return rc;
  const StagedCopy cp[4] = {{ptr.p, h_ptr, sizeof(int64_t) * ((size_t) n + 1)},
                            {idx.p, h_idx, sizeof(int32_t) * (size_t) nnz},
                            {init.p, h_init, sizeof(int32_t) * (size_t) n},
                            {colour.p, h_colour, sizeof(int32_t) * (size_t) n}};
  return staged_h2d(device, cp, h_colour ? 4 : 3);
}

int DeviceOrder::prealloc(int64_t n_, int64_t n_own_) {   // the sort's buffers: allocated by the caller's thread while upload() copies
  int rc;
  const size_t m = (size_t) std::max<int64_t>(n_own_, 1);
  if ((rc = key.alloc(m)) || (rc = val.alloc(m)) || (rc = key2.alloc(m)) || (rc = val2.alloc(m)) ||
      (rc = inv.alloc((size_t) n_)) || (rc = flags.alloc(3)))
    return rc;
  return BSB200_OK;
}

int DeviceOrder::scan(const int32_t *h_colour_new, int ncol_, int G_, int g0_, int g1_, int64_t lo_, int64_t hi_, cudaStream_t s,
                      std::vector<int64_t> &cnt, int *maxdeg, bool *proper) {
  ncol = ncol_; G = G_; g0 = g0_; g1 = g1_; lo = lo_; hi = hi_;
  const int nb = (g1 - g0) * ncol;
  int rc;
  if ((rc = dcnt.alloc((size_t) nb))) return rc;
  if (h_colour_new) CK(cudaMemcpyAsync(colour.p, h_colour_new, sizeof(int32_t) * (size_t) n, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(dcnt.p, 0, sizeof(unsigned long long) * nb, s));
  const int32_t f0[3] = {0, 0, 0x7fffffff};   // max degree, colouring faults, first spot with a colour out of range
  CK(cudaMemcpyAsync(flags.p, f0, sizeof(f0), cudaMemcpyHostToDevice, s));
  k_graph_scan<<<grid_for(n), kScanBlock, nb <= kMaxSmemBuckets ? sizeof(uint32_t) * nb : 0, s>>>(ptr.p, idx.p, colour.p, n, ncol, G, g0, nb, lo, hi, key.p, val.p,
                                                                      dcnt.p, flags.p);
  count_launch();
  CK(cudaGetLastError());
  std::vector<unsigned long long> hc((size_t) nb);
  int32_t hf[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hc.data(), dcnt.p, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hf, flags.p, sizeof(hf), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  cnt.assign(hc.begin(), hc.end());
  if (hf[1] & 2) return fail(BSB200_ERR_INVALID, "colour[%lld] out of range", (long long) hf[2]);
  *maxdeg = hf[0];
  *proper = hf[1] == 0;
  return BSB200_OK;
}

int DeviceOrder::positions(const std::vector<int64_t> &cnt, const std::vector<int64_t> &start, const std::vector<int32_t> &ghost_ids,
                           int64_t n_pos, std::vector<int32_t> *own_inv, cudaStream_t s) {
  const int nb = (g1 - g0) * ncol;
  const int64_t n_own = hi - lo;
  std::vector<int64_t> ustart((size_t) nb + 1, 0);
  for (int b = 0; b < nb; b++) ustart[(size_t) b + 1] = ustart[(size_t) b] + cnt[(size_t) b];
  DevBuf<int64_t> dstart, dustart;
  DevBuf<unsigned char> temp;
  int rc;
  if ((rc = dstart.alloc((size_t) nb)) || (rc = dustart.alloc((size_t) nb + 1)) || (rc = ghosts.alloc(ghost_ids.size()))) return rc;
  CK(cudaMemcpyAsync(dstart.p, start.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dustart.p, ustart.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, s));
  if (!ghost_ids.empty())
    CK(cudaMemcpyAsync(ghosts.p, ghost_ids.data(), sizeof(int32_t) * ghost_ids.size(), cudaMemcpyHostToDevice, s));
  int bits = 1;
  while ((1 << bits) < nb) bits++;
  size_t temp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key.p, key2.p, val.p, val2.p, (int64_t) n_own, 0, bits, s));
  if ((rc = temp.alloc(temp_bytes))) return rc;
  CK(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, key.p, key2.p, val.p, val2.p, (int64_t) n_own, 0, bits, s));   // stable
  count_launch();
  k_positions<<<grid_for(n_own), kScanBlock, 0, s>>>(key2.p, val2.p, n_own, dstart.p, dustart.p, inv.p);
  count_launch();
  if (!ghost_ids.empty()) {
    k_ghost_positions<<<grid_for((int64_t) ghost_ids.size()), kScanBlock, 0, s>>>(ghosts.p, (int64_t) ghost_ids.size(), n_pos, inv.p);
    count_launch();
  }
  CK(cudaGetLastError());
  if (own_inv) {   // the halo plan of a sharded chain (host) addresses the own spots by position
    own_inv->resize((size_t) n_own);
    CK(cudaMemcpyAsync(own_inv->data(), inv.p + lo, sizeof(int32_t) * (size_t) n_own, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));   // the temporaries go out of scope
  return BSB200_OK;
}

void DeviceOrder::release() {
  ptr.release(); idx.release(); init.release(); colour.release(); val.release(); val2.release(); inv.release();
  ghosts.release(); key.release(); key2.release(); dcnt.release(); flags.release();
}

int DeviceOrder::place(int64_t ld, int maxdeg, int32_t *d_orig, uint8_t *d_z, int32_t *d_ell, cudaStream_t s) {
  const int64_t n_own = hi - lo;
  const size_t ell_rows = (size_t) std::max(maxdeg, 1);
  CK(cudaMemsetAsync(d_orig, 0xff, sizeof(int32_t) * ((size_t) ld + kSweepPad), s));                 // -1: holes, padding
  CK(cudaMemsetAsync(d_ell, 0xff, sizeof(int32_t) * (ell_rows * (size_t) ld + kSweepPad), s));
  CK(cudaMemsetAsync(d_z, 0, (size_t) ld, s));
  CK(cudaMemsetAsync(d_z + ld, 255, (size_t) kSweepPad, s));   // the label no cluster has (padding neighbour entries)
  k_place<<<grid_for(n_own), kScanBlock, 0, s>>>(val2.p, n_own, inv.p, init.p, d_orig, d_z);
  count_launch();
  if (ghosts.count) {
    k_place<<<grid_for((int64_t) ghosts.count), kScanBlock, 0, s>>>(ghosts.p, (int64_t) ghosts.count, inv.p, init.p, d_orig, d_z);
    count_launch();
  }
  k_ell<<<grid_for(n_own), kScanBlock, 0, s>>>(ptr.p, idx.p, lo, hi, inv.p, d_ell, ld);
  count_launch();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  return BSB200_OK;
}

}   // namespace bsb
