// chain.cu -- host layer of the four cluster samplers behind the C ABI (include/bsb200.h):
// argument validation, colouring / ordering / segmentation of the spots, spot sharding over several
// GPUs, HBM staging, the per-iteration launch schedule (captured once as a CUDA graph, NCCL exchanges
// included), snapshots and error mapping.
//
// Replaces the bodies of iterate / iterate_vvv / iterate_t / iterate_t_vvv
// (/root/reference/src/cluster.cpp:217-710) as seen through their Rcpp glue
// (src/RcppExports.cpp:14-105).  Output containers follow SURVEY.md App. D.
#include "chain.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <mutex>
#include <string>
#include <thread>

#include "comm.hpp"
#include "setup_kernels.hpp"
#include "shard.hpp"
#include "tma_maps.hpp"

namespace bsb {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(BSB200_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

KernelsForD *kernel_table() {
  static KernelsForD table[kMaxD + 1] = {};
  return table;
}
KernelRegistrar::KernelRegistrar(int d, const KernelsForD &k) {
  KernelsForD &t = kernel_table()[d];
  if (k.sweep) { t.sweep = k.sweep; t.sweep_smem = k.sweep_smem; t.sweep_max_blocks_per_sm = k.sweep_max_blocks_per_sm; }
  if (k.loglik) t.loglik = k.loglik;
  if (k.deconv) { t.deconv = k.deconv; t.deconv_smem = k.deconv_smem; }
}
FastRegistrar::FastRegistrar(int d, void (*launch)(int, const SweepArgs &, int, cudaStream_t),
                             size_t (*smem)(int, const SweepArgs &)) {
  kernel_table()[d].sweep_fast = launch;
  kernel_table()[d].sweep_fast_smem = smem;
}
static bool use_fast(const KernelsForD &K, int vvv, const SweepArgs &a) {
  static const bool off = std::getenv("BSB200_NO_FAST_SWEEP") != nullptr;   // A/B switch for measurements
  return K.sweep_fast && a.maps && !vvv && a.PL.q <= kFastMaxQ && !off;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libbsb200.so does not link libcuda).
int encode_sweep_maps(const double *Y, int64_t ld, int d, const int32_t *ell, int maxdeg, const int32_t *orig,
                      SweepMaps *out, const char **msg) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
      *msg = "cuTensorMapEncodeTiled is not available from this driver";
      return 1;
    }
    encode = (encode_fn) fn;
  }
  std::memset(out, 0, sizeof(*out));
  const cuuint32_t ones[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {(cuuint64_t) ld, (cuuint64_t) d}, strides[1] = {(cuuint64_t) ld * sizeof(double)};
    const cuuint32_t box[2] = {16, (cuuint32_t) d};
    if (encode(&out->y, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(Y), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      *msg = "tensor map of Y";
      return 1;
    }
  }
  if (ell && maxdeg > 0) {
    const cuuint64_t dims[2] = {(cuuint64_t) ld, (cuuint64_t) maxdeg}, strides[1] = {(cuuint64_t) ld * sizeof(int32_t)};
    const cuuint32_t box[2] = {32, (cuuint32_t) maxdeg};
    if (encode(&out->ell, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<int32_t *>(ell), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      *msg = "tensor map of the neighbour ids";
      return 1;
    }
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t) ld, 1}, strides[1] = {(cuuint64_t) ld * sizeof(int32_t)};
    const cuuint32_t box[2] = {32, 1};
    if (encode(&out->orig, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<int32_t *>(orig), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      *msg = "tensor map of the original ids";
      return 1;
    }
  }
  return 0;
}
void launch_sweep_any(const KernelsForD &K, int tdist, int vvv, const SweepArgs &a, int grid, cudaStream_t s) {
  if (use_fast(K, vvv, a)) K.sweep_fast(tdist, a, grid, s);
  else K.sweep(tdist, vvv, a, grid, s);
}
size_t sweep_smem_any(const KernelsForD &K, int tdist, int vvv, const SweepArgs &a) {
  return use_fast(K, vvv, a) ? K.sweep_fast_smem(tdist, a) : K.sweep_smem(tdist, vvv, a);
}

size_t device_smem_limit() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return (size_t) v;
}

cudaError_t pool_alloc(void **p, size_t bytes) {
  cudaError_t e = cudaMallocAsync(p, bytes, cudaStreamPerThread);
  if (e == cudaErrorMemoryAllocation) {   // cached blocks of other sizes may be in the way: hand them back and retry
    cudaGetLastError();
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      cudaDeviceSynchronize();
      cudaMemPoolTrimTo(pool, 0);
      e = cudaMallocAsync(p, bytes, cudaStreamPerThread);
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);   // valid on every stream from here on
  return e;
}
void pool_free(void *p) {
  cudaDeviceSynchronize();   // what cudaFree implied: nothing in flight reads the block
  cudaFreeAsync(p, cudaStreamPerThread);
}

int select_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(BSB200_ERR_CUDA, "no CUDA device available (%s); libbsb200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= count) return fail(BSB200_ERR_INVALID, "device %d out of range (have %d)", device, count);
  CK(cudaSetDevice(device));
  // once per device and process: the architecture check (cudaGetDeviceProperties alone takes milliseconds, a third of
  // the set-up of a Visium-sized chain) and the pool's release threshold
  static std::mutex mu;
  static bool checked[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (device < 64 && checked[device]) return BSB200_OK;
  int major = 0, minor = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major < 10) return fail(BSB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
  cudaMemPool_t pool;   // keep freed blocks mapped (chain.hpp: pool_alloc)
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = ~0ull;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (device < 64) checked[device] = true;
  return BSB200_OK;
}

// ------------------------------------------------------------------------------------------------
// Spot ordering of the spots [lo, hi) a rank owns: internal order = (group, colour, original id);
// segments = runs of at most `seg` spots inside one (group, colour) range; slots numbered in internal
// order so that every group's partial records are contiguous.  Groups are global (virtual shards).
// ------------------------------------------------------------------------------------------------
int SpotOrder::build(int64_t n, const int32_t *colour, int ncolours, int ngroups_total, int seg_spots, int g0, int g1,
                     int64_t lo, int64_t hi, int align) {
  if (g1 < 0) { g0 = 0; g1 = ngroups_total; lo = 0; hi = n; }
  if (align < 1) align = 1;
  ncol = ncolours;
  ngroups = g1 - g0;
  const int64_t n_own = hi - lo;
  const size_t nb = (size_t) ngroups * ncol;
  std::vector<int64_t> cnt(nb, 0), start(nb + 1, 0);
  auto grp = [&](int64_t j) { return (int) ((j * (int64_t) ngroups_total) / n) - g0; };
  // counting sort by (group, colour), stable in the original id; chunks of spots counted and placed by host threads
  const int nchunk = parallel_chunks(n_own);
  std::vector<std::vector<int64_t>> ccnt((size_t) nchunk, std::vector<int64_t>(nb, 0));
  parallel_for(n_own, [&](int64_t a, int64_t b, int t) {
    std::vector<int64_t> &c = ccnt[(size_t) t];
    for (int64_t j = lo + a; j < lo + b; j++) c[(size_t) grp(j) * ncol + colour[j]]++;
  });
  for (size_t b = 0; b < nb; b++)
    for (int t = 0; t < nchunk; t++) cnt[b] += ccnt[(size_t) t][b];
  layout(cnt, seg_spots, align, start);
  perm.assign((size_t) npos, -1);
  inv.resize((size_t) n_own);
  for (size_t b = 0; b < nb; b++) {   // first position of chunk t inside bucket b
    int64_t at = start[b];
    for (int t = 0; t < nchunk; t++) { const int64_t c = ccnt[(size_t) t][b]; ccnt[(size_t) t][b] = at; at += c; }
  }
  parallel_for(n_own, [&](int64_t a, int64_t b, int t) {
    std::vector<int64_t> &fill = ccnt[(size_t) t];
    for (int64_t j = lo + a; j < lo + b; j++) {
      const int64_t pos = fill[(size_t) grp(j) * ncol + colour[j]]++;
      perm[(size_t) pos] = (int32_t) j;
      inv[(size_t) (j - lo)] = (int32_t) pos;
    }
  });
  return BSB200_OK;
}

// Positions of the (shard, colour) buckets (aligned starts, holes between them) and the segment / slot tables, from the
// bucket sizes alone: shared by the host order above and the device order (setup_kernels.cu).  ncol, ngroups set.
void SpotOrder::layout(const std::vector<int64_t> &cnt, int seg_spots, int align, std::vector<int64_t> &start) {
  const size_t nb = (size_t) ngroups * ncol;
  if (align < 1) align = 1;
  start.assign(nb + 1, 0);
  for (size_t b = 0; b < nb; b++) start[b + 1] = (start[b] + cnt[b] + align - 1) / align * align;
  npos = nb ? start[nb - 1] + cnt[nb - 1] : 0;
  seg_begin.clear(); seg_count.clear(); seg_slot.clear();
  colour_seg_start.assign((size_t) ncol + 1, 0);
  group_start.assign((size_t) ngroups + 1, 0);
  // slots in (group, colour) order; launch lists in (colour, group) order
  std::vector<std::vector<int32_t>> cb((size_t) ncol), cc((size_t) ncol), cs((size_t) ncol);
  int32_t slot = 0;
  for (int g = 0; g < ngroups; g++) {
    group_start[(size_t) g] = slot;
    for (int c = 0; c < ncol; c++) {
      const int64_t b = start[(size_t) g * ncol + c], e = b + cnt[(size_t) g * ncol + c];
      for (int64_t s = b; s < e; s += seg_spots) {
        cb[(size_t) c].push_back((int32_t) s);
        cc[(size_t) c].push_back((int32_t) std::min<int64_t>(seg_spots, e - s));
        cs[(size_t) c].push_back(slot++);
      }
    }
  }
  group_start[(size_t) ngroups] = slot;
  nslots = slot;
  for (int c = 0; c < ncol; c++) {
    colour_seg_start[(size_t) c] = (int32_t) seg_begin.size();
    seg_begin.insert(seg_begin.end(), cb[(size_t) c].begin(), cb[(size_t) c].end());
    seg_count.insert(seg_count.end(), cc[(size_t) c].begin(), cc[(size_t) c].end());
    seg_slot.insert(seg_slot.end(), cs[(size_t) c].begin(), cs[(size_t) c].end());
  }
  colour_seg_start[(size_t) ncol] = (int32_t) seg_begin.size();
}

int choose_groups(int64_t n) { return n >= 262144 ? kGroups : 1; }
int choose_segment(int64_t n, int ncol, int d) {
  // Tiles per segment (= per CTA of a sweep launch): a function of the GLOBAL problem only (n, colours, d), never of
  // the GPU count, because the fold order -- hence the bits of a chain -- depends on it.
  int64_t tps = n / ((int64_t) kBlock * 2048);
  tps = std::max<int64_t>(1, std::min<int64_t>(8, tps));
  if (choose_groups(n) == kGroups && ncol > 0) {
    // Sharded sizes: a colour phase of W ranks has (kGroups / W) * ceil(tiles / t) segments for 148 SMs x (resident
    // CTAs of k_sweep_fast<d>) persistent CTAs; its time is about rounds * (t + fixed cost of a segment, ~half a tile:
    // record fold and write -- one-tile segments measured slower than two-tile ones at 700k bins).  Pick the t whose
    // cost, summed over W = 1, 2, 4, 8, is smallest (10M spots, d = 15: t = 9 -> 8 / 4 / 2 / 1 rounds).
    const int64_t slots = 148 * (d <= 16 ? 4 : (d <= 24 ? 3 : 2));
    const int64_t tiles = ((n / kGroups + ncol - 1) / ncol + kBlock - 1) / kBlock;   // per (shard, colour)
    int64_t best = 0;
    double best_cost = 0;
    for (int64_t t = 2; t <= 12; t++) {
      const int64_t segs = (tiles + t - 1) / t;
      double cost = 0;
      for (int W = 1; W <= 8; W *= 2) cost += (double) ((kGroups / W * segs + slots - 1) / slots) * ((double) t + 0.5) * W;
      if (best == 0 || cost < best_cost) { best = t; best_cost = cost; }
    }
    tps = best;
  }
  if (const char *e = getenv("BSB200_SEG_TILES")) tps = std::max(1, atoi(e));   // tuning knob
  return (int) tps * kBlock;
}

template <typename T>
static int upload(DevBuf<T> &buf, const T *host, size_t count, cudaStream_t s) {
  int rc = buf.alloc(count);
  if (rc) return rc;
  if (count) CK(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  return BSB200_OK;
}

}   // namespace bsb

using namespace bsb;

namespace bsb {
// Page-locked snapshot staging (120 MB at 10M spots) is expensive to lock and unlock: the buffers of the last chain that
// was destroyed are kept and handed to the next one that fits (one set per process; freed with the CUDA context).
struct PinnedCache {
  std::mutex mu;
  void *z = nullptr, *w = nullptr;
  size_t count = 0;
  int device = -1;
  bool take(int dev, size_t n, int32_t **pz, double **pw) {
    std::lock_guard<std::mutex> l(mu);
    if (!z || device != dev || count < n) return false;
    *pz = (int32_t *) z; *pw = (double *) w;
    z = w = nullptr; count = 0;
    return true;
  }
  void give(int dev, size_t n, int32_t *pz, double *pw) {
    std::lock_guard<std::mutex> l(mu);
    if (z) { cudaFreeHost(z); cudaFreeHost(w); }
    z = pz; w = pw; count = n; device = dev;
  }
};
static PinnedCache g_pinned;
}   // namespace bsb

struct bsb200_chain {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t n = 0, n_own = 0, n_pos = 0, ld = 0, lo = 0, hi = 0;   // n_pos: internal positions of the own spots (holes included)
  int d = 0, q = 0, nl = 1, maxdeg = 0, G = 1;
  bool tdist = false, vvv = false;
  uint32_t compat = 0, sweep_flags = 0;
  int nrep = 0, thin = 1;
  double gamma = 0, alpha = 0, beta = 0;
  uint64_t seed = 0;
  ShardPlan plan;
  HaloPlan halo;
  Comm comm;
  SpotOrder ord;
  ParamLayout PL;
  StatLayout SL;
  KernelsForD K{};
  DevBuf<double> Y, w, params, partials, gsum, stats, T_fixed, mu0c, lambda0, lm0, ybar, plogLik, snap_mu, snap_lambda, snap_w;
  DevBuf<uint8_t> z, halo_send, halo_recv;
  DevBuf<int32_t> ell, orig, seg_begin, seg_count, seg_slot, group_start, snap_z, send_pos, recv_pos;
  DevBuf<uint32_t> iter, mode_cnt, mode_first;   // mode_*: q x n_own label histogram of the kept snapshots
  DevBuf<int> err;
  std::vector<double> ybar_host;
  int32_t *h_snap_z = nullptr;   // pinned staging of one snapshot
  double *h_snap_w = nullptr;
  size_t h_snap_count = 0;       // elements each of them holds
  cudaGraphExec_t graph_iter = nullptr, graph_batch = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t copy_stream = nullptr;                    // snapshots leave the device beside the next batch of sweeps
  cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;    // snapshot written (main stream) / on the host (copy stream)
  int64_t iters_done = 0;
  int kernels_per_iter = 0;
  double setup_ms = 0;
  int comm_rc = 0;               // first NCCL failure seen while enqueuing (reported by the next sync point)
  double *gsum_p = nullptr;      // shard sums: own buffer, or the exported window of the peer transport
  SweepMaps maps;                // TMA descriptors of Y / ell / orig
  bool fuse_colours = false;     // small single-GPU chains: all colour classes in one launch (ColourFuse)
  DevBuf<int32_t> colour_seg_start;
  DevBuf<uint32_t> colour_done;
  bool fused_halo = false;       // peer transport + k_sweep_fast: the sweep launches exchange the halos themselves
  std::vector<int> nbound;       // per colour: boundary segments (first in the launch list)
  DevBuf<uint32_t> halo_done;    // per colour: finished boundary segments of the running launch
  bool have_maps = false;

  ~bsb200_chain() {
    const bool timing = std::getenv("BSB200_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t = now();
    auto mark = [&](const char *what) {
      if (!timing) return;
      const double t2 = now();
      std::fprintf(stderr, "[bsb200 destroy] %-22s %8.1f ms\n", what, t2 - t);
      t = t2;
    };
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);   // a snapshot copy of an aborted run may still be in flight
    mark("sync");
    if (graph_iter) cudaGraphExecDestroy(graph_iter);     // graphs first: NCCL waits for the graphs that captured it
    if (graph_batch) cudaGraphExecDestroy(graph_batch);
    mark("graphs");
    comm.shutdown();               // unmap the peers before the exported label array goes away
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_snap) cudaEventDestroy(ev_snap);
    if (ev_copied) cudaEventDestroy(ev_copied);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (h_snap_z && h_snap_w) g_pinned.give(device, h_snap_count, h_snap_z, h_snap_w);
    else {
      if (h_snap_z) cudaFreeHost(h_snap_z);
      if (h_snap_w) cudaFreeHost(h_snap_w);
    }
    mark("pinned buffers");
    if (stream) cudaStreamDestroy(stream);
    Y.release(); ell.release(); mark("Y, ell");
  }

  int world() const { return plan.world; }

  SweepArgs sweep_args(int colour, uint32_t flags, uint32_t iter_add) const {
    SweepArgs a{};
    a.maps = have_maps ? &maps : nullptr;
    a.Y = Y.p; a.ld = ld; a.z = z.p; a.w = w.p; a.ell = ell.p; a.maxdeg = maxdeg; a.orig = orig.p;
    const int s0 = colour < 0 ? 0 : ord.colour_seg_start[(size_t) colour];
    const int s1 = colour < 0 ? (int) ord.seg_begin.size() : ord.colour_seg_start[(size_t) colour + 1];
    a.seg_begin = seg_begin.p + s0; a.seg_count = seg_count.p + s0; a.seg_slot = seg_slot.p + s0;
    a.nseg = s1 - s0;
    a.params = params.p; a.PL = PL; a.partials = partials.p; a.SL = SL;
    a.iter = iter.p; a.iter_add = iter_add;
    a.key0 = (uint32_t) seed; a.key1 = (uint32_t) (seed >> 32);
    a.gamma = gamma; a.flags = flags | sweep_flags; a.err = err.p;
    if (fuse_colours && colour < 0 && flags == 0) {
      a.cfuse.on = 1; a.cfuse.ncol = ord.ncol;
      a.cfuse.colour_seg_start = colour_seg_start.p; a.cfuse.done = colour_done.p;
    }
    if (fused_halo && colour >= 0 && !(flags & SW_STATS_ONLY)) {
      a.halo.legs = comm.halo_legs + (size_t) colour * (size_t) (world() - 1);
      a.halo.nlegs = comm.halo_leg_count[(size_t) colour];
      a.halo.nbound = nbound[(size_t) colour];
      a.halo.ncol = ord.ncol; a.halo.colour = colour;
      a.halo.send_pos = send_pos.p; a.halo.send_dst = comm.send_dst;
      a.halo.done = halo_done.p + colour;
      a.halo.timeout_ns = (unsigned long long) comm.peer_timeout_ns;
    }
    return a;
  }
  ParamArgs param_args(bool finalize_only) const {
    ParamArgs a{};
    a.PL = PL; a.SL = SL; a.params = params.p; a.partials = partials.p; a.group_start = group_start.p;
    a.ngroups = G; a.gsum = gsum_p; a.gsum_ready = G > 1; a.stats = stats.p;
    if (comm.peer()) { a.gepoch = comm.gather_epoch(); a.gsum_stride = comm.gsum_stride; }
    a.T_fixed = T_fixed.p; a.mu0c = mu0c.p; a.lambda0 = lambda0.p; a.lm0 = lm0.p; a.ybar = ybar.p;
    a.alpha = alpha; a.beta = beta; a.n = n; a.tdist = tdist; a.vvv = vvv; a.compat = compat;
    a.key0 = (uint32_t) seed; a.key1 = (uint32_t) (seed >> 32);
    a.iter = iter.p; a.nrep = nrep; a.thin = thin; a.plogLik = plogLik.p;
    a.snap_mu = snap_mu.p; a.snap_lambda = snap_lambda.p;
    a.finalize_only = finalize_only; a.precision_form = 0; a.err = err.p;
    if (fuse_colours) { a.zero_counters = colour_done.p; a.n_zero = ord.ncol; }
    return a;
  }
  // level 1 of the statistics fold on the shards this rank owns, allgather of the shard sums, k_param
  void enqueue_reduce_and_param(bool finalize_only, cudaStream_t s) {
    if (G > 1) {
      const bool peer = comm.peer();
      launch_reduce_groups(partials.p, group_start.p, ord.ngroups, SL.E, gsum_p + (size_t) plan.g0 * SL.E, s,
                           peer ? comm.gather_epoch() : nullptr, comm.gsum_stride);
      if (peer) {
        comm.enqueue_gather_peer((size_t) plan.g0 * SL.E, (size_t) ord.ngroups * SL.E, err.p, s);
      } else if (world() > 1) {
        int rc = comm.allgather_f64(gsum_p, (size_t) ord.ngroups * SL.E, s);
        if (rc && !comm_rc) comm_rc = rc;
      }
    }
    launch_param(param_args(finalize_only), s);
    g_launches++;
  }
  void enqueue_sweep(int colour, uint32_t flags, uint32_t iter_add, cudaStream_t s) {
    SweepArgs a = sweep_args(colour, flags, iter_add);
    if (a.nseg <= 0) return;
    launch_sweep_any(K, tdist, vvv, a, a.nseg, s);
    g_launches++;
  }
  // labels of the boundary spots of one colour travel to the ranks that hold them as ghosts
  void enqueue_halo(int colour, cudaStream_t s) {
    if (world() <= 1) return;
    if (fused_halo) {   // the sweep launch of this colour pushed and released already; an empty phase still owes its epoch
      if (ord.colour_seg_start[(size_t) colour + 1] == ord.colour_seg_start[(size_t) colour]) comm.enqueue_halo_release(colour, iter.p, s);
      return;
    }
    if (comm.peer()) {
      comm.enqueue_halo_peer(colour, z.p, send_pos.p, err.p, s);
      return;
    }
    const int W = world();
    const int32_t *so = halo.send_off.data() + (size_t) colour * W, *ro = halo.recv_off.data() + (size_t) colour * W;
    const int ns = so[W] - so[0], nr = ro[W] - ro[0];
    launch_halo_pack(z.p, send_pos.p + so[0], ns, halo_send.p + so[0], s);
    int rc = comm.exchange_u8(halo_send.p, so, halo_recv.p, ro, s);
    if (rc && !comm_rc) comm_rc = rc;
    launch_halo_unpack(z.p, recv_pos.p + ro[0], nr, halo_recv.p + ro[0], s);
  }
  void enqueue_iteration(cudaStream_t s) {
    enqueue_reduce_and_param(false, s);
    if (fuse_colours) {   // every colour class in one launch, ordered by in-kernel counters
      enqueue_sweep(-1, 0, 0, s);
      return;
    }
    for (int c = 0; c < ord.ncol; c++) {
      enqueue_sweep(c, 0, 0, s);
      enqueue_halo(c, s);
    }
  }
};

namespace bsb {

static int validate(const bsb200_cluster_args *a) {
  if (!a) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (a->n <= 0 || a->n > 0x7fffff00ll) return fail(BSB200_ERR_INVALID, "n = %lld out of range", (long long) a->n);
  if (a->d < 1 || a->d > kMaxD) return fail(BSB200_ERR_UNSUPPORTED, "d = %d unsupported (1..%d)", a->d, kMaxD);
  if (a->q < 2 || a->q > 255) return fail(BSB200_ERR_UNSUPPORTED, "q = %d unsupported (2..255)", a->q);
  if (a->nrep < 1 || a->thin < 1) return fail(BSB200_ERR_INVALID, "nrep and thin must be positive");
  if (!a->Y || !a->nbr_ptr || !a->init || !a->mu0 || !a->lambda0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (a->model != BSB200_MODEL_NORMAL && a->model != BSB200_MODEL_T) return fail(BSB200_ERR_INVALID, "unknown model %d", a->model);
  if (a->precision != BSB200_PRECISION_EQUAL && a->precision != BSB200_PRECISION_VARIABLE) return fail(BSB200_ERR_INVALID, "unknown precision %d", a->precision);
  // the first offending spot (lowest id) is reported, whichever host thread finds it
  std::vector<int64_t> first_bad((size_t) parallel_chunks(a->n), -1);
  std::vector<int> first_kind(first_bad.size(), 0);
  parallel_for(a->n, [&](int64_t lo, int64_t hi, int t) {
    for (int64_t j = lo; j < hi; j++) {
      int kind = 0;
      const int64_t dg = a->nbr_ptr[j + 1] - a->nbr_ptr[j];
      if (a->init[j] < 1 || a->init[j] > a->q) kind = 1;
      else if (dg < 0 || dg > kMaxDeg) kind = 2;
      else
        for (int64_t e = a->nbr_ptr[j]; e < a->nbr_ptr[j + 1]; e++)
          if (a->nbr_idx[e] < 0 || a->nbr_idx[e] >= a->n) { kind = 3; break; }
      if (kind) { first_bad[(size_t) t] = j; first_kind[(size_t) t] = kind; return; }
    }
  });
  for (size_t t = 0; t < first_bad.size(); t++) {
    const int64_t j = first_bad[t];
    if (j < 0) continue;
    if (first_kind[t] == 1) return fail(BSB200_ERR_INVALID, "init[%lld] = %d outside 1..q", (long long) j, a->init[j]);
    if (first_kind[t] == 2)
      return fail(BSB200_ERR_UNSUPPORTED, "spot %lld has %lld neighbours (max %d)", (long long) j,
                  (long long) (a->nbr_ptr[j + 1] - a->nbr_ptr[j]), kMaxDeg);
    return fail(BSB200_ERR_INVALID, "neighbour index out of range at spot %lld", (long long) j);
  }
  return BSB200_OK;
}

static int check_device_error(bsb200_chain *ch) {
  int e = 0;
  CK(cudaMemcpyAsync(&e, ch->err.p, sizeof(int), cudaMemcpyDeviceToHost, ch->stream));
  CK(cudaStreamSynchronize(ch->stream));
  if (ch->comm_rc) return ch->comm_rc;
  if (e == BSB_DEVERR_PEER_TIMEOUT)
    return fail(BSB200_ERR_CUDA, "a peer rank did not answer within %.0f s (lost rank, or ranks running different calls)",
                1e-9 * (double) BSB_PEER_TIMEOUT_NS);
  if (e == BSB_DEVERR_NUMERIC)
    return fail(BSB200_ERR_NUMERIC, "NA in probability vector, empty cluster (Wishart df <= 0) or matrix not positive definite");
  return BSB200_OK;
}

static int build_graphs(bsb200_chain *ch) {
  cudaGraph_t g = nullptr;
  const long long launches_before = g_launches.load();
  CK(cudaStreamBeginCapture(ch->stream, cudaStreamCaptureModeThreadLocal));
  ch->enqueue_iteration(ch->stream);
  CK(cudaStreamEndCapture(ch->stream, &g));
  ch->kernels_per_iter = (int) (g_launches.load() - launches_before);   // kernel nodes of one iteration
  CK(cudaGraphInstantiate(&ch->graph_iter, g, 0));
  cudaGraphDestroy(g);
  if (ch->thin > 1 && ch->thin <= 256 && ch->n <= 200000 && ch->world() == 1) {   // launch-bound sizes: one graph per thin batch
    CK(cudaStreamBeginCapture(ch->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < ch->thin; i++) ch->enqueue_iteration(ch->stream);
    CK(cudaStreamEndCapture(ch->stream, &g));
    CK(cudaGraphInstantiate(&ch->graph_batch, g, 0));
    cudaGraphDestroy(g);
  }
  g_launches.store(launches_before);   // captured nodes are counted when the graphs are launched
  if (ch->comm_rc) return ch->comm_rc;
  return BSB200_OK;
}

int chain_create_impl(const bsb200_cluster_args *a, bsb200_chain **out) {
  const double t0 = now_ms();
  // BSB200_TIMING=1: wall-clock marks of the set-up phases on stderr
  static const bool timing = std::getenv("BSB200_TIMING") != nullptr;
  double t_last = t0;
  auto mark = [&](const char *what) {
    if (!timing) return;
    const double t = now_ms();
    std::fprintf(stderr, "[bsb200 setup r%d] %-22s %8.1f ms\n", a->world > 1 ? a->rank : 0, what, t - t_last);
    t_last = t;
  };
  set_host_share(a && a->world > 1 ? a->world : 1);
  int rc = validate(a);
  if (rc) return rc;
  rc = select_device(a->device);
  if (rc) return rc;
  KernelsForD K = kernel_table()[a->d];
  if (!K.sweep) return fail(BSB200_ERR_UNSUPPORTED, "library was built without kernels for d = %d", a->d);

  std::unique_ptr<bsb200_chain> ch(new bsb200_chain());
  ch->device = a->device; ch->K = K;
  ch->n = a->n; ch->d = a->d; ch->q = a->q;
  ch->tdist = a->model == BSB200_MODEL_T; ch->vvv = a->precision == BSB200_PRECISION_VARIABLE;
  ch->nl = ch->vvv ? a->q : 1;
  ch->compat = a->compat; ch->nrep = a->nrep; ch->thin = a->thin;
  ch->gamma = a->gamma; ch->alpha = a->alpha; ch->beta = a->beta; ch->seed = a->seed;
  const int64_t n = a->n;
  const int d = a->d, q = a->q;
  if ((a->compat & BSB200_COMPAT_Q3) && ch->vvv && !ch->tdist) ch->sweep_flags |= SW_Q3;
  if (!ch->tdist && !ch->vvv) ch->sweep_flags |= SW_NO_PAIRS;

  // ---- sharding ----------------------------------------------------------------------------------
  const int world = a->world > 1 ? a->world : 1, rank = world > 1 ? a->rank : 0;
  const int G = a->groups > 0 ? a->groups : (world > 1 ? kGroups : choose_groups(n));
  if (G > 4096) return fail(BSB200_ERR_INVALID, "groups = %d out of range", G);
  rc = make_shard_plan(n, a->nbr_ptr, a->nbr_idx, G, world, rank, ch->plan);
  if (rc) return rc;
  mark("validate+shard plan");
  const ShardPlan &plan = ch->plan;
  ch->G = G; ch->lo = plan.lo; ch->hi = plan.hi; ch->n_own = plan.n_own();
  const int64_t n_own = ch->n_own;
  // The PCs are the bulk of the input (1.2 GB at 10M spots): their upload runs on a second host thread while this one
  // colours, orders and stages the graph.  Own rows of Y (R's column-major n x d): one strided copy.
  // (everything the upload thread touches is declared before the thread's owner, i.e. outlives its join on any return)
  DevBuf<double> Yraw;
  DeviceOrder dorder;
  std::promise<int> csr_promise;
  std::future<int> csr_ready = csr_promise.get_future();
  struct Uploader {
    std::thread th;
    int rc = BSB200_OK;
    std::string msg;
    ~Uploader() { if (th.joinable()) th.join(); }
  } yup;
  struct Pinner {   // page-locked snapshot staging (120 MB at 10M spots), allocated off the critical path
    std::thread th;
    cudaError_t e = cudaSuccess;
    ~Pinner() { if (th.joinable()) th.join(); }
  } pin;
  pin.th = std::thread([&, device = a->device, chp = ch.get()]() {
    pin.e = cudaSetDevice(device);
    const size_t cnt = (size_t) std::max<int64_t>(n_own, 1);
    chp->h_snap_count = cnt;
    if (pin.e == cudaSuccess && g_pinned.take(device, cnt, &chp->h_snap_z, &chp->h_snap_w)) return;
    if (pin.e == cudaSuccess) pin.e = cudaMallocHost(&chp->h_snap_z, sizeof(int32_t) * cnt);
    if (pin.e == cudaSuccess) pin.e = cudaMallocHost(&chp->h_snap_w, sizeof(double) * cnt);
  });
  // The spots are ordered and the neighbour lists built on the device (setup_kernels.cu) from the CSR arrays, which
  // therefore go first; BSB200_HOST_SETUP=1 keeps the host version (same order, tests/test_gpu_zv_setup.py).
  const bool dev_setup = !std::getenv("BSB200_HOST_SETUP");
  yup.th = std::thread([&, device = a->device]() {
    cudaError_t e = cudaSetDevice(device);
    if (dev_setup) {
      int rcu = e == cudaSuccess ? dorder.upload(a->nbr_ptr, a->nbr_idx, a->init, a->colour, n, device) : BSB200_ERR_CUDA;
      if (rcu) yup.msg = g_last_error;   // thread-local message of this thread
      csr_promise.set_value(rcu);
      if (rcu) { yup.rc = rcu; return; }
    }
    if (e == cudaSuccess && Yraw.alloc((size_t) n_own * d)) { yup.rc = BSB200_ERR_CUDA; yup.msg = g_last_error; return; }
    if (e == cudaSuccess) {
      std::vector<StagedCopy> cols;   // own rows of every column of R's column-major n x d matrix
      if (n_own == n) cols.push_back({Yraw.p, a->Y, sizeof(double) * (size_t) n * d});
      else
        for (int c = 0; c < d; c++) cols.push_back({Yraw.p + (size_t) c * n_own, a->Y + (size_t) c * n + plan.lo, sizeof(double) * (size_t) n_own});
      const int rcu = staged_h2d(device, cols.data(), (int) cols.size());
      if (rcu) { yup.rc = rcu; yup.msg = g_last_error; }
    }
    if (e != cudaSuccess) { yup.rc = BSB200_ERR_CUDA; yup.msg = cudaGetErrorString(e); }
    if (timing) std::fprintf(stderr, "[bsb200 setup] Y upload thread done at %8.1f ms\n", now_ms() - t0);
  });

  // ---- colouring (of the whole graph: every rank derives the same classes) ------------------------
  std::vector<int32_t> colour;
  int32_t ncol = 0;
  std::vector<int64_t> bucket_cnt, bucket_start;
  int dev_maxdeg = 0;
  if (a->colour && (a->ncolours < 1 || a->ncolours > 64)) return fail(BSB200_ERR_INVALID, "ncolours = %d out of range", a->ncolours);
  if (dev_setup) {
    CK(cudaStreamCreateWithFlags(&ch->stream, cudaStreamNonBlocking));
    if ((rc = dorder.prealloc(n, n_own))) { csr_ready.wait(); return rc; }
    if ((rc = csr_ready.get())) return fail(rc, "%s", yup.msg.c_str());
    mark("graph upload");
    bool proper = false;
    if (a->colour) {   // range and properness are checked by the scan
      ncol = a->ncolours;
      if ((rc = dorder.scan(nullptr, ncol, G, plan.g0, plan.g1, plan.lo, plan.hi, ch->stream, bucket_cnt, &dev_maxdeg, &proper))) return rc;
    }
    if (!proper) {   // none given, or e.g. hex offsets on a filled rectangle: colour the graph ourselves
      rc = colour_graph(a->nbr_ptr, a->nbr_idx, n, colour, &ncol);
      if (rc) return rc;
      if ((rc = dorder.scan(colour.data(), ncol, G, plan.g0, plan.g1, plan.lo, plan.hi, ch->stream, bucket_cnt, &dev_maxdeg, &proper))) return rc;
      if (!proper) return fail(BSB200_ERR_NEIGHBORS, "internal error: greedy colouring is not proper");
    }
    mark("colouring");
    ch->ord.ncol = ncol;
    ch->ord.ngroups = plan.g1 - plan.g0;
    ch->ord.layout(bucket_cnt, choose_segment(n, ncol, d), kOrderAlign, bucket_start);
    // own spots into their buckets, ghosts behind them; a sharded chain's halo plan (host) reads the own positions
    if ((rc = dorder.positions(bucket_cnt, bucket_start, plan.ghosts, ch->ord.npos, world > 1 ? &ch->ord.inv : nullptr, ch->stream)))
      return rc;
  } else {
    if (a->colour) {
      colour.assign(a->colour, a->colour + n);
      ncol = a->ncolours;
      for (int64_t j = 0; j < n; j++)
        if (colour[(size_t) j] < 0 || colour[(size_t) j] >= ncol) return fail(BSB200_ERR_INVALID, "colour[%lld] out of range", (long long) j);
      if (!colouring_is_proper(a->nbr_ptr, a->nbr_idx, n, colour.data())) {
        colour.clear();   // e.g. hex offsets on a filled rectangle: colour the graph ourselves
      }
    }
    if (colour.empty()) {
      rc = colour_graph(a->nbr_ptr, a->nbr_idx, n, colour, &ncol);
      if (rc) return rc;
    }
    mark("colouring");
    ch->ord.build(n, colour.data(), ncol, G, choose_segment(n, ncol, d), plan.g0, plan.g1, plan.lo, plan.hi, kOrderAlign);
  }
  const SpotOrder &ord = ch->ord;
  const int64_t n_pos = ch->n_pos = ord.npos;   // ghosts follow the own positions
  ch->ld = (n_pos + (int64_t) plan.ghosts.size() + 31) / 32 * 32;
  mark("spot order");
  if (world > 1) {
    rc = make_halo_plan(plan, a->nbr_ptr, a->nbr_idx, colour.empty() ? a->colour : colour.data(), ncol, ord.inv, n_pos, ch->halo);
    if (rc) return rc;
    mark("halo plan");
    rc = ch->comm.init(world, rank, a->comm_id);
    if (rc) return rc;
    mark("comm init");
    // Fused halo exchange (layout.hpp HaloFuse): segments that hold boundary spots move to the front of their
    // colour's launch list.  The list order does not enter the arithmetic (records are addressed by slot).
    static const bool no_fuse = std::getenv("BSB200_NO_FUSED_HALO") != nullptr;
    ch->fused_halo = ch->comm.peer() && !no_fuse && K.sweep_fast && !ch->vvv && q <= kFastMaxQ &&
                     !std::getenv("BSB200_NO_FAST_SWEEP");
    if (ch->fused_halo) {
      SpotOrder &o = ch->ord;
      ch->nbound.assign((size_t) ncol, 0);
      for (int c = 0; c < ncol; c++) {
        const int s0 = o.colour_seg_start[(size_t) c], s1 = o.colour_seg_start[(size_t) c + 1];
        if (s1 == s0) continue;
        std::vector<char> isb((size_t) (s1 - s0), 0);
        for (int p = 0; p < world; p++) {
          const size_t e = (size_t) c * world + p;
          for (int32_t k = ch->halo.send_off[e]; k < ch->halo.send_off[e + 1]; k++) {
            const int32_t pos = ch->halo.send_pos[(size_t) k];
            // segments of a colour are sorted by their first position
            const auto it2 = std::upper_bound(o.seg_begin.begin() + s0, o.seg_begin.begin() + s1, pos);
            const int sidx = (int) (it2 - (o.seg_begin.begin() + s0)) - 1;
            if (sidx >= 0) isb[(size_t) sidx] = 1;
          }
        }
        std::vector<int32_t> nb2, nc2, ns2;
        int nbnd = 0;
        for (int pass = 0; pass < 2; pass++)
          for (int i = 0; i < s1 - s0; i++)
            if ((isb[(size_t) i] != 0) == (pass == 0)) {
              nb2.push_back(o.seg_begin[(size_t) (s0 + i)]); nc2.push_back(o.seg_count[(size_t) (s0 + i)]); ns2.push_back(o.seg_slot[(size_t) (s0 + i)]);
              if (pass == 0) nbnd++;
            }
        std::copy(nb2.begin(), nb2.end(), o.seg_begin.begin() + s0);
        std::copy(nc2.begin(), nc2.end(), o.seg_count.begin() + s0);
        std::copy(ns2.begin(), ns2.end(), o.seg_slot.begin() + s0);
        ch->nbound[(size_t) c] = std::max(nbnd, 1);   // no boundary spot of this colour: the first segment keeps the handshake
      }
    }
  }

  // ---- device staging --------------------------------------------------------------------------
  if (!ch->stream) CK(cudaStreamCreateWithFlags(&ch->stream, cudaStreamNonBlocking));
  cudaStream_t s = ch->stream;
  CK(cudaEventCreate(&ch->ev0));
  CK(cudaEventCreate(&ch->ev1));
  CK(cudaStreamCreateWithFlags(&ch->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ch->ev_snap, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ch->ev_copied, cudaEventDisableTiming));
  const int nlT = (ch->tdist || ch->vvv) ? ch->nl : 0;
  const StatLayout SL_run = make_stat_layout(q, d, nlT);
  const StatLayout SL_init = make_stat_layout(q, d, 1);   // one-off pass that also yields sum y y^T
  const size_t Emax = (size_t) std::max(SL_run.E, SL_init.E);
  if (ch->comm.peer()) {   // labels, receive list and shard sums live in the arena the peers map
    if ((rc = ch->comm.alloc_arena((size_t) ch->ld + kSweepPad, ch->halo.recv_pos.size(), (size_t) G * Emax, s))) return rc;
    ch->z.adopt(ch->comm.z(), (size_t) ch->ld);
    ch->recv_pos.adopt(ch->comm.recv_pos(), ch->halo.recv_pos.size());
  }
  int maxdeg = dev_maxdeg;
  if (!dev_setup) {
    std::vector<int> md((size_t) parallel_chunks(n), 0);
    parallel_for(n, [&](int64_t lo, int64_t hi, int t) {
      int m = 0;
      for (int64_t j = lo; j < hi; j++) m = std::max<int>(m, (int) (a->nbr_ptr[j + 1] - a->nbr_ptr[j]));
      md[(size_t) t] = m;
    });
    for (int m : md) maxdeg = std::max(maxdeg, m);
  }
  ch->maxdeg = maxdeg;
  if (dev_setup) {
    // original ids, initial labels and neighbour lists in internal order, written by the device from the uploaded graph
    const size_t ell_rows = (size_t) std::max(maxdeg, 1), ldp = (size_t) ch->ld;
    if ((rc = ch->ell.alloc(ell_rows * ldp + kSweepPad)) || (rc = ch->orig.alloc(ldp + kSweepPad))) return rc;
    if (!ch->comm.peer() && (rc = ch->z.alloc(ldp + kSweepPad))) return rc;   // peer transport: the labels live in the arena
    if ((rc = dorder.place(ch->ld, maxdeg, ch->orig.p, ch->z.p, ch->ell.p, s))) return rc;
    dorder.release();   // 0.6 GB of graph copies and sort buffers at 10M spots go back to the pool before Y is laid out
  } else {
    // neighbour ids in internal order (ELL, -1 padded), initial labels and original ids, filled by host threads
    const size_t ell_rows = (size_t) std::max(maxdeg, 1), ldp = (size_t) ch->ld;
    std::unique_ptr<int32_t[]> ell(new int32_t[ell_rows * ldp + kSweepPad]);
    // z[ld ..] = 255: the label that matches no cluster, read through padding entries of the neighbour lists
    std::unique_ptr<uint8_t[]> z0(new uint8_t[ldp + kSweepPad]);
    std::memset(z0.get() + ldp, 255, (size_t) kSweepPad);
    std::unique_ptr<int32_t[]> origp(new int32_t[ldp + kSweepPad]);
    parallel_for((int64_t) ldp, [&](int64_t i0, int64_t i1, int) {
      for (int64_t i = i0; i < i1; i++) {
        const int64_t j = i < n_pos ? (int64_t) ord.perm[(size_t) i] : (int64_t) -1;
        int e = 0;
        if (j >= 0) {
          z0[(size_t) i] = (uint8_t) (a->init[j] - 1);
          origp[(size_t) i] = (int32_t) j;
          for (int64_t p = a->nbr_ptr[j]; p < a->nbr_ptr[j + 1]; p++, e++) {
            const int32_t u = a->nbr_idx[p];
            int32_t pos;
            if (u >= plan.lo && u < plan.hi) pos = ord.inv[(size_t) (u - plan.lo)];
            else pos = (int32_t) (n_pos + (std::lower_bound(plan.ghosts.begin(), plan.ghosts.end(), u) - plan.ghosts.begin()));
            ell[(size_t) e * ldp + (size_t) i] = pos;
          }
        } else {   // hole of the internal order, ghost or padding
          z0[(size_t) i] = 0;
          origp[(size_t) i] = -1;
        }
        for (; e < (int) ell_rows; e++) ell[(size_t) e * ldp + (size_t) i] = -1;
      }
    });
    for (size_t e = 0; e < (size_t) kSweepPad; e++) { ell[ell_rows * ldp + e] = -1; origp[ldp + e] = -1; }
    for (size_t gi = 0; gi < plan.ghosts.size(); gi++) {
      z0[(size_t) n_pos + gi] = (uint8_t) (a->init[plan.ghosts[gi]] - 1);
      origp[(size_t) n_pos + gi] = plan.ghosts[gi];
    }
    if ((rc = upload(ch->ell, ell.get(), ell_rows * ldp + kSweepPad, s))) return rc;
    if (ch->comm.peer()) CK(cudaMemcpyAsync(ch->z.p, z0.get(), ldp + kSweepPad, cudaMemcpyHostToDevice, s));
    else if ((rc = upload(ch->z, z0.get(), ldp + kSweepPad, s))) return rc;
    if ((rc = upload(ch->orig, origp.get(), ldp + kSweepPad, s))) return rc;
    CK(cudaStreamSynchronize(s));   // host vectors go out of scope
  }
  mark("ELL/z/orig staging");
  if ((rc = upload(ch->seg_begin, ord.seg_begin.data(), ord.seg_begin.size(), s))) return rc;
  if ((rc = upload(ch->seg_count, ord.seg_count.data(), ord.seg_count.size(), s))) return rc;
  if ((rc = upload(ch->seg_slot, ord.seg_slot.data(), ord.seg_slot.size(), s))) return rc;
  if ((rc = upload(ch->group_start, ord.group_start.data(), ord.group_start.size(), s))) return rc;
  if (world > 1) {
    if ((rc = upload(ch->send_pos, ch->halo.send_pos.data(), ch->halo.send_pos.size(), s))) return rc;
    if (ch->comm.peer()) {
      if (!ch->halo.recv_pos.empty())
        CK(cudaMemcpyAsync(ch->recv_pos.p, ch->halo.recv_pos.data(), sizeof(int32_t) * ch->halo.recv_pos.size(),
                           cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));
    } else if ((rc = upload(ch->recv_pos, ch->halo.recv_pos.data(), ch->halo.recv_pos.size(), s))) {
      return rc;
    }
    if (!ch->comm.peer() &&
        ((rc = ch->halo_send.alloc(ch->halo.send_pos.size())) || (rc = ch->halo_recv.alloc(ch->halo.recv_pos.size()))))
      return rc;
  }
  {
    // column means from per-shard sums (the upload of Y itself was started above)
    DevBuf<double> gcol;
    DevBuf<int64_t> glo;
    if ((rc = gcol.alloc((size_t) G * d)) || (rc = ch->ybar.alloc((size_t) d)) || (rc = ch->Y.alloc((size_t) d * ch->ld + kSweepPad)))
      return rc;
    yup.th.join();
    if (yup.rc) return fail(BSB200_ERR_CUDA, "upload of Y failed: %s", yup.msg.c_str());
    if ((rc = upload(glo, plan.group_lo.data(), plan.group_lo.size(), s))) return rc;
    CK(cudaMemsetAsync(ch->Y.p, 0, sizeof(double) * ((size_t) d * ch->ld + kSweepPad), s));
    CK(cudaMemsetAsync(gcol.p, 0, sizeof(double) * G * d, s));
    launch_colsums_groups(Yraw.p, n_own, glo.p, plan.lo, d, plan.g0, plan.g1 - plan.g0, gcol.p, s);
    if (world > 1 && !ch->comm.peer() && (rc = ch->comm.allgather_f64(gcol.p, (size_t) (plan.g1 - plan.g0) * d, s))) return rc;
    std::vector<double> gs((size_t) G * d);
    CK(cudaMemcpyAsync(gs.data(), gcol.p, sizeof(double) * G * d, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ch->comm.peer()) {   // the shard sums of the other ranks come through the host rendezvous
      const size_t cnt = (size_t) (plan.g1 - plan.g0) * d;
      std::vector<double> mine(gs.begin() + (size_t) plan.g0 * d, gs.begin() + (size_t) plan.g0 * d + cnt);
      if ((rc = ch->comm.host_allgather(mine.data(), cnt * sizeof(double), gs.data()))) return rc;
    }
    ch->ybar_host.assign((size_t) d, 0.0);
    for (int c = 0; c < d; c++) {
      double acc = 0.0;
      for (int g = 0; g < G; g++) acc += gs[(size_t) g * d + c];   // shard order: independent of the GPU count
      ch->ybar_host[(size_t) c] = acc / (double) n;
    }
    CK(cudaMemcpyAsync(ch->ybar.p, ch->ybar_host.data(), sizeof(double) * d, cudaMemcpyHostToDevice, s));
    launch_permute_center(Yraw.p, n_own, n_pos, d, ch->orig.p, ch->ybar.p, ch->Y.p, ch->ld, s, plan.lo);
    CK(cudaStreamSynchronize(s));
  }
  mark("Y upload+centre");
  {
    if ((rc = ch->w.alloc((size_t) ch->ld))) return rc;
    launch_fill_f64(ch->w.p, ch->ld, 1.0, s);
    std::vector<double> mu0c((size_t) d), lm0((size_t) d, 0.0);
    for (int c = 0; c < d; c++) mu0c[(size_t) c] = a->mu0[c] - ch->ybar_host[(size_t) c];
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) lm0[(size_t) i] += a->lambda0[(size_t) j * d + i] * mu0c[(size_t) j];
    if ((rc = upload(ch->mu0c, mu0c.data(), (size_t) d, s))) return rc;
    if ((rc = upload(ch->lm0, lm0.data(), (size_t) d, s))) return rc;
    if ((rc = upload(ch->lambda0, a->lambda0, (size_t) d * d, s))) return rc;
    ch->PL = make_param_layout(q, d, ch->nl);
    std::vector<double> p0((size_t) ch->PL.total, 0.0);
    for (int k = 0; k < q; k++)
      for (int c = 0; c < d; c++) p0[(size_t) ch->PL.mu + k * d + c] = mu0c[(size_t) c];
    for (int k = 0; k < ch->nl; k++) std::memcpy(&p0[(size_t) ch->PL.lam + (size_t) k * d * d], a->lambda0, sizeof(double) * d * d);
    if ((rc = upload(ch->params, p0.data(), p0.size(), s))) return rc;
    CK(cudaStreamSynchronize(s));
  }
  const int nsnap = a->nrep / a->thin + 1;
  if ((rc = ch->partials.alloc((size_t) std::max(ord.nslots, 1) * Emax))) return rc;
  if (ch->comm.peer()) {
    if ((rc = ch->comm.connect(ch->halo, ncol, s))) return rc;
    ch->gsum_p = ch->comm.gsum();
    mark("peer windows");
  } else {
    if ((rc = ch->gsum.alloc((size_t) G * Emax))) return rc;
    ch->gsum_p = ch->gsum.p;
    CK(cudaMemsetAsync(ch->gsum.p, 0, sizeof(double) * G * Emax, s));
  }
  if ((rc = ch->stats.alloc(Emax))) return rc;
  if ((rc = ch->T_fixed.alloc((size_t) ch->PL.np))) return rc;
  if ((rc = ch->plogLik.alloc((size_t) a->nrep))) return rc;
  if ((rc = ch->snap_mu.alloc((size_t) nsnap * q * d))) return rc;
  if ((rc = ch->snap_lambda.alloc((size_t) nsnap * ch->nl * d * d))) return rc;
  if ((rc = ch->snap_z.alloc((size_t) n_own))) return rc;
  if ((rc = ch->snap_w.alloc((size_t) n_own))) return rc;
  if ((rc = ch->iter.alloc(1))) return rc;
  if ((rc = ch->halo_done.alloc((size_t) std::max(ncol, 1)))) return rc;
  CK(cudaMemsetAsync(ch->halo_done.p, 0, sizeof(uint32_t) * std::max(ncol, 1), s));
  if ((rc = ch->err.alloc(1))) return rc;
  CK(cudaMemsetAsync(ch->iter.p, 0, sizeof(uint32_t), s));
  CK(cudaMemsetAsync(ch->err.p, 0, sizeof(int), s));
  CK(cudaMemsetAsync(ch->snap_mu.p, 0, sizeof(double) * nsnap * q * d, s));
  CK(cudaMemsetAsync(ch->snap_lambda.p, 0, sizeof(double) * nsnap * ch->nl * d * d, s));
  {
    std::vector<double> nanv((size_t) a->nrep, std::numeric_limits<double>::quiet_NaN());
    CK(cudaMemcpyAsync(ch->plogLik.p, nanv.data(), sizeof(double) * a->nrep, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  pin.th.join();
  if (pin.e != cudaSuccess) return fail(BSB200_ERR_CUDA, "pinned snapshot staging: %s", cudaGetErrorString(pin.e));

  {
    const char *msg = "";
    if (encode_sweep_maps(ch->Y.p, ch->ld, d, ch->ell.p, maxdeg, ch->orig.p, &ch->maps, &msg))
      return fail(BSB200_ERR_CUDA, "cannot encode the TMA descriptors (%s)", msg);
    ch->have_maps = true;
  }
  {   // all colour classes in one launch when every CTA of that launch is resident at once
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
    static const bool off = std::getenv("BSB200_NO_FUSED_COLOURS") != nullptr;
    ch->fuse_colours = !off && world == 1 && ncol > 1 && K.sweep_fast && !ch->vvv && q <= kFastMaxQ &&
                       !std::getenv("BSB200_NO_FAST_SWEEP") && (int64_t) ord.seg_begin.size() <= (int64_t) sms * (d <= 24 ? 2 : 1);
    if (ch->fuse_colours) {
      if ((rc = upload(ch->colour_seg_start, ord.colour_seg_start.data(), ord.colour_seg_start.size(), s)) ||
          (rc = ch->colour_done.alloc((size_t) ncol)))
        return rc;
      CK(cudaMemsetAsync(ch->colour_done.p, 0, sizeof(uint32_t) * ncol, s));
    }
  }
  mark("buffers");
  // ---- the configuration must fit the shared memory of one CTA (sweep: per-cluster parameter blocks + tiles;
  //      k_param: 3 q d + 4 d^2 doubles) -- say so instead of failing inside a launch
  {
    ch->SL = make_stat_layout(q, d, std::max(nlT, 1));
    const size_t limit = device_smem_limit();
    const size_t need_sweep = sweep_smem_any(K, ch->tdist, ch->vvv, ch->sweep_args(0, 0, 0));
    const size_t need_param = param_smem_min(q, d);
    if (limit && (need_sweep > limit || need_param > std::min<size_t>(limit, 220 * 1024)))
      return fail(BSB200_ERR_UNSUPPORTED,
                  "d = %d, q = %d, %s / %s precision needs %zu (sweep) / %zu (parameter update) bytes of shared memory per "
                  "CTA, the device offers %zu: reduce q or d", d, q, ch->tdist ? "t" : "normal", ch->vvv ? "variable" : "equal",
                  need_sweep, need_param, limit);
  }
  // ---- statistics of the initial state -----------------------------------------------------------
  param_kernel_init();
  if (nlT == 0) {   // normal model, shared precision: sum_j y_j y_j^T never changes -> T_fixed
    ch->SL = SL_init;
    const uint32_t keep = ch->sweep_flags;
    ch->sweep_flags &= ~SW_NO_PAIRS;
    ch->enqueue_sweep(-1, SW_STATS_ONLY, 0, s);
    ch->sweep_flags = keep;
    ch->enqueue_reduce_and_param(true, s);
    CK(cudaMemcpyAsync(ch->T_fixed.p, ch->stats.p + SL_init.T, sizeof(double) * ch->PL.np, cudaMemcpyDeviceToDevice, s));
  }
  ch->SL = SL_run;
  ch->enqueue_sweep(-1, SW_STATS_ONLY, 0, s);
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (ch->comm_rc) return ch->comm_rc;
  mark("initial statistics");
  if ((rc = build_graphs(ch.get()))) return rc;
  mark("graph capture");
  ch->setup_ms = now_ms() - t0;
  *out = ch.release();
  return BSB200_OK;
}

// Runs `iters` iterations.  When `out` is given, snapshots go to its arrays (cluster_impl); the
// resident-chain interface passes nullptr and snapshots stop at the pinned staging buffers.
int chain_run_impl(bsb200_chain *ch, int iters, bsb200_cluster_out *out, double *device_ms, bool honour_cancel) {
  CK(cudaSetDevice(ch->device));
  cudaStream_t s = ch->stream;
  double total_ms = 0.0;
  int64_t remaining = iters;
  // A snapshot staged in pinned memory is copied into the caller's matrices while the GPU already runs the next batch
  // (the copy touches fresh pages of a 1.3 GB allocation at 10M spots: ~30 ms per row that used to idle the GPU).
  int64_t pending_row = -1;
  int sink_rc = BSB200_OK;
  bool copy_in_flight = false;
  auto flush_pending = [&]() {
    if (pending_row < 0) return;
    if (copy_in_flight) { cudaEventSynchronize(ch->ev_copied); copy_in_flight = false; }
    const size_t r = (size_t) pending_row;
    // host threads: the destination is fresh pages of the caller's matrices (page faults, not bandwidth)
    parallel_for(ch->n_own, [&](int64_t a, int64_t b, int) {
      if (out->z) std::memcpy(out->z + r * ch->n + ch->lo + a, ch->h_snap_z + a, sizeof(int32_t) * (size_t) (b - a));
      if (out->weights && ch->tdist) std::memcpy(out->weights + r * ch->n + ch->lo + a, ch->h_snap_w + a, sizeof(double) * (size_t) (b - a));
    });
    if (sink_active() && ch->world() == 1) {   // chain persistence: the row goes to the sink as it is produced
      const int64_t dims[1] = {ch->n};
      int rc = sink_write("z", (int64_t) r, ch->h_snap_z, ch->n, 0, dims, 1);
      if (!rc && ch->tdist) rc = sink_write("weights", (int64_t) r, ch->h_snap_w, ch->n, 1, dims, 1);
      if (rc && !sink_rc) sink_rc = rc;
    }
    out->rows_done = (int32_t) r + 1;
    pending_row = -1;
  };
  while (remaining > 0) {
    const int64_t it0 = ch->iters_done;   // iterations completed; next iteration index = it0 + 1
    const int64_t next_multiple = ((it0 + 2 + ch->thin - 1) / ch->thin) * ch->thin;
    const int64_t to_boundary = next_multiple - 1 - it0;   // iterations up to and including the snapshot one
    const int64_t run = std::min<int64_t>(to_boundary, remaining);
    CK(cudaEventRecord(ch->ev0, s));
    if (run == ch->thin && ch->graph_batch) {
      CK(cudaGraphLaunch(ch->graph_batch, s));
    } else {
      for (int64_t i = 0; i < run; i++) CK(cudaGraphLaunch(ch->graph_iter, s));
    }
    g_launches += (long long) ch->kernels_per_iter * run;
    ch->iters_done += run;
    remaining -= run;
    flush_pending();   // host copy of the previous snapshot, overlapped with the batch just enqueued
    const bool boundary = (ch->iters_done + 1) % ch->thin == 0;
    if (boundary) {
      const int64_t row = (ch->iters_done + 1) / ch->thin;
      const bool sink = sink_active() && out && ch->world() == 1;
      const bool want_w = ch->tdist && (!out || out->weights || sink);
      // the previous snapshot must have left snap_z / snap_w (flush_pending above waited for it when a caller takes rows)
      if (copy_in_flight) CK(cudaStreamWaitEvent(s, ch->ev_copied, 0));
      launch_snapshot(ch->z.p, want_w ? ch->w.p : nullptr, ch->orig.p, ch->n_pos, ch->snap_z.p,
                      want_w ? ch->snap_w.p : nullptr, s, ch->lo);
      if (out && out->z_mode && row >= out->mode_from && row < ch->nrep / ch->thin + 1) {
        if (!ch->mode_cnt.p) {
          int rc;
          if ((rc = ch->mode_cnt.alloc((size_t) ch->q * ch->n_own)) || (rc = ch->mode_first.alloc((size_t) ch->q * ch->n_own)))
            return rc;
          CK(cudaMemsetAsync(ch->mode_cnt.p, 0, sizeof(uint32_t) * ch->q * ch->n_own, s));
        }
        launch_mode_accumulate(ch->snap_z.p, ch->n_own, ch->q, (uint32_t) row, ch->mode_cnt.p, ch->mode_first.p, s);
      }
      // the snapshots travel to the host only when the caller asked for the matrices
      // ... on the copy stream, so that the next batch of sweeps does not wait for 120 MB over PCIe
      CK(cudaEventRecord(ch->ev_snap, s));
      CK(cudaStreamWaitEvent(ch->copy_stream, ch->ev_snap, 0));
      if (!out || out->z || sink) CK(cudaMemcpyAsync(ch->h_snap_z, ch->snap_z.p, sizeof(int32_t) * ch->n_own, cudaMemcpyDeviceToHost, ch->copy_stream));
      if (want_w) CK(cudaMemcpyAsync(ch->h_snap_w, ch->snap_w.p, sizeof(double) * ch->n_own, cudaMemcpyDeviceToHost, ch->copy_stream));
      CK(cudaEventRecord(ch->ev_copied, ch->copy_stream));
      copy_in_flight = true;
    }
    CK(cudaEventRecord(ch->ev1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ch->ev0, ch->ev1));
    total_ms += ms;
    if (boundary) {
      const size_t r = (size_t) ((ch->iters_done + 1) / ch->thin);
      if (out && r < (size_t) (ch->nrep / ch->thin + 1)) pending_row = (int64_t) r;
      int rc = check_device_error(ch);
      if (rc) { flush_pending(); return rc; }
    }
    // early_stop at a thin boundary (src/cluster.cpp:1106) / user interrupt (checkUserInterrupt, :243-244): polled
    // after every batch of at most `thin` iterations
    if (honour_cancel && cancel_requested()) {
      flush_pending();
      if (device_ms) *device_ms = total_ms;
      return fail(BSB200_ERR_CANCELLED, "Stopping...");
    }
  }
  flush_pending();
  if (device_ms) *device_ms = total_ms;
  if (sink_rc) return sink_rc;
  return check_device_error(ch);
}

// Closes plogLik of the last sweep and copies the device-resident outputs to the host.
int chain_finish_impl(bsb200_chain *ch, bsb200_cluster_out *out) {
  cudaStream_t s = ch->stream;
  ch->enqueue_reduce_and_param(true, s);
  const int nsnap = ch->nrep / ch->thin + 1, q = ch->q, d = ch->d;
  if (out->plogLik) CK(cudaMemcpyAsync(out->plogLik, ch->plogLik.p, sizeof(double) * ch->nrep, cudaMemcpyDeviceToHost, s));
  if (out->mu) CK(cudaMemcpyAsync(out->mu, ch->snap_mu.p, sizeof(double) * nsnap * q * d, cudaMemcpyDeviceToHost, s));
  if (out->lambda) CK(cudaMemcpyAsync(out->lambda, ch->snap_lambda.p, sizeof(double) * nsnap * ch->nl * d * d, cudaMemcpyDeviceToHost, s));
  if (out->z_mode) {
    if (ch->mode_cnt.p) {   // mode of the kept label snapshots, written for the rank's own spots
      launch_mode_final(ch->mode_cnt.p, ch->mode_first.p, ch->n_own, q, ch->snap_z.p, s);
      CK(cudaMemcpyAsync(out->z_mode + ch->lo, ch->snap_z.p, sizeof(int32_t) * ch->n_own, cudaMemcpyDeviceToHost, s));
    } else {
      std::memset(out->z_mode + ch->lo, 0, sizeof(int32_t) * ch->n_own);   // no row reached mode_from
    }
  }
  CK(cudaStreamSynchronize(s));
  if (ch->comm_rc) return ch->comm_rc;
  return BSB200_OK;
}

}   // namespace bsb

extern "C" {

int bsb200_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}

int64_t bsb200_kernel_launches(void) { return (int64_t) g_launches.load(); }

int bsb200_release_memory(int32_t device) {
  int rc = select_device(device);
  if (rc) return rc;
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemPoolTrimTo(pool, 0));
  int32_t *z = nullptr;
  double *w = nullptr;
  if (g_pinned.take(device, 0, &z, &w)) { cudaFreeHost(z); cudaFreeHost(w); }
  staged_release();
  return BSB200_OK;
}

int bsb200_chain_create(const bsb200_cluster_args *args, bsb200_chain **chain) {
  if (!chain) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  *chain = nullptr;
  return chain_create_impl(args, chain);
}

int bsb200_chain_run(bsb200_chain *chain, int32_t iters, double *device_ms) {
  if (!chain || iters < 0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  return chain_run_impl(chain, iters, nullptr, device_ms, false);
}

int bsb200_chain_profile(bsb200_chain *ch, int32_t iters, double *sweep_ms, int64_t *launches, double *param_ms,
                         double *halo_ms) {
  if (!ch || iters < 0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  CK(cudaSetDevice(ch->device));
  cudaStream_t s = ch->stream;
  const int per_iter = 2 * ch->ord.ncol + 1;
  std::vector<cudaEvent_t> ev((size_t) iters * per_iter * 2);
  for (auto &e : ev) CK(cudaEventCreate(&e));
  size_t k = 0;
  for (int i = 0; i < iters; i++) {
    CK(cudaEventRecord(ev[k++], s));
    ch->enqueue_reduce_and_param(false, s);
    CK(cudaEventRecord(ev[k++], s));
    for (int c = 0; c < ch->ord.ncol; c++) {
      CK(cudaEventRecord(ev[k++], s));
      ch->enqueue_sweep(c, 0, 0, s);
      CK(cudaEventRecord(ev[k++], s));
      CK(cudaEventRecord(ev[k++], s));
      ch->enqueue_halo(c, s);
      CK(cudaEventRecord(ev[k++], s));
    }
    ch->iters_done++;
  }
  CK(cudaStreamSynchronize(s));
  double sw = 0.0, pm = 0.0, hm = 0.0;
  int64_t nl = 0;
  k = 0;
  for (int i = 0; i < iters; i++) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
    pm += ms; k += 2;
    for (int c = 0; c < ch->ord.ncol; c++) {
      CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      sw += ms; k += 2; nl++;
      CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      hm += ms; k += 2;
    }
  }
  for (auto &e : ev) cudaEventDestroy(e);
  if (sweep_ms) *sweep_ms = sw;
  if (param_ms) *param_ms = pm;
  if (halo_ms) *halo_ms = hm;
  if (launches) *launches = nl;
  return check_device_error(ch);
}

int bsb200_chain_state(bsb200_chain *ch, int32_t *z, double *weights, double *mu, double *lambda,
                       double *plogLik, int64_t *iters_done) {
  if (!ch) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  CK(cudaSetDevice(ch->device));
  cudaStream_t s = ch->stream;
  const int q = ch->q, d = ch->d;
  if (z || weights) {   // own spots only; entries of other ranks are left untouched
    launch_snapshot(ch->z.p, ch->w.p, ch->orig.p, ch->n_pos, ch->snap_z.p, ch->snap_w.p, s, ch->lo);
    if (z) CK(cudaMemcpyAsync(z + ch->lo, ch->snap_z.p, sizeof(int32_t) * ch->n_own, cudaMemcpyDeviceToHost, s));
    if (weights) CK(cudaMemcpyAsync(weights + ch->lo, ch->snap_w.p, sizeof(double) * ch->n_own, cudaMemcpyDeviceToHost, s));
  }
  std::vector<double> p((size_t) ch->PL.total);
  CK(cudaMemcpyAsync(p.data(), ch->params.p, sizeof(double) * p.size(), cudaMemcpyDeviceToHost, s));
  if (plogLik && ch->iters_done >= 1) {
    ch->enqueue_reduce_and_param(true, s);   // closes plogLik of the last sweep (idempotent)
    CK(cudaMemcpyAsync(plogLik, ch->plogLik.p, sizeof(double) * std::min<int64_t>(ch->iters_done + 1, ch->nrep),
                       cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  if (mu)
    for (int k = 0; k < q; k++)
      for (int c = 0; c < d; c++) mu[k * d + c] = p[(size_t) ch->PL.mu + k * d + c] + ch->ybar_host[(size_t) c];
  if (lambda) std::memcpy(lambda, &p[(size_t) ch->PL.lam], sizeof(double) * ch->nl * d * d);
  if (iters_done) *iters_done = ch->iters_done;
  return BSB200_OK;
}

int bsb200_chain_dry_sweep(bsb200_chain *ch, int32_t *z_new, double *w_new, double *h_prev, double *h_new,
                           double *prob, int32_t *accepted) {
  if (!ch) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  CK(cudaSetDevice(ch->device));
  cudaStream_t s = ch->stream;
  const size_t n = (size_t) ch->n;   // outputs are indexed by original spot id; own spots are filled
  DevBuf<int32_t> dz, da;
  DevBuf<double> dw, dhp, dhn, dpr;
  int rc;
  if ((rc = dz.alloc(n)) || (rc = da.alloc(n)) || (rc = dw.alloc(n)) || (rc = dhp.alloc(n)) || (rc = dhn.alloc(n)) || (rc = dpr.alloc(n))) return rc;
  // parameters of the next iteration are drawn first (k_param advances the iteration counter), the
  // dry sweep then evaluates that iteration's proposals on the unchanged labels / weights.
  DevBuf<double> params_keep;
  if ((rc = params_keep.alloc((size_t) ch->PL.total))) return rc;
  CK(cudaMemcpyAsync(params_keep.p, ch->params.p, sizeof(double) * ch->PL.total, cudaMemcpyDeviceToDevice, s));
  ch->enqueue_reduce_and_param(false, s);
  for (int c = 0; c < ch->ord.ncol; c++) {
    SweepArgs a = ch->sweep_args(c, SW_DRY, 0);
    a.dry_znew = dz.p; a.dry_w = dw.p; a.dry_hp = dhp.p; a.dry_hn = dhn.p; a.dry_prob = dpr.p; a.dry_acc = da.p;
    if (a.nseg > 0) { launch_sweep_any(ch->K, ch->tdist, ch->vvv, a, a.nseg, s); g_launches++; }
  }
  // roll the chain back: parameters and iteration counter as before (the statistics were not touched)
  CK(cudaMemcpyAsync(ch->params.p, params_keep.p, sizeof(double) * ch->PL.total, cudaMemcpyDeviceToDevice, s));
  const uint32_t it = (uint32_t) ch->iters_done;
  CK(cudaMemcpyAsync(ch->iter.p, &it, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  if (z_new) CK(cudaMemcpyAsync(z_new, dz.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  if (accepted) CK(cudaMemcpyAsync(accepted, da.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  if (w_new) CK(cudaMemcpyAsync(w_new, dw.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (h_prev) CK(cudaMemcpyAsync(h_prev, dhp.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (h_new) CK(cudaMemcpyAsync(h_new, dhn.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (prob) CK(cudaMemcpyAsync(prob, dpr.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return check_device_error(ch);
}

int bsb200_chain_destroy(bsb200_chain *chain) {
  if (chain) {
    cudaSetDevice(chain->device);
    delete chain;
  }
  return BSB200_OK;
}

int bsb200_cluster(const bsb200_cluster_args *args, bsb200_cluster_out *out) {
  if (!args || !out) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  g_cancel.store(0);
  bsb200_chain *ch = nullptr;
  static const bool timing = std::getenv("BSB200_TIMING") != nullptr;
  double t_last = now_ms();
  auto mark = [&](const char *what) {   // BSB200_TIMING=1: wall-clock marks on stderr
    if (!timing) return;
    const double t = now_ms();
    std::fprintf(stderr, "[bsb200 cluster] %-24s %8.1f ms\n", what, t - t_last);
    t_last = t;
  };
  int rc = chain_create_impl(args, &ch);
  if (rc) return rc;
  mark("chain create");
  struct Guard {   // destroys the chain, timed
    bsb200_chain *ch;
    decltype(mark) &mk;
    ~Guard() { delete ch; mk("chain destroy"); }
  } guard{ch, mark};
  const int64_t n = args->n;
  const int d = args->d, q = args->q, nsnap = args->nrep / args->thin + 1, nl = ch->nl;
  // row 0 of every output = initial state (src/cluster.cpp:232-237, 462-469); rows the chain does not reach stay
  // zero as in the reference.  Single GPU: every later row is overwritten in full by its snapshot, so only the rows
  // of an interrupted run are zeroed (at the end) -- a second pass over 1.3 GB of host memory at 10M spots otherwise.
  // Sharded: a rank writes its own columns [own_lo, own_hi) of rows 1.. and nothing else beyond row 0 (include/bsb200.h).
  const bool sharded = ch->world() > 1;
  parallel_for(n, [&](int64_t a, int64_t b, int) {
    if (out->z) std::memcpy(out->z + a, args->init + a, sizeof(int32_t) * (size_t) (b - a));
    if (out->weights && ch->tdist)
      for (int64_t j = a; j < b; j++) out->weights[j] = 1.0;
  });
  if (!sharded && args->thin == 1 && nsnap > 1) {
    // iteration i is stored in row (i + 1) / thin: with thin = 1 the first iteration lands in row 2 and row 1 is never
    // produced (it stays zero in the reference, src/cluster.cpp:310-313) -- do not leave the caller's bytes there
    if (out->z) std::memset(out->z + n, 0, sizeof(int32_t) * n);
    if (out->weights && ch->tdist) std::memset(out->weights + n, 0, sizeof(double) * n);
  }
  if (sink_active() && !sharded) {   // row 0 = initial state
    const int64_t dims[1] = {n};
    if ((rc = sink_write("z", 0, args->init, n, 0, dims, 1))) return rc;
    if (ch->tdist) {
      std::vector<double> ones((size_t) n, 1.0);
      if ((rc = sink_write("weights", 0, ones.data(), n, 1, dims, 1))) return rc;
    }
  }
  out->rows_done = 1;
  out->ncolours = ch->ord.ncol;
  out->setup_ms = ch->setup_ms;
  out->own_lo = ch->lo;
  out->own_hi = ch->hi;
  mark("row 0");
  double ms = 0.0;
  rc = chain_run_impl(ch, args->nrep - 1, out, &ms, true);
  out->device_ms = ms;
  mark("chain run (wall)");
  if (!sharded && out->rows_done < nsnap) {   // interrupted (or failed) run: the rows never reached read as zero
    const size_t r0 = (size_t) out->rows_done, nr = (size_t) (nsnap - out->rows_done);
    if (out->z) std::memset(out->z + r0 * n, 0, sizeof(int32_t) * nr * n);
    if (out->weights && ch->tdist) std::memset(out->weights + r0 * n, 0, sizeof(double) * nr * n);
  }
  if (rc && rc != BSB200_ERR_CANCELLED) return rc;
  int rc2 = chain_finish_impl(ch, out);
  if (rc2) return rc2;
  mark("finish");
  if (out->mu)
    for (int k = 0; k < q; k++)
      for (int c = 0; c < d; c++) out->mu[k * d + c] = args->mu0[c];
  if (out->lambda)
    for (int k = 0; k < nl; k++) std::memcpy(out->lambda + (size_t) k * d * d, args->lambda0, sizeof(double) * d * d);
  if (rc == BSB200_OK) out->rows_done = nsnap;
  if (sink_active() && !sharded && out->mu && out->lambda && out->plogLik) {   // the small parameters, all rows at once
    const int64_t dmu[2] = {q, d}, dlam[3] = {nl, d, d}, one[1] = {1};
    int rc3 = BSB200_OK;
    for (int r = 0; r < out->rows_done && !rc3; r++) {
      rc3 = sink_write("mu", r, out->mu + (size_t) r * q * d, (int64_t) q * d, 1, dmu, 2);
      if (!rc3) rc3 = sink_write("lambda", r, out->lambda + (size_t) r * nl * d * d, (int64_t) nl * d * d, 1, nl > 1 ? dlam : dlam + 1, nl > 1 ? 3 : 2);
    }
    for (int i = 0; i < args->nrep && !rc3; i++) rc3 = sink_write("plogLik", i, out->plogLik + i, 1, 1, one, 1);
    if (rc3) return rc3;
  }
  return rc;
}

}   // extern "C"
