// deconv.cu -- host layer of bsb200_deconv, the replacement of iterate_deconv
// (/root/reference/src/cluster.cpp:742-1141 behind src/RcppExports.cpp:107-138).
//
// Schedule per iteration: k_param (mu, Lambda; plogLik / Ychange of the previous sweep), then one
// k_deconv launch per PARENT colour class.  Parents of one class are never adjacent, so no subspot
// handled in a launch reads a label that the same launch writes, except inside its own parent,
// where the S updates run sequentially as in the reference.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "chain.hpp"
#include "deconv.hpp"

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

template <typename T>
static int upload(DevBuf<T> &buf, const T *host, size_t count, cudaStream_t s) {
  int rc = buf.alloc(count);
  if (rc) return rc;
  if (count) CK(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  return BSB200_OK;
}

struct DeconvRun {
  cudaStream_t stream = nullptr;
  cudaGraphExec_t graph_iter = nullptr, graph_batch = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ~DeconvRun() {
    if (graph_iter) cudaGraphExecDestroy(graph_iter);
    if (graph_batch) cudaGraphExecDestroy(graph_batch);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

static int deconv_impl(const bsb200_deconv_args *a, bsb200_deconv_out *out) {
  const double t0 = now_ms();
  if (!a || !out) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (!a->subspot_positions || !a->spot_nbr_ptr || !a->Y || !a->init || !a->mu0 || !a->lambda0)
    return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  const int64_t n = a->n, n0 = a->n0;
  const int d = a->d, q = a->q, S = a->subspots;
  if (n <= 0 || n0 <= 0 || n > 0x7fffff00ll) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (d < 1 || d > kMaxD) return fail(BSB200_ERR_UNSUPPORTED, "d = %d unsupported (1..%d)", d, kMaxD);
  if (q < 2 || q > 255) return fail(BSB200_ERR_UNSUPPORTED, "q = %d unsupported (2..255)", q);
  if (a->nrep < 1 || a->thin < 1) return fail(BSB200_ERR_INVALID, "nrep and thin must be positive");
  if (S < 1 || S > 32) return fail(BSB200_ERR_UNSUPPORTED, "subspots = %d unsupported (1..32)", S);
  // find_subspot_neighbors first: it is where the reference raises "Invalid arguments!" (:168-174)
  std::vector<std::vector<int32_t>> sub_adj;
  int rc = subspot_neighbors(a->subspot_positions, n, n0, a->spot_nbr_ptr, a->spot_nbr_idx, a->dist, sub_adj);
  if (rc) return rc;
  if ((int64_t) S * n0 != n) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  for (int64_t s = 0; s < n; s++)
    if (a->init[s] < 1 || a->init[s] > q) return fail(BSB200_ERR_INVALID, "init[%lld] = %d outside 1..q", (long long) s, a->init[s]);
  int maxdeg = 0;
  for (const auto &v : sub_adj) maxdeg = std::max(maxdeg, (int) v.size());
  if (maxdeg > kMaxDeg) return fail(BSB200_ERR_UNSUPPORTED, "a subspot has %d neighbours (max %d)", maxdeg, kMaxDeg);
  if (out->df_j) {   // convert_neighbors2mtx (:147-159)
    if (maxdeg > 4) return fail(BSB200_ERR_UNSUPPORTED, "convert_neighbors2mtx: more than 4 neighbours");
    std::memset(out->df_j, 0, sizeof(int64_t) * 4 * (size_t) n);
    for (int64_t s = 0; s < n; s++)
      for (size_t e = 0; e < sub_adj[(size_t) s].size(); e++) out->df_j[e * (size_t) n + s] = (int64_t) sub_adj[(size_t) s][e] + 1;
  }
  rc = select_device(a->device);
  if (rc) return rc;
  KernelsForD K = kernel_table()[d];
  if (!K.deconv) return fail(BSB200_ERR_UNSUPPORTED, "library was built without kernels for d = %d", d);
  g_cancel.store(0);

  // ---- parent colouring and order ----------------------------------------------------------------
  std::vector<int32_t> colour;
  int32_t ncol = 0;
  rc = colour_graph(a->spot_nbr_ptr, a->spot_nbr_idx, n0, colour, &ncol);
  if (rc) return rc;
  const int gpw = 32 / S, ppb = (kBlock / 32) * gpw;
  SpotOrder ord;
  ord.build(n0, colour.data(), ncol, choose_groups(n), std::max(1, choose_segment(n) / kBlock) * ppb);
  const int64_t ld0 = (n0 + 31) / 32 * 32;
  const double js = a->jitter_scale / d;   // :814
  const bool adaptive = js == 0.0;
  const double err_sd = adaptive ? 1.0 : std::sqrt(js);
  const int nsnap = a->nrep / a->thin + 1;
  const int np = d * (d + 1) / 2;

  DeconvRun run;
  CK(cudaStreamCreateWithFlags(&run.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&run.ev0));
  CK(cudaEventCreate(&run.ev1));
  cudaStream_t s = run.stream;
  DevBuf<double> Yraw, Y, Y0, w, ybar, params, partials, gsum, stats, mu0c, lambda0, lm0, plogLik, ychange, snap_mu,
      snap_lambda, snap_w, snap_Y, Ymean, A;
  DevBuf<uint8_t> z;
  DevBuf<int32_t> ell, orig0, seg_begin, seg_count, seg_slot, group_start, snap_z, nacc;
  DevBuf<uint32_t> iter;
  DevBuf<int> err;

  if ((rc = upload(Yraw, a->Y, (size_t) n * d, s))) return rc;
  if ((rc = ybar.alloc((size_t) d)) || (rc = Y.alloc((size_t) d * S * ld0)) || (rc = Y0.alloc((size_t) d * ld0))) return rc;
  {
    std::vector<int32_t> o0((size_t) ld0, 0);
    std::copy(ord.perm.begin(), ord.perm.end(), o0.begin());
    if ((rc = upload(orig0, o0.data(), o0.size(), s))) return rc;
    std::vector<uint8_t> z0((size_t) S * ld0, 0);
    std::vector<int32_t> ellh((size_t) std::max(maxdeg, 1) * S * ld0, -1);
    for (int r = 0; r < S; r++)
      for (int64_t ip = 0; ip < n0; ip++) {
        const int64_t so = (int64_t) r * n0 + ord.perm[(size_t) ip];
        z0[(size_t) r * ld0 + ip] = (uint8_t) (a->init[so] - 1);
        const auto &nb = sub_adj[(size_t) so];
        for (size_t e = 0; e < nb.size(); e++) {
          const int64_t rr = nb[e] / n0, jj = nb[e] % n0;
          ellh[((size_t) e * S + r) * ld0 + ip] = (int32_t) (rr * ld0 + ord.inv[(size_t) jj]);
        }
      }
    if ((rc = upload(z, z0.data(), z0.size(), s))) return rc;
    if ((rc = upload(ell, ellh.data(), ellh.size(), s))) return rc;
    std::vector<double> w1((size_t) S * ld0, 1.0);
    if ((rc = upload(w, w1.data(), w1.size(), s))) return rc;
    CK(cudaStreamSynchronize(s));
  }
  CK(cudaMemsetAsync(Y.p, 0, sizeof(double) * d * S * ld0, s));
  CK(cudaMemsetAsync(Y0.p, 0, sizeof(double) * d * ld0, s));
  launch_colmeans(Yraw.p, n, d, ybar.p, s);
  for (int r = 0; r < S; r++)
    launch_permute_center(Yraw.p + (size_t) r * n0, n, n0, d, orig0.p, ybar.p, Y.p + (size_t) r * ld0, (int64_t) S * ld0, s);
  launch_permute_center(Yraw.p, n, n0, d, orig0.p, ybar.p, Y0.p, ld0, s);   // Y0 = Y.rows(0, n0-1) (:777)
  std::vector<double> ybar_h((size_t) d);
  CK(cudaMemcpyAsync(ybar_h.data(), ybar.p, sizeof(double) * d, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));

  const ParamLayout PL = make_param_layout(q, d, 1);
  const StatLayout SL = make_stat_layout(q, d, 1);
  {
    std::vector<double> m0((size_t) d), l0((size_t) d, 0.0), p0((size_t) PL.total, 0.0);
    for (int c = 0; c < d; c++) m0[(size_t) c] = a->mu0[c] - ybar_h[(size_t) c];
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) l0[(size_t) i] += a->lambda0[(size_t) j * d + i] * m0[(size_t) j];
    std::memcpy(&p0[(size_t) PL.lam], a->lambda0, sizeof(double) * d * d);
    if ((rc = upload(mu0c, m0.data(), (size_t) d, s)) || (rc = upload(lm0, l0.data(), (size_t) d, s)) ||
        (rc = upload(lambda0, a->lambda0, (size_t) d * d, s)) || (rc = upload(params, p0.data(), p0.size(), s)))
      return rc;
    CK(cudaStreamSynchronize(s));
  }
  if ((rc = upload(seg_begin, ord.seg_begin.data(), ord.seg_begin.size(), s)) ||
      (rc = upload(seg_count, ord.seg_count.data(), ord.seg_count.size(), s)) ||
      (rc = upload(seg_slot, ord.seg_slot.data(), ord.seg_slot.size(), s)) ||
      (rc = upload(group_start, ord.group_start.data(), ord.group_start.size(), s)))
    return rc;
  if ((rc = partials.alloc((size_t) ord.nslots * SL.E)) || (rc = gsum.alloc((size_t) ord.ngroups * SL.E)) ||
      (rc = stats.alloc((size_t) SL.E)) || (rc = plogLik.alloc((size_t) a->nrep)) || (rc = ychange.alloc((size_t) a->nrep)) ||
      (rc = snap_mu.alloc((size_t) nsnap * q * d)) || (rc = snap_lambda.alloc((size_t) nsnap * d * d)) ||
      (rc = snap_z.alloc((size_t) n)) || (rc = snap_w.alloc((size_t) n)) || (rc = snap_Y.alloc((size_t) n * d)) ||
      (rc = iter.alloc(1)) || (rc = err.alloc(1)) || (rc = nacc.alloc((size_t) ld0)))
    return rc;
  if (out->Y_mean && (rc = Ymean.alloc((size_t) n * d))) return rc;
  DevBuf<uint32_t> mode_cnt, mode_first;   // q x n label histogram of the kept snapshots
  if (out->z_mode && ((rc = mode_cnt.alloc((size_t) q * n)) || (rc = mode_first.alloc((size_t) q * n)))) return rc;
  if (adaptive) {
    if ((rc = A.alloc((size_t) np * S * ld0))) return rc;
    std::vector<double> Ah((size_t) np * S * ld0, 0.0);
    for (int i = 0; i < d; i++) {
      double *row = &Ah[(size_t) (i * (i + 1) / 2 + i) * S * ld0];
      std::fill(row, row + (size_t) S * ld0, 1.0);
    }
    CK(cudaMemcpyAsync(A.p, Ah.data(), sizeof(double) * Ah.size(), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  {
    std::vector<double> nanv((size_t) a->nrep, std::numeric_limits<double>::quiet_NaN());
    CK(cudaMemcpyAsync(plogLik.p, nanv.data(), sizeof(double) * a->nrep, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ychange.p, nanv.data(), sizeof(double) * a->nrep, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  CK(cudaMemsetAsync(iter.p, 0, sizeof(uint32_t), s));
  CK(cudaMemsetAsync(err.p, 0, sizeof(int), s));
  CK(cudaMemsetAsync(nacc.p, 0, sizeof(int32_t) * ld0, s));
  CK(cudaMemsetAsync(snap_mu.p, 0, sizeof(double) * nsnap * q * d, s));
  CK(cudaMemsetAsync(snap_lambda.p, 0, sizeof(double) * nsnap * d * d, s));
  if (out->Y_mean) CK(cudaMemsetAsync(Ymean.p, 0, sizeof(double) * n * d, s));
  if (out->z_mode) CK(cudaMemsetAsync(mode_cnt.p, 0, sizeof(uint32_t) * q * n, s));

  // one launch for all colour classes needs every CTA resident at once (CTAs of a later class spin on the earlier ones)
  DevBuf<int32_t> colour_seg_start;
  DevBuf<uint32_t> colour_done;
  if ((rc = upload(colour_seg_start, ord.colour_seg_start.data(), ord.colour_seg_start.size(), s)) ||
      (rc = colour_done.alloc((size_t) std::max<int>(ncol, 1))))
    return rc;
  CK(cudaMemsetAsync(colour_done.p, 0, sizeof(uint32_t) * std::max<int>(ncol, 1), s));
  bool fuse_colours = false;
  {
    int sms = 0, dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static const bool off = std::getenv("BSB200_NO_FUSED_DECONV") != nullptr;
    // conservative: one CTA per SM is always resident whatever the kernel's registers / shared memory
    fuse_colours = !off && ncol > 1 && (int) ord.seg_begin.size() <= 2 * sms;
  }
  auto deconv_args = [&](int colour_id, int stats_only) {
    DeconvArgs k{};
    k.Y = Y.p; k.Y0 = Y0.p; k.z = z.p; k.w = w.p; k.ell = ell.p; k.maxdeg = maxdeg; k.orig0 = orig0.p;
    k.A = A.p; k.nacc = nacc.p; k.ld0 = ld0; k.n0 = n0; k.S = S;
    const int s0 = colour_id < 0 ? 0 : ord.colour_seg_start[(size_t) colour_id];
    const int s1 = colour_id < 0 ? (int) ord.seg_begin.size() : ord.colour_seg_start[(size_t) colour_id + 1];
    k.seg_begin = seg_begin.p + s0; k.seg_count = seg_count.p + s0; k.seg_slot = seg_slot.p + s0; k.nseg = s1 - s0;
    k.params = params.p; k.PL = PL; k.partials = partials.p; k.SL = SL; k.iter = iter.p;
    k.key0 = (uint32_t) a->seed; k.key1 = (uint32_t) (a->seed >> 32);
    k.gamma = a->gamma; k.c = a->c; k.err_sd = err_sd; k.tdist = a->tdist ? 1 : 0; k.adaptive = adaptive ? 1 : 0;
    k.adapt_before = a->adapt_before; k.stats_only = stats_only; k.err = err.p;
    k.fuse = (fuse_colours && colour_id < 0 && !stats_only) ? 1 : 0; k.ncol = ncol;
    k.colour_seg_start = colour_seg_start.p; k.done = colour_done.p;
    return k;
  };
  auto param_args = [&](int finalize_only) {
    ParamArgs p{};
    p.PL = PL; p.SL = SL; p.params = params.p; p.partials = partials.p; p.group_start = group_start.p;
    p.ngroups = ord.ngroups; p.gsum = gsum.p; p.gsum_ready = ord.ngroups > 1; p.stats = stats.p; p.T_fixed = nullptr;
    p.mu0c = mu0c.p; p.lambda0 = lambda0.p; p.lm0 = lm0.p; p.ybar = ybar.p; p.alpha = a->alpha; p.beta = a->beta;
    p.n = n; p.tdist = 1; p.vvv = 0; p.compat = 0; p.key0 = (uint32_t) a->seed; p.key1 = (uint32_t) (a->seed >> 32);
    p.iter = iter.p; p.nrep = a->nrep; p.thin = a->thin; p.plogLik = plogLik.p; p.ychange = ychange.p; p.n0 = (double) n0;
    p.snap_mu = snap_mu.p; p.snap_lambda = snap_lambda.p; p.finalize_only = finalize_only; p.precision_form = 1;
    p.err = err.p;
    return p;
  };
  auto enqueue_param = [&](int finalize_only) {
    if (ord.ngroups > 1) launch_reduce_groups(partials.p, group_start.p, ord.ngroups, SL.E, gsum.p, s);
    launch_param(param_args(finalize_only), s);
    g_launches++;
  };
  auto enqueue_iteration = [&]() {
    enqueue_param(0);
    if (fuse_colours) {   // every colour class in one launch, ordered by in-kernel counters (deconv.hpp)
      DeconvArgs k = deconv_args(-1, 0);
      K.deconv(k, k.nseg, s);
      g_launches++;
      return;
    }
    for (int c = 0; c < ncol; c++) {
      DeconvArgs k = deconv_args(c, 0);
      if (k.nseg > 0) { K.deconv(k, k.nseg, s); g_launches++; }
    }
  };
  int kernels_per_iter = 0;
  param_kernel_init();
  {   // the configuration must fit the shared memory of one CTA: say so instead of failing inside a launch
    const size_t limit = device_smem_limit(), need = K.deconv_smem(deconv_args(-1, 1)), need_param = param_smem_min(q, d);
    if (limit && (need > limit || need_param > std::min<size_t>(limit, 220 * 1024)))
      return fail(BSB200_ERR_UNSUPPORTED,
                  "d = %d, q = %d needs %zu (deconvolution sweep) / %zu (parameter update) bytes of shared memory per CTA, "
                  "the device offers %zu: reduce q or d", d, q, need, need_param, limit);
  }
  {   // statistics of the initial state
    DeconvArgs k = deconv_args(-1, 1);
    K.deconv(k, k.nseg, s);
    g_launches++;
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  {
    cudaGraph_t g = nullptr;
    const long long launches_before = g_launches.load();
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    enqueue_iteration();
    CK(cudaStreamEndCapture(s, &g));
    kernels_per_iter = (int) (g_launches.load() - launches_before);
    CK(cudaGraphInstantiate(&run.graph_iter, g, 0));
    cudaGraphDestroy(g);
    if (a->thin > 1 && a->thin <= 256 && n <= 400000) {
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < a->thin; i++) enqueue_iteration();
      CK(cudaStreamEndCapture(s, &g));
      CK(cudaGraphInstantiate(&run.graph_batch, g, 0));
      cudaGraphDestroy(g);
    }
    g_launches.store(launches_before);
  }
  // ---- chain persistence: rows go to the registered sink as they are produced (staged here when the caller did not
  //      ask for the matrices themselves)
  const bool sink = sink_active();
  std::vector<int32_t> sink_z;
  std::vector<double> sink_w, sink_Y, sink_Yt;
  if (sink) {
    if (!out->z) sink_z.resize((size_t) n);
    if (!out->weights) sink_w.resize((size_t) n);
    if (!out->Y) sink_Y.resize((size_t) n * d);
    sink_Yt.resize((size_t) n * d);
  }
  auto sink_row = [&](size_t r, const int32_t *zr, const double *wr, const double *Yr) -> int {
    const int64_t dn[1] = {n}, dY[2] = {n, d};
    int rc2 = sink_write("z", (int64_t) r, zr, n, 0, dn, 1);
    if (!rc2) rc2 = sink_write("weights", (int64_t) r, wr, n, 1, dn, 1);
    if (!rc2) {   // as.vector(t(Y_r)): spot-major rows of d values (R/mcmcChain.R:206-211)
      for (int64_t j = 0; j < n; j++)
        for (int c = 0; c < d; c++) sink_Yt[(size_t) j * d + c] = Yr[(size_t) c * n + j];
      rc2 = sink_write("Y", (int64_t) r, sink_Yt.data(), n * d, 1, dY, 2);
    }
    return rc2;
  };
  if (sink) {
    std::vector<double> ones((size_t) n, 1.0);
    if ((rc = sink_row(0, a->init, ones.data(), a->Y))) return rc;
  }
  // ---- row 0 of the outputs (:793-801) -------------------------------------------------------------
  if (out->z) {
    std::memset(out->z, 0, sizeof(int32_t) * (size_t) nsnap * n);
    std::memcpy(out->z, a->init, sizeof(int32_t) * n);
  }
  if (out->weights) {
    std::memset(out->weights, 0, sizeof(double) * (size_t) nsnap * n);
    for (int64_t j = 0; j < n; j++) out->weights[j] = 1.0;
  }
  if (out->Y) {
    std::memset(out->Y, 0, sizeof(double) * (size_t) nsnap * n * d);
    std::memcpy(out->Y, a->Y, sizeof(double) * (size_t) n * d);
  }
  out->rows_done = 1;
  out->ncolours = ncol;
  out->setup_ms = now_ms() - t0;

  // ---- sampling loop ---------------------------------------------------------------------------------
  double total_ms = 0.0;
  int64_t done = 0, remaining = a->nrep - 1;
  int status = BSB200_OK, mean_count = 0;
  while (remaining > 0) {
    const int64_t next_multiple = ((done + 2 + a->thin - 1) / a->thin) * a->thin;
    const int64_t runit = std::min<int64_t>(next_multiple - 1 - done, remaining);
    CK(cudaEventRecord(run.ev0, s));
    if (runit == a->thin && run.graph_batch) CK(cudaGraphLaunch(run.graph_batch, s));
    else for (int64_t i = 0; i < runit; i++) CK(cudaGraphLaunch(run.graph_iter, s));
    g_launches += (long long) kernels_per_iter * runit;
    done += runit;
    remaining -= runit;
    const bool boundary = (done + 1) % a->thin == 0;
    const size_t r = (size_t) ((done + 1) / a->thin);
    if (boundary) {
      for (int rr = 0; rr < S; rr++) {
        launch_snapshot(z.p + (size_t) rr * ld0, w.p + (size_t) rr * ld0, orig0.p, n0, snap_z.p + (size_t) rr * n0,
                        snap_w.p + (size_t) rr * n0, s);
        if (out->Y || out->Y_mean || sink)
          launch_snapshot_Y(Y.p + (size_t) rr * ld0, (int64_t) S * ld0, n0, d, orig0.p, ybar.p, snap_Y.p + (size_t) rr * n0, n, s);
      }
      if (out->Y_mean && (int) r >= out->mean_from) {
        mean_count++;
        launch_running_mean(snap_Y.p, Ymean.p, n * d, 1.0 / mean_count, s);
      }
      if (out->z_mode && (int) r >= out->mode_from && r < (size_t) nsnap)
        launch_mode_accumulate(snap_z.p, n, q, (uint32_t) r, mode_cnt.p, mode_first.p, s);
      if (out->z) CK(cudaMemcpyAsync(out->z + r * n, snap_z.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
      else if (sink) CK(cudaMemcpyAsync(sink_z.data(), snap_z.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
      if (out->weights) CK(cudaMemcpyAsync(out->weights + r * n, snap_w.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
      else if (sink) CK(cudaMemcpyAsync(sink_w.data(), snap_w.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
      if (out->Y) CK(cudaMemcpyAsync(out->Y + r * (size_t) n * d, snap_Y.p, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s));
      else if (sink) CK(cudaMemcpyAsync(sink_Y.data(), snap_Y.p, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(run.ev1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, run.ev0, run.ev1));
    total_ms += ms;
    if (boundary) {
      if (sink && r < (size_t) nsnap) {
        const int rc2 = sink_row(r, out->z ? out->z + r * n : sink_z.data(), out->weights ? out->weights + r * n : sink_w.data(),
                                 out->Y ? out->Y + r * (size_t) n * d : sink_Y.data());
        if (rc2) return rc2;
      }
      out->rows_done = (int32_t) r + 1;
      int e = 0;
      CK(cudaMemcpy(&e, err.p, sizeof(int), cudaMemcpyDeviceToHost));
      if (e) return fail(BSB200_ERR_NUMERIC, "NA in probability vector, Wishart df <= 0 or matrix not positive definite");
    }
    // early stop (:1106-1107) / user interrupt (:867-868), polled after every batch of at most `thin` iterations
    if (cancel_requested()) { status = fail(BSB200_ERR_CANCELLED, "Stopping..."); break; }
  }
  {
    int e = 0;
    CK(cudaMemcpy(&e, err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return fail(BSB200_ERR_NUMERIC, "NA in probability vector, Wishart df <= 0 or matrix not positive definite");
  }
  enqueue_param(1);   // closes plogLik / Ychange of the last sweep
  out->device_ms = total_ms;
  if (out->plogLik) CK(cudaMemcpyAsync(out->plogLik, plogLik.p, sizeof(double) * a->nrep, cudaMemcpyDeviceToHost, s));
  if (out->Ychange) CK(cudaMemcpyAsync(out->Ychange, ychange.p, sizeof(double) * a->nrep, cudaMemcpyDeviceToHost, s));
  if (out->mu) CK(cudaMemcpyAsync(out->mu, snap_mu.p, sizeof(double) * nsnap * q * d, cudaMemcpyDeviceToHost, s));
  if (out->lambda) CK(cudaMemcpyAsync(out->lambda, snap_lambda.p, sizeof(double) * nsnap * d * d, cudaMemcpyDeviceToHost, s));
  if (out->Y_mean) CK(cudaMemcpyAsync(out->Y_mean, Ymean.p, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s));
  if (out->z_mode) {
    launch_mode_final(mode_cnt.p, mode_first.p, n, q, snap_z.p, s);
    CK(cudaMemcpyAsync(out->z_mode, snap_z.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  if (out->mu)
    for (int k = 0; k < q; k++)
      for (int c = 0; c < d; c++) out->mu[k * d + c] = a->mu0[c];
  if (out->lambda) std::memcpy(out->lambda, a->lambda0, sizeof(double) * d * d);
  if (sink && out->mu && out->lambda && out->plogLik && out->Ychange) {   // the small parameters, all rows at once
    const int64_t dmu[2] = {q, d}, dlam[2] = {d, d}, one[1] = {1};
    int rc3 = BSB200_OK;
    for (int r = 0; r < out->rows_done && !rc3; r++) {
      rc3 = sink_write("mu", r, out->mu + (size_t) r * q * d, (int64_t) q * d, 1, dmu, 2);
      if (!rc3) rc3 = sink_write("lambda", r, out->lambda + (size_t) r * d * d, (int64_t) d * d, 1, dlam, 2);
    }
    for (int i = 0; i < a->nrep && !rc3; i++) {
      rc3 = sink_write("plogLik", i, out->plogLik + i, 1, 1, one, 1);
      if (!rc3) rc3 = sink_write("Ychange", i, out->Ychange + i, 1, 1, one, 1);
    }
    if (rc3) return rc3;
  }
  return status;
}

}   // namespace bsb

extern "C" int bsb200_deconv(const bsb200_deconv_args *args, bsb200_deconv_out *out) {
  return bsb::deconv_impl(args, out);
}
