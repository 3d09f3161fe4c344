// probes.cu -- state-level parity probes of the C ABI: each evaluates one quantity of the sampler
// on a caller-supplied state with the same device code the samplers run.
#include <cstring>

#include "chain.hpp"
#include "rng.cuh"
#include "tma_maps.hpp"

namespace bsb {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(BSB200_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

__global__ void k_debug_rng(int kind, Key key, uint32_t purpose, uint32_t index, uint32_t iter, uint32_t slot,
                            double shape, double scale, int count, double *out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (kind == 0) {
    if (t == 0) rng_uniform2(key, purpose, index, iter, slot, out[0], out[1]);
  } else if (kind == 1) {
    if (t < count) out[t] = rng_normal(key, purpose, index, iter, slot + (uint32_t) t);
  } else if (kind == 2) {
    if (t < count) out[t] = rng_gamma(key, purpose, index + (uint32_t) t, iter, shape, scale);
  } else if (kind == 3) {   // raw Philox block: counter (index, slot, iter, purpose)
    if (t == 0) {
      const uint4 o = philox4x32_10(index, slot, iter, purpose, key);
      out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
    }
  }
}

}   // namespace bsb

using namespace bsb;

extern "C" {

int bsb200_debug_rng(int32_t device, int32_t kind, uint64_t seed, uint32_t purpose, uint32_t index,
                     uint32_t iter, uint32_t slot, double shape, double scale, int32_t count, double *out) {
  if (!out || count < 1) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  int rc = select_device(device);
  if (rc) return rc;
  DevBuf<double> d;
  const int need = kind == 0 ? 2 : (kind == 3 ? 4 : count);
  if ((rc = d.alloc((size_t) need))) return rc;
  k_debug_rng<<<(need + 127) / 128, 128>>>(kind, Key{(uint32_t) seed, (uint32_t) (seed >> 32)}, purpose, index, iter,
                                            slot, shape, scale, count, d.p);
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d.p, sizeof(double) * need, cudaMemcpyDeviceToHost));
  return BSB200_OK;
}

int bsb200_eval_loglik(int32_t device, int32_t form, uint32_t compat, const double *Y, int64_t n, int32_t d,
                       int32_t q, const double *mu, const double *mat, int32_t nl, const double *w, double *out) {
  if (!Y || !mu || !mat || !out || n <= 0 || q < 1 || (nl != 1 && nl != q) || (form != 0 && form != 1))
    return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (d < 1 || d > kMaxD) return fail(BSB200_ERR_UNSUPPORTED, "d = %d unsupported (1..%d)", d, kMaxD);
  int rc = select_device(device);
  if (rc) return rc;
  KernelsForD K = kernel_table()[d];
  if (!K.loglik) return fail(BSB200_ERR_UNSUPPORTED, "library was built without kernels for d = %d", d);
  DevBuf<double> dY, dmu, dmat, droot, dld, dw, dout;
  DevBuf<int> derr;
  if ((rc = dY.alloc((size_t) n * d)) || (rc = dmu.alloc((size_t) q * d)) || (rc = dmat.alloc((size_t) nl * d * d)) ||
      (rc = droot.alloc((size_t) nl * d * d)) || (rc = dld.alloc((size_t) nl)) || (rc = dout.alloc((size_t) n * q)) ||
      (rc = derr.alloc(1)))
    return rc;
  CK(cudaMemset(derr.p, 0, sizeof(int)));
  CK(cudaMemcpy(dY.p, Y, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmu.p, mu, sizeof(double) * q * d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmat.p, mat, sizeof(double) * nl * d * d, cudaMemcpyHostToDevice));
  if (w) {
    if ((rc = dw.alloc((size_t) n))) return rc;
    CK(cudaMemcpy(dw.p, w, sizeof(double) * n, cudaMemcpyHostToDevice));
  }
  launch_factor(dmat.p, nl, d, form, droot.p, dld.p, derr.p, 0);
  LoglikArgs a{};
  a.Y = dY.p; a.n = n; a.q = q; a.nl = nl; a.mu = dmu.p; a.rooti = droot.p; a.logdet = dld.p;
  a.w = w ? dw.p : nullptr;
  a.transposed = (form == 0 && !(compat & BSB200_COMPAT_Q1)) ? 1 : 0;
  a.out = dout.p;
  K.loglik(a, 0);
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout.p, sizeof(double) * n * q, cudaMemcpyDeviceToHost));
  int e = 0;
  CK(cudaMemcpy(&e, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return fail(BSB200_ERR_NUMERIC, "matrix not positive definite");
  return BSB200_OK;
}

int bsb200_eval_suffstats(int32_t device, const double *Y, int64_t n, int32_t d, int32_t q, const int32_t *z,
                          const double *w, const double *mu, int32_t nl, double *nk, double *sk, double *scatter) {
  if (!Y || !z || !mu || n <= 0 || q < 1 || q > 255 || (nl != 1 && nl != q)) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (d < 1 || d > kMaxD) return fail(BSB200_ERR_UNSUPPORTED, "d = %d unsupported (1..%d)", d, kMaxD);
  int rc = select_device(device);
  if (rc) return rc;
  KernelsForD K = kernel_table()[d];
  if (!K.sweep) return fail(BSB200_ERR_UNSUPPORTED, "library was built without kernels for d = %d", d);
  for (int64_t j = 0; j < n; j++)
    if (z[j] < 1 || z[j] > q) return fail(BSB200_ERR_INVALID, "z[%lld] outside 1..q", (long long) j);
  // Same staging as a chain: centred Y in identity order, one colour, the chain's segmentation.
  std::vector<int32_t> colour((size_t) n, 0);
  SpotOrder ord;
  ord.build(n, colour.data(), 1, choose_groups(n), choose_segment(n), 0, -1, 0, 0, kOrderAlign);
  const int64_t ld = (ord.npos + 31) / 32 * 32;
  const StatLayout SL = make_stat_layout(q, d, nl);
  const ParamLayout PL = make_param_layout(q, d, nl);
  DevBuf<double> dYraw, dY, dw, dybar, dpart, dgsum, dstats, dmu, dS, dparams, dplog;
  DevBuf<uint8_t> dz;
  DevBuf<int32_t> dorig, dsb, dsc, dss, dgs;
  DevBuf<uint32_t> diter;
  DevBuf<int> derr;
  if ((rc = dYraw.alloc((size_t) n * d)) || (rc = dY.alloc((size_t) d * ld + kSweepPad)) || (rc = dw.alloc((size_t) ld)) ||
      (rc = dybar.alloc((size_t) d)) || (rc = dpart.alloc((size_t) ord.nslots * SL.E)) ||
      (rc = dgsum.alloc((size_t) ord.ngroups * SL.E)) || (rc = dstats.alloc((size_t) SL.E)) ||
      (rc = dmu.alloc((size_t) q * d)) || (rc = dS.alloc((size_t) nl * d * d)) || (rc = dparams.alloc((size_t) PL.total)) ||
      (rc = dplog.alloc(4)) || (rc = dz.alloc((size_t) ld + kSweepPad)) || (rc = dorig.alloc((size_t) ld + kSweepPad)) ||
      (rc = dsb.alloc(ord.seg_begin.size())) || (rc = dsc.alloc(ord.seg_begin.size())) ||
      (rc = dss.alloc(ord.seg_begin.size())) || (rc = dgs.alloc(ord.group_start.size())) || (rc = diter.alloc(1)) ||
      (rc = derr.alloc(1)))
    return rc;
  std::vector<uint8_t> z0((size_t) ld, 0);
  std::vector<double> w0((size_t) ld, 1.0);
  std::vector<int32_t> orig((size_t) ld, -1);
  for (int64_t j = 0; j < n; j++) {
    const size_t pos = (size_t) ord.inv[(size_t) j];
    z0[pos] = (uint8_t) (z[j] - 1);
    if (w) w0[pos] = w[j];
    orig[pos] = (int32_t) j;
  }
  CK(cudaMemset(derr.p, 0, sizeof(int)));
  CK(cudaMemset(diter.p, 0, sizeof(uint32_t)));
  CK(cudaMemset(dY.p, 0, sizeof(double) * ((size_t) d * ld + kSweepPad)));
  CK(cudaMemset(dorig.p, 0, sizeof(int32_t) * ((size_t) ld + kSweepPad)));
  CK(cudaMemcpy(dYraw.p, Y, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dz.p, z0.data(), (size_t) ld, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, w0.data(), sizeof(double) * ld, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dorig.p, orig.data(), sizeof(int32_t) * ld, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsb.p, ord.seg_begin.data(), sizeof(int32_t) * ord.seg_begin.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsc.p, ord.seg_count.data(), sizeof(int32_t) * ord.seg_count.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dss.p, ord.seg_slot.data(), sizeof(int32_t) * ord.seg_slot.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dgs.p, ord.group_start.data(), sizeof(int32_t) * ord.group_start.size(), cudaMemcpyHostToDevice));
  launch_colmeans(dYraw.p, n, d, dybar.p, 0);
  launch_permute_center(dYraw.p, n, ord.npos, d, dorig.p, dybar.p, dY.p, ld, 0);
  std::vector<double> ybar((size_t) d), muc((size_t) q * d);
  CK(cudaMemcpy(ybar.data(), dybar.p, sizeof(double) * d, cudaMemcpyDeviceToHost));
  for (int k = 0; k < q; k++)
    for (int c = 0; c < d; c++) muc[(size_t) k * d + c] = mu[k * d + c] - ybar[(size_t) c];
  CK(cudaMemcpy(dmu.p, muc.data(), sizeof(double) * q * d, cudaMemcpyHostToDevice));
  SweepMaps maps;
  {
    const char *msg = "";
    if (encode_sweep_maps(dY.p, ld, d, nullptr, 0, dorig.p, &maps, &msg))
      return fail(BSB200_ERR_CUDA, "cannot encode the TMA descriptors (%s)", msg);
  }
  SweepArgs a{};
  a.maps = &maps;
  a.Y = dY.p; a.ld = ld; a.z = dz.p; a.w = dw.p; a.ell = nullptr; a.maxdeg = 0; a.orig = dorig.p;
  a.seg_begin = dsb.p; a.seg_count = dsc.p; a.seg_slot = dss.p; a.nseg = (int) ord.seg_begin.size();
  a.params = dparams.p; a.PL = PL; a.partials = dpart.p; a.SL = SL; a.iter = diter.p; a.flags = SW_STATS_ONLY;
  a.err = derr.p;
  launch_sweep_any(K, 1, nl > 1 ? 1 : 0, a, a.nseg, 0);
  g_launches++;
  ParamArgs p{};
  p.PL = PL; p.SL = SL; p.params = dparams.p; p.partials = dpart.p; p.group_start = dgs.p; p.ngroups = ord.ngroups;
  p.gsum = dgsum.p; p.gsum_ready = 0; p.stats = dstats.p; p.iter = diter.p; p.nrep = 1; p.thin = 1; p.plogLik = dplog.p;
  p.finalize_only = 1; p.err = derr.p; p.n = n;
  param_kernel_init();
  launch_param(p, 0);
  g_launches++;
  launch_scatter_probe(dstats.p, SL, dmu.p, nl, dS.p, 0);
  CK(cudaGetLastError());
  std::vector<double> st((size_t) SL.E);
  CK(cudaMemcpy(st.data(), dstats.p, sizeof(double) * SL.E, cudaMemcpyDeviceToHost));
  if (scatter) CK(cudaMemcpy(scatter, dS.p, sizeof(double) * nl * d * d, cudaMemcpyDeviceToHost));
  for (int k = 0; k < q; k++) {
    if (nk) nk[k] = st[(size_t) SL.nk + k];
    if (sk)
      for (int c = 0; c < d; c++) sk[k * d + c] = st[(size_t) SL.sk + k * d + c] + st[(size_t) SL.nk + k] * ybar[(size_t) c];
  }
  return BSB200_OK;
}

}   // extern "C"
