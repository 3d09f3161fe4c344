// util_kernels.cu -- staging / bookkeeping kernels shared by the samplers (dimension-agnostic).
#include "util_kernels.hpp"

#include "smallmat.cuh"

namespace bsb {

std::atomic<long long> g_launches{0};

// Column means of an n x d column-major matrix in a fixed (grid-independent) summation order:
// one CTA per column, 256 strided partial sums, then a fixed tree.
__global__ void __launch_bounds__(256) k_colmeans(const double *Y, int64_t n, double *ybar) {
  __shared__ double red[256];
  const double *col = Y + (size_t) blockIdx.x * n;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += 256) s += col[j];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ybar[blockIdx.x] = red[0] / (double) n;
}

// Yp[c][i] = Y[c][orig[i]] - ybar[c]   (gather into internal order + centring)
__global__ void k_permute_center(const double *Y, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                                 const double *ybar, double *Yp, int64_t dst_ld, int64_t orig_off) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (orig[i] < 0) {   // hole of the internal order
    for (int c = 0; c < d; c++) Yp[(size_t) c * dst_ld + i] = 0.0;
    return;
  }
  const int64_t o = orig[i] - orig_off;
  for (int c = 0; c < d; c++) Yp[(size_t) c * dst_ld + i] = Y[(size_t) c * src_ld + o] - ybar[c];
}

// Snapshot of the labels / weights in the caller's spot order (src/cluster.cpp:313, :559-560).
__global__ void k_snapshot(const uint8_t *z, const double *w, const int32_t *orig, int64_t n,
                           int32_t *z_out, double *w_out, int64_t orig_off) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || orig[i] < 0) return;
  const int64_t o = orig[i] - orig_off;
  z_out[o] = (int32_t) z[i] + 1;
  if (w_out) w_out[o] = w[i];
}

// Y snapshot for the deconvolution sampler: out[c][orig[i]] = Yp[c][i] + ybar[c]
__global__ void k_snapshot_Y(const double *Yp, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                             const double *ybar, double *out, int64_t dst_ld) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t o = orig[i];
  for (int c = 0; c < d; c++) out[(size_t) c * dst_ld + o] = Yp[(size_t) c * src_ld + i] + ybar[c];
}

// Per-spot label histogram over the kept snapshots, for the mode the R callers compute with apply(zs, 2, Mode)
// (R/spatialCluster.R:204-213, R/spatialEnhance.R:450-458).  z1: one snapshot, 1-based labels in the caller's
// order.  cnt / first: q x n; `first` keeps the row in which a label first appeared (Mode()'s tie rule).
__global__ void k_mode_accumulate(const int32_t *z1, int64_t n, int q, uint32_t row, uint32_t *cnt, uint32_t *first) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = z1[i] - 1;
  if ((unsigned) k >= (unsigned) q) return;
  const size_t e = (size_t) k * n + i;
  const uint32_t c = cnt[e];
  if (c == 0u) first[e] = row;
  cnt[e] = c + 1u;
}

// R/utils.R:63-66  Mode(x) = ux[which.max(tabulate(match(x, ux)))], ux = unique(x): the most frequent label,
// ties broken by the earliest first appearance.  0 where no snapshot was kept.
__global__ void k_mode_final(const uint32_t *cnt, const uint32_t *first, int64_t n, int q, int32_t *out) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int best = -1;
  uint32_t bc = 0u, bf = 0u;
  for (int k = 0; k < q; k++) {
    const uint32_t c = cnt[(size_t) k * n + i];
    if (c == 0u) continue;
    const uint32_t f = first[(size_t) k * n + i];
    if (c > bc || (c == bc && f < bf)) { best = k; bc = c; bf = f; }
  }
  out[i] = best + 1;
}

// running mean of Y snapshots: mean += (Y - mean) / count   (R/spatialEnhance.R:441)
__global__ void k_running_mean(const double *snap, double *mean, int64_t len, double inv_count) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) mean[i] += (snap[i] - mean[i]) * inv_count;
}

// Level 1 of the two-level statistics reduction when there are many segments: one thread per
// (group, entry), sequential over the group's slots (fixed order).
__global__ void __launch_bounds__(256) k_reduce_groups(const double *partials, const int32_t *group_start, int E,
                                                      double *gsum, const uint32_t *gepoch, size_t par_stride) {
  // CTA = 32 entries x 8 sub-ranges of the group's slots.  Each thread folds its sub-range with four interleaved
  // running sums, the 8 sub-range sums are then folded in order: a fixed association that depends only on the
  // group's slot count (a global quantity), never on the grid or the GPU count.
  __shared__ double sub[8][33];
  const int le = threadIdx.x & 31, sr = threadIdx.x >> 5, e = blockIdx.x * 32 + le, g = blockIdx.y;
  if (gepoch) gsum += (size_t) ((*gepoch + 1u) & 1u) * par_stride;   // the parity the coming gather publishes
  const int begin = group_start[g], end = group_start[g + 1];
  const int len = (end - begin + 7) / 8;
  int slot = begin + sr * len;
  const int stop = min(end, slot + len);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (e < E) {
    const double *p = partials + (size_t) slot * E + e;
    for (; slot + 3 < stop; slot += 4, p += 4 * (size_t) E) {
      a0 += __ldcg(p);
      a1 += __ldcg(p + E);
      a2 += __ldcg(p + 2 * (size_t) E);
      a3 += __ldcg(p + 3 * (size_t) E);
    }
    for (; slot < stop; slot++, p += E) a0 += __ldcg(p);
  }
  sub[sr][le] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (sr == 0 && e < E) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) s += sub[r][le];
    gsum[(size_t) g * E + e] = s;
  }
}

// rooti / logdet per matrix for the log-density probe (one warp per matrix):
// form 0: mat = Sigma  -> rooti = inv(chol(Sigma))   (src/cluster.cpp:91-92)
// form 1: mat = Lambda -> rooti = chol(Lambda)       (:117-118)
__global__ void k_factor(const double *mat, int d, int form, double *rooti, double *logdet, int *err) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x, k = blockIdx.x;
  double *A = sm, *R = A + d * d, *T = R + d * d;
  for (int e = lane; e < d * d; e += 32) A[e] = mat[(size_t) k * d * d + e];
  __syncwarp();
  bool ok = warp_chol_upper(A, R, d, lane);
  const double *out = R;
  if (form == 0) {
    warp_inv_upper(R, T, d, lane);
    out = T;
  }
  for (int e = lane; e < d * d; e += 32) rooti[(size_t) k * d * d + e] = out[e];
  if (lane == 0) {
    double ld = 0.0;
    for (int i = 0; i < d; i++) ld += log(out[(size_t) i * d + i]);
    logdet[k] = ld;
    if (!ok) atomicExch(err, BSB_DEVERR_NUMERIC);
  }
}

// Scatter about the cluster means from the moment sums (the identity k_param uses), for the
// sufficient-statistics probe: S = T - sum_k (mu_k s_k^T + s_k mu_k^T - n_k mu_k mu_k^T).
__global__ void k_scatter_probe(const double *stats, StatLayout SL, const double *mu, int nl, double *S) {
  const int d = SL.d, q = SL.q, np = SL.np;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nl * d * d; idx += gridDim.x * blockDim.x) {
    const int kk = idx / (d * d), e = idx % (d * d), ia = e % d, ib = e / d;
    const int lo = ia < ib ? ia : ib, hi = ia < ib ? ib : ia;
    double s = stats[SL.T + (size_t) (SL.nlT > 1 ? kk : 0) * np + (lo * d - lo * (lo - 1) / 2 + (hi - lo))];
    const int k0 = nl > 1 ? kk : 0, k1 = nl > 1 ? kk + 1 : q;
    for (int k = k0; k < k1; k++) {
      const double ma = mu[k * d + ia], mb = mu[k * d + ib];
      const double sa = stats[SL.sk + k * d + ia], sb = stats[SL.sk + k * d + ib];
      s -= ma * sb + sa * mb - stats[SL.nk + k] * ma * mb;
    }
    S[idx] = s;
  }
}

// Column sums of Y per virtual shard (rows [lo_g, hi_g) of a matrix with leading dimension ldy): one CTA per
// (column, shard), 256 strided partial sums then a fixed tree.  The shard sums are folded in shard order on
// the host, so the column means do not depend on how many GPUs share the chain.
__global__ void __launch_bounds__(256) k_colsums_groups(const double *Y, int64_t ldy, const int64_t *group_lo,
                                                       int64_t row_off, int d, double *gsums) {
  __shared__ double red[256];
  const int c = blockIdx.x, g = blockIdx.y;
  const double *col = Y + (size_t) c * ldy - row_off;
  double s = 0.0;
  for (int64_t j = group_lo[g] + threadIdx.x; j < group_lo[g + 1]; j += 256) s += col[j];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gsums[(size_t) g * d + c] = red[0];
}

// Halo exchange of boundary labels: gather into / scatter from the contiguous NCCL buffers.
__global__ void k_halo_pack(const uint8_t *z, const int32_t *pos, int count, uint8_t *buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) buf[i] = z[pos[i]];
}
__global__ void k_halo_unpack(uint8_t *z, const int32_t *pos, int count, const uint8_t *buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) z[pos[i]] = buf[i];
}

void launch_colsums_groups(const double *Y, int64_t ldy, const int64_t *group_lo, int64_t row_off, int d, int g_first,
                           int g_count, double *gsums, cudaStream_t s) {
  dim3 grid(d, g_count);
  k_colsums_groups<<<grid, 256, 0, s>>>(Y, ldy, group_lo + g_first, row_off, d, gsums + (size_t) g_first * d);
  g_launches++;
}
void launch_halo_pack(const uint8_t *z, const int32_t *pos, int count, uint8_t *buf, cudaStream_t s) {
  if (count <= 0) return;
  k_halo_pack<<<(count + 255) / 256, 256, 0, s>>>(z, pos, count, buf);
  g_launches++;
}
void launch_halo_unpack(uint8_t *z, const int32_t *pos, int count, const uint8_t *buf, cudaStream_t s) {
  if (count <= 0) return;
  k_halo_unpack<<<(count + 255) / 256, 256, 0, s>>>(z, pos, count, buf);
  g_launches++;
}

void launch_colmeans(const double *Y, int64_t n, int d, double *ybar, cudaStream_t s) {
  k_colmeans<<<d, 256, 0, s>>>(Y, n, ybar);
  g_launches++;
}
void launch_permute_center(const double *Y, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                           const double *ybar, double *Yp, int64_t dst_ld, cudaStream_t s, int64_t orig_off) {
  k_permute_center<<<(unsigned) ((count + 255) / 256), 256, 0, s>>>(Y, src_ld, count, d, orig, ybar, Yp, dst_ld, orig_off);
  g_launches++;
}
__global__ void k_fill_f64(double *p, int64_t n, double v) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_f64(double *p, int64_t n, double v, cudaStream_t s) {
  if (n <= 0) return;
  k_fill_f64<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(p, n, v);
  g_launches++;
}
void launch_snapshot(const uint8_t *z, const double *w, const int32_t *orig, int64_t n, int32_t *z_out,
                     double *w_out, cudaStream_t s, int64_t orig_off) {
  k_snapshot<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(z, w, orig, n, z_out, w_out, orig_off);
  g_launches++;
}
void launch_snapshot_Y(const double *Yp, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                       const double *ybar, double *out, int64_t dst_ld, cudaStream_t s) {
  k_snapshot_Y<<<(unsigned) ((count + 255) / 256), 256, 0, s>>>(Yp, src_ld, count, d, orig, ybar, out, dst_ld);
  g_launches++;
}
void launch_mode_accumulate(const int32_t *z1, int64_t n, int q, uint32_t row, uint32_t *cnt, uint32_t *first,
                            cudaStream_t s) {
  if (n <= 0) return;
  k_mode_accumulate<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(z1, n, q, row, cnt, first);
  g_launches++;
}
void launch_mode_final(const uint32_t *cnt, const uint32_t *first, int64_t n, int q, int32_t *out, cudaStream_t s) {
  if (n <= 0) return;
  k_mode_final<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(cnt, first, n, q, out);
  g_launches++;
}
void launch_running_mean(const double *snap, double *mean, int64_t len, double inv_count, cudaStream_t s) {
  k_running_mean<<<(unsigned) ((len + 255) / 256), 256, 0, s>>>(snap, mean, len, inv_count);
  g_launches++;
}
void launch_reduce_groups(const double *partials, const int32_t *group_start, int ngroups, int E,
                          double *gsum, cudaStream_t s, const uint32_t *gepoch, size_t par_stride) {
  dim3 grid((E + 31) / 32, ngroups);
  k_reduce_groups<<<grid, 256, 0, s>>>(partials, group_start, E, gsum, gepoch, par_stride);
  g_launches++;
}
void launch_factor(const double *mat, int nl, int d, int form, double *rooti, double *logdet, int *err,
                   cudaStream_t s) {
  k_factor<<<nl, 32, 3 * (size_t) d * d * sizeof(double), s>>>(mat, d, form, rooti, logdet, err);
  g_launches++;
}
void launch_scatter_probe(const double *stats, const StatLayout &SL, const double *mu, int nl, double *S,
                          cudaStream_t s) {
  k_scatter_probe<<<8, 256, 0, s>>>(stats, SL, mu, nl, S);
  g_launches++;
}

}   // namespace bsb
