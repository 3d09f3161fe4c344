// init_cluster.cu -- the step BEFORE the sampler: initial cluster labels, on the device.
//
// Replaces .init_cluster (/root/reference/R/spatialCluster.R:317-328), which calls stats::kmeans(Y, centers = q)
// or mclust::Mclust(Y, q, "EEE") in R.  At Visium-HD sizes (700k .. 10M spots) that R step takes far longer than the
// sampler itself once the sampler runs on a B200.  Both third-party routines draw from R's RNG and use
// algorithms of their own (Hartigan-Wong; model-based agglomeration + EM), so their labels cannot be reproduced bit for
// bit; what is kept is the contract the sampler relies on -- n labels in 1..q, every label present, a sensible
// partition of the PCs, deterministic under the seed -- with standard algorithms:
//   method 0 "kmeans":  k-means++ seeding (D^2 sampling on Philox-keyed uniforms) + Lloyd iterations until no label
//                       changes (or max_iter);
//   method 1 "mclust":  EM for the Gaussian mixture with one shared covariance ("EEE": ellipsoidal, equal volume, shape
//                       and orientation), started from the k-means partition, until the relative change of the
//                       log-likelihood falls below 1e-5 (mclust's default tolerance) or max_iter; labels = argmax of
//                       the responsibilities.
// Everything that touches all n spots runs in kernels; the q x d and d x d algebra of a step runs on the host between
// them.  Sums over spots are folded in a fixed order (per-block column sums, then blocks in index order), so a call
// is reproducible and independent of the launch geometry of other work on the GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "chain.hpp"

namespace bsb {
namespace {

constexpr int kChunk = 1024;          // spots per block of the summation kernels
constexpr uint32_t P_INIT = 8;        // Philox purpose of the seeding draws: counter (step, 0, 0, P_INIT)

// Philox4x32-10 on the host (same function as csrc/rng.cuh; the seeding uniforms are drawn on the host).
void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1, n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
double seed_uniform(uint64_t seed, uint32_t step) {
  uint32_t o[4];
  philox_host(step, 0u, 0u, P_INIT, (uint32_t) seed, (uint32_t) (seed >> 32), o);
  const uint64_t x = ((uint64_t) o[1] << 32) | o[0];
  return ((double) (x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// d2[j] = min(d2[j], |y_j - c|^2); bsum[b] = sum of d2 over the block's spots, folded in index order by one thread
__global__ void __launch_bounds__(256) k_d2_update(const double *Y, int64_t n, int d, const double *centre, double *d2, int first,
                                                   double *bsum) {
  __shared__ double s_c[kMaxD];
  __shared__ double s_d2[kChunk];
  if (threadIdx.x < d) s_c[threadIdx.x] = centre[threadIdx.x];
  __syncthreads();
  const int64_t j0 = (int64_t) blockIdx.x * kChunk;
  for (int t = threadIdx.x; t < kChunk; t += 256) {
    const int64_t j = j0 + t;
    double v = 0.0;
    if (j < n) {
      for (int c = 0; c < d; c++) { const double x = Y[(size_t) c * n + j] - s_c[c]; v = fma(x, x, v); }
      if (!first) v = fmin(v, d2[j]);
      d2[j] = v;
    }
    s_d2[t] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int t = 0; t < kChunk; t++) s += s_d2[t];
    bsum[blockIdx.x] = s;
  }
}

// first index of block b whose running sum (index order) reaches target; the last spot of the block otherwise
__global__ void k_pick_in_block(const double *d2, int64_t n, int64_t b, double target, int64_t *picked) {
  const int64_t j0 = b * kChunk, j1 = min(n, j0 + kChunk);
  double s = 0.0;
  int64_t at = j1 - 1;
  for (int64_t j = j0; j < j1; j++) {
    s += d2[j];
    if (s >= target) { at = j; break; }
  }
  *picked = at;
}

// hard assignment to the nearest centre (ties: lowest index); counts label changes
__global__ void __launch_bounds__(256) k_assign(const double *Y, int64_t n, int d, int q, const double *centres, uint8_t *z,
                                                unsigned long long *changed) {
  extern __shared__ double s_cen[];   // q x d
  for (int e = threadIdx.x; e < q * d; e += 256) s_cen[e] = centres[e];
  __syncthreads();
  const int64_t j = (int64_t) blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  double y[kMaxD];
  for (int c = 0; c < d; c++) y[c] = Y[(size_t) c * n + j];
  int best = 0;
  double bd = __longlong_as_double(0x7ff0000000000000ll);
  for (int k = 0; k < q; k++) {
    double v = 0.0;
    for (int c = 0; c < d; c++) { const double x = y[c] - s_cen[k * d + c]; v = fma(x, x, v); }
    if (v < bd) { bd = v; best = k; }
  }
  if (z[j] != (uint8_t) best) { z[j] = (uint8_t) best; atomicAdd(changed, 1ull); }
}

// E step of the shared-covariance mixture: a_jk = lc[k] + y_j . g_k (the quadratic term y^T M y is common to all k);
// resp[k][j] = softmax_k(a_jk), z = argmax, per-block sum of logsumexp folded in index order
__global__ void __launch_bounds__(256) k_estep(const double *Y, int64_t n, int d, int q, const double *g, const double *lc,
                                               double *resp, uint8_t *z, double *blse) {
  extern __shared__ double sh[];   // g (q x d), lc (q), lse (256)
  double *s_g = sh, *s_lc = sh + q * d, *s_lse = s_lc + q;
  for (int e = threadIdx.x; e < q * d; e += 256) s_g[e] = g[e];
  for (int e = threadIdx.x; e < q; e += 256) s_lc[e] = lc[e];
  __syncthreads();
  const int64_t j = (int64_t) blockIdx.x * 256 + threadIdx.x;
  double lse = 0.0;
  if (j < n) {
    double y[kMaxD];
    for (int c = 0; c < d; c++) y[c] = Y[(size_t) c * n + j];
    double mx = -__longlong_as_double(0x7ff0000000000000ll);
    int best = 0;
    for (int k = 0; k < q; k++) {
      double a = s_lc[k];
      for (int c = 0; c < d; c++) a = fma(y[c], s_g[k * d + c], a);
      resp[(size_t) k * n + j] = a;
      if (a > mx) { mx = a; best = k; }
    }
    double s = 0.0;
    for (int k = 0; k < q; k++) s += exp(resp[(size_t) k * n + j] - mx);
    const double inv = 1.0 / s;
    for (int k = 0; k < q; k++) resp[(size_t) k * n + j] = exp(resp[(size_t) k * n + j] - mx) * inv;
    z[j] = (uint8_t) best;
    lse = mx + log(s);
  }
  s_lse[threadIdx.x] = lse;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int t = 0; t < 256; t++) s += s_lse[t];
    blse[blockIdx.x] = s;
  }
}

// Per-block cluster sums, deterministic: thread c owns column c (c < d: sum r y_c; c == d: sum r) of EVERY cluster and
// walks the block's spots in index order.  resp == nullptr: hard labels z (r = 1 for the spot's cluster).
// part: nblocks x q x (d + 1)
__global__ void __launch_bounds__(64) k_cluster_sums(const double *Y, int64_t n, int d, int q, const uint8_t *z, const double *resp,
                                                     double *part) {
  extern __shared__ double s_acc[];   // q x (d + 1)
  const int c = threadIdx.x, nc = d + 1;
  for (int e = threadIdx.x; e < q * nc; e += blockDim.x) s_acc[e] = 0.0;
  __syncthreads();
  const int64_t j0 = (int64_t) blockIdx.x * kChunk, j1 = min(n, j0 + kChunk);
  if (c < nc) {
    const double *col = c < d ? Y + (size_t) c * n : nullptr;
    for (int64_t j = j0; j < j1; j++) {
      const double v = col ? col[j] : 1.0;
      if (resp) {
        for (int k = 0; k < q; k++) s_acc[k * nc + c] = fma(resp[(size_t) k * n + j], v, s_acc[k * nc + c]);
      } else {
        s_acc[z[j] * nc + c] += v;
      }
    }
  }
  __syncthreads();
  double *out = part + (size_t) blockIdx.x * q * nc;
  for (int e = threadIdx.x; e < q * nc; e += blockDim.x) out[e] = s_acc[e];
}

// Per-block second moments sum_j y_a y_b (upper pairs), one pair per thread, spots in index order.  part: nblocks x np
__global__ void __launch_bounds__(256) k_pair_sums(const double *Y, int64_t n, int d, double *part) {
  const int np = d * (d + 1) / 2;
  const int64_t j0 = (int64_t) blockIdx.x * kChunk, j1 = min(n, j0 + kChunk);
  for (int p = threadIdx.x; p < np; p += 256) {
    int a = 0, rem = p, len = d;
    while (rem >= len) { rem -= len; len--; a++; }
    const int b = a + rem;
    const double *ca = Y + (size_t) a * n, *cb = Y + (size_t) b * n;
    double s = 0.0;
    for (int64_t j = j0; j < j1; j++) s = fma(ca[j], cb[j], s);
    part[(size_t) blockIdx.x * np + p] = s;
  }
}

// out[e] = sum over blocks (index order) of part[b][e]
__global__ void __launch_bounds__(256) k_fold_blocks(const double *part, int64_t nblocks, int len, double *out) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= len) return;
  double s = 0.0;
  for (int64_t b = 0; b < nblocks; b++) s += part[(size_t) b * len + e];
  out[e] = s;
}

__global__ void k_labels_out(const uint8_t *z, int64_t n, int32_t *out) {
  const int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (int32_t) z[j] + 1;
}

// Cholesky of a d x d SPD matrix (col-major, lower), in place; false if not positive definite
bool chol_lower_host(std::vector<double> &A, int d) {
  for (int j = 0; j < d; j++) {
    double s = A[(size_t) j * d + j];
    for (int k = 0; k < j; k++) s -= A[(size_t) k * d + j] * A[(size_t) k * d + j];
    if (!(s > 0.0)) return false;
    const double ljj = std::sqrt(s);
    A[(size_t) j * d + j] = ljj;
    for (int i = j + 1; i < d; i++) {
      double t = A[(size_t) j * d + i];
      for (int k = 0; k < j; k++) t -= A[(size_t) k * d + i] * A[(size_t) k * d + j];
      A[(size_t) j * d + i] = t / ljj;
    }
    for (int i = 0; i < j; i++) A[(size_t) j * d + i] = 0.0;
  }
  return true;
}

}   // namespace
}   // namespace bsb

using namespace bsb;

#define CKI(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(BSB200_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);   \
  } while (0)

extern "C" int bsb200_init_cluster(int32_t device, const double *Y, int64_t n, int32_t d, int32_t q, int32_t method,
                                   uint64_t seed, int32_t max_iter, int32_t *init, double *centers, int32_t *iters_done) {
  if (!Y || !init || n <= 0) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (d < 1 || d > kMaxD) return fail(BSB200_ERR_UNSUPPORTED, "d = %d unsupported (1..%d)", d, kMaxD);
  if (q < 1 || q > 255 || q > n) return fail(BSB200_ERR_UNSUPPORTED, "q = %d unsupported (1..min(255, n))", q);
  if (method != BSB200_INIT_KMEANS && method != BSB200_INIT_MCLUST) return fail(BSB200_ERR_INVALID, "unknown init.method %d", method);
  if ((size_t) q * (d + 1) * sizeof(double) > 40 * 1024) return fail(BSB200_ERR_UNSUPPORTED, "q (d + 1) = %d too large for the initialiser", q * (d + 1));
  if (max_iter <= 0) max_iter = 100;
  int rc = select_device(device);
  if (rc) return rc;
  const int nc = d + 1, np = d * (d + 1) / 2;
  const int64_t nblocks = (n + kChunk - 1) / kChunk, nb256 = (n + 255) / 256;
  DevBuf<double> dY, dd2, dbsum, dcen, dpart, dsums, dresp, dg, dlc, dblse, dT;
  DevBuf<uint8_t> dz;
  DevBuf<int32_t> dout;
  DevBuf<int64_t> dpick;
  DevBuf<unsigned long long> dchanged;
  if ((rc = dY.alloc((size_t) n * d)) || (rc = dd2.alloc((size_t) n)) || (rc = dbsum.alloc((size_t) nblocks)) ||
      (rc = dcen.alloc((size_t) q * d)) || (rc = dpart.alloc((size_t) nblocks * std::max(q * nc, np))) ||
      (rc = dsums.alloc((size_t) std::max(q * nc, np))) || (rc = dz.alloc((size_t) n)) || (rc = dout.alloc((size_t) n)) ||
      (rc = dpick.alloc(1)) || (rc = dchanged.alloc(1)))
    return rc;
  CKI(cudaMemcpy(dY.p, Y, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  CKI(cudaMemset(dz.p, 0xff, (size_t) n));

  // ---- k-means++ seeding -------------------------------------------------------------------------------------------
  std::vector<double> cen((size_t) q * d), bsum((size_t) nblocks);
  auto fetch_centre = [&](int k, int64_t j) -> int {
    for (int c = 0; c < d; c++) CKI(cudaMemcpy(&cen[(size_t) k * d + c], dY.p + (size_t) c * n + j, sizeof(double), cudaMemcpyDeviceToHost));
    return BSB200_OK;
  };
  int64_t first = (int64_t) std::floor(seed_uniform(seed, 0) * (double) n);
  if (first >= n) first = n - 1;
  if ((rc = fetch_centre(0, first))) return rc;
  for (int k = 1; k < q; k++) {
    CKI(cudaMemcpy(dcen.p, &cen[(size_t) (k - 1) * d], sizeof(double) * d, cudaMemcpyHostToDevice));
    k_d2_update<<<(unsigned) nblocks, 256>>>(dY.p, n, d, dcen.p, dd2.p, k == 1, dbsum.p);
    g_launches++;
    CKI(cudaMemcpy(bsum.data(), dbsum.p, sizeof(double) * nblocks, cudaMemcpyDeviceToHost));
    double total = 0.0;
    for (int64_t b = 0; b < nblocks; b++) total += bsum[(size_t) b];
    int64_t pick;
    if (!(total > 0.0)) {   // fewer than q distinct points: fall back to equally spaced spots
      pick = (int64_t) ((double) k / q * (double) n);
    } else {
      const double target = seed_uniform(seed, (uint32_t) k) * total;
      double run = 0.0;
      int64_t b = 0;
      for (; b < nblocks - 1; b++) {
        if (run + bsum[(size_t) b] >= target) break;
        run += bsum[(size_t) b];
      }
      k_pick_in_block<<<1, 1>>>(dd2.p, n, b, target - run, dpick.p);
      g_launches++;
      CKI(cudaMemcpy(&pick, dpick.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    if ((rc = fetch_centre(k, pick))) return rc;
  }

  // ---- Lloyd iterations --------------------------------------------------------------------------------------------
  std::vector<double> sums((size_t) std::max(q * nc, np));
  int it = 0;
  auto cluster_sums = [&](const double *resp) -> int {
    k_cluster_sums<<<(unsigned) nblocks, 64, sizeof(double) * q * nc>>>(dY.p, n, d, q, dz.p, resp, dpart.p);
    k_fold_blocks<<<(q * nc + 255) / 256, 256>>>(dpart.p, nblocks, q * nc, dsums.p);
    g_launches += 2;
    CKI(cudaMemcpy(sums.data(), dsums.p, sizeof(double) * q * nc, cudaMemcpyDeviceToHost));
    return BSB200_OK;
  };
  for (it = 0; it < max_iter; it++) {
    CKI(cudaMemcpy(dcen.p, cen.data(), sizeof(double) * q * d, cudaMemcpyHostToDevice));
    CKI(cudaMemset(dchanged.p, 0, sizeof(unsigned long long)));
    k_assign<<<(unsigned) nb256, 256, sizeof(double) * q * d>>>(dY.p, n, d, q, dcen.p, dz.p, dchanged.p);
    g_launches++;
    unsigned long long changed = 0;
    CKI(cudaMemcpy(&changed, dchanged.p, sizeof(changed), cudaMemcpyDeviceToHost));
    if (changed == 0) break;
    if ((rc = cluster_sums(nullptr))) return rc;
    for (int k = 0; k < q; k++) {
      const double nk = sums[(size_t) k * nc + d];
      if (nk > 0.0)   // an empty cluster keeps its centre
        for (int c = 0; c < d; c++) cen[(size_t) k * d + c] = sums[(size_t) k * nc + c] / nk;
    }
  }
  int iters = it;

  // ---- EM for the shared-covariance mixture ("EEE") ----------------------------------------------------------------
  if (method == BSB200_INIT_MCLUST && q > 1) {
    if ((rc = dresp.alloc((size_t) n * q)) || (rc = dg.alloc((size_t) q * d)) || (rc = dlc.alloc((size_t) q)) ||
        (rc = dblse.alloc((size_t) nb256)) || (rc = dT.alloc((size_t) np)))
      return rc;
    std::vector<double> T((size_t) np), blse((size_t) nb256);
    k_pair_sums<<<(unsigned) nblocks, 256>>>(dY.p, n, d, dpart.p);
    k_fold_blocks<<<(np + 255) / 256, 256>>>(dpart.p, nblocks, np, dT.p);
    g_launches += 2;
    CKI(cudaMemcpy(T.data(), dT.p, sizeof(double) * np, cudaMemcpyDeviceToHost));
    if ((rc = cluster_sums(nullptr))) return rc;   // M step from the k-means partition
    std::vector<double> pi((size_t) q), mu((size_t) q * d), Sig((size_t) d * d), Lc, M((size_t) d * d), g((size_t) q * d), lc((size_t) q);
    double prev = -std::numeric_limits<double>::infinity();
    const double LOG2PI = 1.8378770664093454835606594728112;
    for (it = 0; it < max_iter; it++) {
      // M step: pi_k, mu_k, Sigma = (T - sum_k n_k mu_k mu_k^T) / n
      for (int k = 0; k < q; k++) {
        const double nk = std::max(sums[(size_t) k * nc + d], 1e-300);
        pi[(size_t) k] = nk / (double) n;
        for (int c = 0; c < d; c++) mu[(size_t) k * d + c] = sums[(size_t) k * nc + c] / nk;
      }
      int p = 0;
      for (int a = 0; a < d; a++)
        for (int b = a; b < d; b++, p++) {
          double s = T[(size_t) p];
          for (int k = 0; k < q; k++) s -= sums[(size_t) k * nc + d] * mu[(size_t) k * d + a] * mu[(size_t) k * d + b];
          Sig[(size_t) a * d + b] = Sig[(size_t) b * d + a] = s / (double) n;
        }
      Lc = Sig;
      if (!chol_lower_host(Lc, d)) return fail(BSB200_ERR_NUMERIC, "init.method mclust: the pooled covariance is not positive definite");
      // M = Sigma^-1 through the factor; log|Sigma| = 2 sum log diag(L)
      double logdet = 0.0;
      for (int i = 0; i < d; i++) logdet += 2.0 * std::log(Lc[(size_t) i * d + i]);
      for (int col = 0; col < d; col++) {   // solve L L^T x = e_col
        std::vector<double> x((size_t) d, 0.0);
        for (int i = 0; i < d; i++) {
          double s = i == col ? 1.0 : 0.0;
          for (int k = 0; k < i; k++) s -= Lc[(size_t) k * d + i] * x[(size_t) k];
          x[(size_t) i] = s / Lc[(size_t) i * d + i];
        }
        for (int i = d - 1; i >= 0; i--) {
          double s = x[(size_t) i];
          for (int k = i + 1; k < d; k++) s -= Lc[(size_t) i * d + k] * x[(size_t) k];
          x[(size_t) i] = s / Lc[(size_t) i * d + i];
        }
        for (int i = 0; i < d; i++) M[(size_t) col * d + i] = x[(size_t) i];
      }
      double trMT = 0.0;
      p = 0;
      for (int a = 0; a < d; a++)
        for (int b = a; b < d; b++, p++) trMT += (a == b ? 1.0 : 2.0) * M[(size_t) a * d + b] * T[(size_t) p];
      for (int k = 0; k < q; k++) {
        double ck = 0.0;
        for (int i = 0; i < d; i++) {
          double s = 0.0;
          for (int j2 = 0; j2 < d; j2++) s += M[(size_t) j2 * d + i] * mu[(size_t) k * d + j2];
          g[(size_t) k * d + i] = s;
          ck += s * mu[(size_t) k * d + i];
        }
        lc[(size_t) k] = std::log(pi[(size_t) k]) - 0.5 * ck;
      }
      // E step
      CKI(cudaMemcpy(dg.p, g.data(), sizeof(double) * q * d, cudaMemcpyHostToDevice));
      CKI(cudaMemcpy(dlc.p, lc.data(), sizeof(double) * q, cudaMemcpyHostToDevice));
      k_estep<<<(unsigned) nb256, 256, sizeof(double) * (q * d + q + 256)>>>(dY.p, n, d, q, dg.p, dlc.p, dresp.p, dz.p, dblse.p);
      g_launches++;
      CKI(cudaMemcpy(blse.data(), dblse.p, sizeof(double) * nb256, cudaMemcpyDeviceToHost));
      double lse = 0.0;
      for (int64_t b = 0; b < nb256; b++) lse += blse[(size_t) b];
      const double loglik = lse - 0.5 * trMT - 0.5 * (double) n * (d * LOG2PI + logdet);
      iters++;
      if (std::fabs(loglik - prev) <= 1e-5 * std::fabs(loglik)) break;
      prev = loglik;
      if ((rc = cluster_sums(dresp.p))) return rc;
    }
    for (int k = 0; k < q; k++)
      for (int c = 0; c < d; c++) cen[(size_t) k * d + c] = mu[(size_t) k * d + c];
  }
  k_labels_out<<<(unsigned) nb256, 256>>>(dz.p, n, dout.p);
  g_launches++;
  CKI(cudaGetLastError());
  CKI(cudaMemcpy(init, dout.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  {   // every label 1..q must be present (the per-cluster Wishart of the VVV samplers cannot survive an empty cluster):
      // a label that lost all its spots gets one spot back, chosen by a keyed uniform
    std::vector<int64_t> cnt((size_t) q + 1, 0);
    for (int64_t j = 0; j < n; j++) cnt[(size_t) init[j]]++;
    for (int k = 1; k <= q; k++)
      if (cnt[(size_t) k] == 0) {
        int64_t j = (int64_t) std::floor(seed_uniform(seed, 1000u + (uint32_t) k) * (double) n);
        if (j >= n) j = n - 1;
        while (cnt[(size_t) init[j]] <= 1) j = (j + 1) % n;   // never empty another cluster
        cnt[(size_t) init[j]]--;
        init[j] = k;
        cnt[(size_t) k] = 1;
      }
  }
  if (centers) std::memcpy(centers, cen.data(), sizeof(double) * q * d);
  if (iters_done) *iters_done = iters;
  return BSB200_OK;
}
