// param_kernel.cu -- k_param: the once-per-iteration "small" kernel of every sampler.
//
// One CTA of 512 threads (256 for the run-time-d reference version).  (0) folds the per-segment statistics of the
// previous sweep in a fixed two-level
// order (bit-reproducible, independent of grid size and GPU count) and closes plogLik of that sweep;
// (1) draws mu_k | . for every cluster; (2) draws Lambda | . from the Wishart (Bartlett) and factorises it
// ONCE for the whole sweep; (3) publishes the parameter block (ParamLayout) the sweep CTAs stage into shared
// memory.
//
// Reference steps replaced (file:line under /root/reference/src/cluster.cpp):
//   mu update      :246-258  :359-372  :478-493  :609-625  :883-892
//   lambda update  :260-270  :374-390  :495-505  :627-644  :896-902
//   chol/inv that dmvnrm_arma_fast repeats per spot (:91) and dmvnrm_prec_arma_fast (:117)
//   plogLik[i] = sum(plogLikj) :307 :430 :553 :695, accu(log_likelihoods) :1056
// RcppDist::rmvnorm / rwish semantics: SURVEY.md App. C.
//
// Linear algebra.  The reference forms inverses and then factorises them (inv(P) then chol inside rmvnorm;
// inv(V+S) then chol inside rwish; inv(Lambda) then chol then inv inside dmvnrm_arma_fast).
// chol_upper(X^-1) is the inverse of the "reverse" Cholesky factor U of X (X = U U^T, U upper), and
// inv(chol_upper(Lambda^-1)) is that factor of Lambda itself, so each of those chains is one factorisation
// (+ one triangular inverse) here.  Same values up to rounding.
//
// This kernel is pure latency (d <= 32): every d x d operation is spread over the whole CTA and over ALL
// matrices of a phase at once (the q posterior precisions of the mu step; the q Wishart draws of the VVV
// samplers), one thread per matrix element, so the critical path is d steps of one FMA + a barrier.
#include <cstdlib>

#include "chain.hpp"
#include "rng.cuh"

namespace bsb {

#define MAT(M, b, i, j) (M)[(b) * dd + (j) * d + (i)]

__device__ __forceinline__ int pair_index(int a, int b, int d) {   // a <= b
  return a * d - a * (a - 1) / 2 + (b - a);
}

#ifdef BSB_PARAM_TIMING
__device__ long long g_param_clock[16];
#define BSB_T(i) do { if (threadIdx.x == 0) g_param_clock[i] = clock64(); } while (0)
#else
#define BSB_T(i) do { } while (0)
#endif

// U (upper, U U^T = A) in place on nb matrices whose lower triangles are zero.  Right-looking from the
// last column: scale column j by 1/sqrt(a_jj), then down-date the leading j x j block by its outer product.
// Threads form a 32 x 8 grid (row = tid & 31, column / matrix = tid >> 5): no integer division on the
// critical path.  Must be called by every thread of the CTA (256 threads).
__device__ __forceinline__ void blk_revchol(int off_U, int nb, int d, int *ok) {
  extern __shared__ double sm[];
  double *U = sm + off_U;   // shared-memory window: keeps the accesses LDS/STS instead of generic LD/ST
  const int dd = d * d;
  const int ti = threadIdx.x & 31, tk = threadIdx.x >> 5;
  for (int j = d - 1; j >= 0; j--) {
    __syncthreads();
    if (ti < j)
      for (int b = tk; b < nb; b += 8) MAT(U, b, ti, j) *= rsqrt(MAT(U, b, j, j));
    __syncthreads();
    if (ti == 31) {   // the diagonal itself: nobody reads it any more in this step
      for (int b = tk; b < nb; b += 8) {
        const double s = MAT(U, b, j, j);
        if (!(s > 0.0)) *ok = 0;
        MAT(U, b, j, j) = sqrt(s);
      }
    } else {
      for (int b = 0; b < nb; b++)
        for (int k = tk; k < j; k += 8)
          if (ti <= k) MAT(U, b, ti, k) -= MAT(U, b, ti, j) * MAT(U, b, k, j);
    }
  }
  __syncthreads();
}

// T = U^-1 for nb upper-triangular matrices, row by row from the bottom (T must not alias U).
// rd: nb*d scratch for the reciprocals of the diagonals.
__device__ __forceinline__ void blk_inv_upper(int off_U, int off_T, int off_rd, int nb, int d) {
  extern __shared__ double sm[];
  const double *U = sm + off_U;
  double *T = sm + off_T, *rd = sm + off_rd;
  const int dd = d * d;
  const int ti = threadIdx.x & 31, tk = threadIdx.x >> 5;
  if (ti < d)
    for (int b = tk; b < nb; b += 8) rd[b * d + ti] = 1.0 / MAT(U, b, ti, ti);
  __syncthreads();
  for (int i = d - 1; i >= 0; i--) {
    if (ti < d)
      for (int b = tk; b < nb; b += 8) {
        double v = 0.0;
        if (ti >= i) {
          double s = i == ti ? 1.0 : 0.0;
          for (int k = i + 1; k <= ti; k++) s -= MAT(U, b, i, k) * MAT(T, b, k, ti);
          v = s * rd[b * d + i];
        }
        MAT(T, b, i, ti) = v;
      }
    __syncthreads();
  }
}

// ---- register-resident warp versions (one warp per matrix, lane k owns column k) ----------------------
// The shared-memory versions above cost ~3,000 cycles per elimination step (dependent LDS -> DFMA -> STS
// chains); with the column of every lane in registers a step is one rsqrt, one column published to shared
// memory by its owner and j broadcast loads, which makes these ~10x faster.  D = d exactly (1..16): every loop
// is fully unrolled with static register indices and no run-time bound, which keeps the code small -- the kernel
// runs once per iteration between sweeps that flush L2, so its instructions are fetched from HBM every time
// (the first version, unrolled for DM = 16 with run-time d, was 445 KB of SASS and twice as slow at 10M spots
// as at 5k spots for that reason alone).
template <int D>
__device__ __noinline__ void warp_revchol_reg(double *U, int lane, int *ok) {
  double c[D];   // column `lane`, rows 0..lane; entries below the diagonal are never read (don't care)
#pragma unroll
  for (int i = 0; i < D; i++) c[i] = (lane < D && i <= lane) ? U[lane * D + i] : 0.0;
#pragma unroll
  for (int j = D - 1; j >= 0; j--) {
    if (lane == j) {   // column j is final: u_ij = a_ij / sqrt(a_jj), stored where the result lives anyway
      const double ajj = c[j];
      if (!(ajj > 0.0)) *ok = 0;
      const double r = rsqrt(ajj);
#pragma unroll
      for (int i = 0; i < j; i++) U[j * D + i] = c[i] * r;
      U[j * D + j] = ajj * r;
    }
    __syncwarp();
    if (j > 0) {
      const double ukj = lane < j ? U[j * D + lane] : 0.0;   // u_kj of this lane's own column k = lane
#pragma unroll
      for (int i = 0; i < j; i++) c[i] = fma(-U[j * D + i], ukj, c[i]);   // broadcast loads; lanes >= j: ukj = 0
    }
  }
  __syncwarp();
}

// T = U^-1 (upper), from the bottom row up; U and T in shared memory (may not alias).  Column `lane` of T lives in
// registers; as soon as t[i] is final it is taken out of the partial sums of all rows above (independent FMAs) instead
// of every row summing its dependent chain when its turn comes: d dependent steps instead of d (d - 1) / 2.
template <int D>
__device__ __noinline__ void warp_inv_upper_reg(const double *U, double *T, int lane) {
  double t[D], s[D];   // t[k] = 0 for k > lane
  if (lane < D) T[lane * D + lane] = 1.0 / U[lane * D + lane];   // reciprocal diagonals, read back as broadcasts
#pragma unroll
  for (int i = 0; i < D; i++) s[i] = lane == i ? 1.0 : 0.0;
  __syncwarp();
#pragma unroll
  for (int i = D - 1; i >= 0; i--) {
    t[i] = lane >= i ? s[i] * T[i * D + i] : 0.0;
#pragma unroll
    for (int i2 = 0; i2 < i; i2++) s[i2] = fma(-U[i * D + i2], t[i], s[i2]);   // broadcast loads of U[i2, i]
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < D; i++)
    if (lane < D) T[lane * D + i] = i <= lane ? t[i] : 0.0;
}

// D = exact dimension 1..32 (register path) or 0 (run-time d up to 32, shared-memory path)
template <int D>
__global__ void __launch_bounds__(D > 0 ? 512 : 256) k_param(ParamArgs a, int stats_in_smem, int batch) {
  extern __shared__ double sm[];
  __shared__ int s_ok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  // d stays a run-time value outside the factorisations: compile-time d would fully unroll every small loop below
  // (+120 KB of code that is fetched from HBM once per iteration for nothing)
  const int d = a.PL.d, q = a.PL.q, nl = a.PL.nl, np = a.PL.np, E = a.SL.E;
  const int dd = d * d;
  const Key key{a.key0, a.key1};
  const uint32_t it_prev = *a.iter, it = it_prev + 1u;
  double *P = a.params;
  double *mu_s = sm;                                      // q*d   new means
  double *va = mu_s + q * d;                              // q*d   work vectors
  double *vb = va + q * d;                                // q*d
  double *st = stats_in_smem ? vb + q * d : a.stats;      // reduced statistics
  double *MA = vb + q * d + (stats_in_smem ? E : 0);      // 4 work matrices per matrix of a batch
  double *MB = MA + (size_t) batch * dd, *MC = MB + (size_t) batch * dd, *MD = MC + (size_t) batch * dd;
  const double LOG2PI = 1.8378770664093454835606594728112;
  if (tid == 0) s_ok = 1;
  if (a.zero_counters && tid < a.n_zero) a.zero_counters[tid] = 0u;
  BSB_T(0);

  // ---- (0) deterministic two-level reduction of the segment partials ---------------------------
  double *const gsum = a.gsum + (a.gepoch ? (size_t) (*a.gepoch & 1u) * a.gsum_stride : (size_t) 0);
  if (!a.gsum_ready) {
    // Level 1 (chains with one virtual shard fold here; the others in k_reduce_groups): the slots of a shard are summed in
    // slot order, by 1, 2 or 4 adjacent lanes per sum when threads are to spare -- each lane a contiguous part of the
    // slots, the parts added in order -- so that a 250-slot fold (deconvolution at 30k subspots) is 8 rounds of eight
    // loads instead of 32.  The association depends on (E, slots) only: deterministic, and no other GPU count shares it.
    const int total = E * a.ngroups;
    int ways = 1;
    while (ways < 4 && total * ways * 2 <= nt) ways *= 2;
    for (int base = 0; base < total * ways; base += nt) {   // uniform trip count: the shuffles below are warp-wide
      const int idx = base + tid, pair = idx / ways, part = idx - pair * ways;
      const bool valid = pair < total;
      const int e = valid ? pair % E : 0, g = valid ? pair / E : 0;
      const int s0 = a.group_start[g], s1 = valid ? a.group_start[g + 1] : s0;
      const int chunk = (s1 - s0 + ways - 1) / ways;
      int slot = min(s1, s0 + part * chunk);
      const int end = min(s1, slot + chunk);
      const double *pp = a.partials + e;
      double s = 0.0;
      for (; slot + 8 <= end; slot += 8) {   // loads issued 8 at a time ahead of the adds
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(pp + (size_t) (slot + u) * E);
#pragma unroll
        for (int u = 0; u < 8; u++) s += v[u];
      }
      for (; slot < end; slot++) s += __ldcg(pp + (size_t) slot * E);
      const double p1 = __shfl_down_sync(0xffffffffu, s, 1), p2 = __shfl_down_sync(0xffffffffu, s, 2),
                   p3 = __shfl_down_sync(0xffffffffu, s, 3);
      if (valid && part == 0) gsum[(size_t) g * E + e] = ways == 1 ? s : (ways == 2 ? s + p1 : ((s + p1) + p2) + p3);
    }
  }
  __syncthreads();
  for (int e = tid; e < E; e += nt) {
    // shard sums folded in shard order (fixed association); loads issued 8 at a time ahead of the adds
    const double *gp = gsum + e;
    double s = 0.0;
    int g = 0;
    for (; g + 8 <= a.ngroups; g += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldcg(gp + (size_t) (g + u) * E);
#pragma unroll
      for (int u = 0; u < 8; u++) s += v[u];
    }
    for (; g < a.ngroups; g++) s += __ldcg(gp + (size_t) g * E);
    st[e] = s;
    if (stats_in_smem) a.stats[e] = s;
  }
  __syncthreads();
  BSB_T(1);

  // plogLik of the sweep that just finished (uses the parameters of that sweep, still in P: the density matrix and its
  // log-determinant are only replaced in phase (2)).  One warp; when a warp is idle during the factorisations of the mu
  // phase (fewer clusters than warps) it runs there, off the critical path.
  auto close_plogLik = [&]() {
    double pl;
    if (!a.vvv) {
      const double *T = a.SL.nlT ? st + a.SL.T : a.T_fixed;
      double tr = 0.0;
      for (int p = lane; p < np; p += 32) tr += P[a.PL.PM + p] * T[p];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
      pl = st[a.SL.hp] + (double) a.n * (P[a.PL.logdet] - 0.5 * d * LOG2PI) - 0.5 * tr;
    } else {
      pl = st[a.SL.hp] - (double) a.n * 0.5 * d * LOG2PI;
    }
    if (lane == 0 && (int) it_prev < a.nrep) {
      a.plogLik[it_prev] = pl;
      if (a.ychange) a.ychange[it_prev] = st[a.SL.nacc] / a.n0;   // src/cluster.cpp:1055
    }
  };
  const bool plog_late = !a.finalize_only && q < nw && D > 0;
  if (it_prev >= 1u && warp == 0 && !plog_late) close_plogLik();
  if (a.finalize_only) return;
  if (*a.err) return;
  __syncthreads();
  BSB_T(2);

  // ---- (1) mu_k ~ N(var (lambda0 mu0 + Lambda_k s_k), var), var = (lambda0 + n_k Lambda_k)^-1 ---
  // rmvnorm(1, mean, var) = eps chol(var) + mean with chol(var) = R = U^-1, lambda0 + n_k Lambda_k = U U^T,
  // and mean = var rhs = R^T (R rhs):   mu_k = R^T (eps + R rhs).   All clusters of a batch at once.
  // Bartlett factor A of the Wishart draw (lower triangular: N(0,1) below the diagonal, sqrt(chi2(df - i)) on it).
  auto bartlett = [&](int kk, int ia, int ib) {
    double av = 0.0;
    if (ia > ib) {
      av = rng_normal(key, P_WISH_NORM, (uint32_t) kk, it, (uint32_t) (ia * (ia - 1) / 2 + ib));
    } else if (ia == ib) {
      const int df = (int) ((a.vvv ? st[a.SL.cnt + kk] : (double) a.n) + a.alpha);   // Q4: int df
      const double chi = rng_gamma(key, P_WISH_CHI, (uint32_t) kk * 1024u + (uint32_t) ia, it, (df - ia) / 2.0, 2.0);
      if (!(chi > 0.0)) s_ok = 0;                // Q9: df - i <= 0 (empty cluster)
      av = sqrt(chi);
    }
    return av;
  };
  // Everything of the iteration that depends on the random numbers and the statistics alone -- the Bartlett factors
  // (-> MD, idle until phase (2)), the normals of the mean draws (-> vb) and the right-hand sides lambda0 mu0 +
  // Lambda_k s_k (-> va) -- is side work: when all matrices fit one batch, the warps without a factorisation in phase
  // (1) do it while the others factorise (4 of 16 warps are busy at q = 4), instead of every thread afterwards.
  const bool early = D > 0 && batch >= q;
  auto side_work = [&](int t0, int nthreads) {
    const int nA = nl * (int) dd, nE = q * d;
    for (int idx = t0; idx < nA + 2 * nE; idx += nthreads) {
      if (idx < nA) {
        const int b = idx / (int) dd, e = idx - b * (int) dd;
        MD[idx] = bartlett(b, e % d, e / d);
      } else if (idx < nA + nE) {
        const int r = idx - nA, k = r / d, i = r - k * d;
        vb[r] = rng_normal(key, P_MU, (uint32_t) k, it, (uint32_t) i);
      } else {
        const int r = idx - nA - nE, k = r / d, i = r - k * d;
        double s = a.lm0[i];
        for (int j = 0; j < d; j++) s += MAT(MC, k, i, j) * st[a.SL.sk + k * d + j];
        va[r] = s;
      }
    }
  };
  for (int k0 = 0; k0 < q; k0 += batch) {
    const int nb = min(batch, q - k0);
    for (int idx = tid; idx < nb * (int) dd; idx += nt) {
      const int b = idx / (int) dd, e = idx - b * (int) dd, i = e % d, j = e / d, k = k0 + b;
      const double l = P[a.PL.lam + (size_t) (a.vvv ? k : 0) * dd + e];
      MC[idx] = l;                                                                      // Lambda_k
      MA[idx] = i <= j ? a.lambda0[e] + st[a.SL.nk + k] * l : 0.0;                      // upper triangle of P_k
    }
    __syncthreads();
    if constexpr (D > 0) {
      // too few idle warps for the side work to finish before the factorisations: everybody, before them
      const int busy = min(nb, nw);
      const bool side_all = early && (nl * (int) dd + 2 * q * d) > 4 * (nt - busy * 32);
      if (plog_late && k0 == 0 && warp == nw - 1 && it_prev >= 1u) close_plogLik();   // an idle warp of this phase
      if (early && (side_all || warp >= busy)) side_work(side_all ? tid : tid - busy * 32, side_all ? nt : nt - busy * 32);
#pragma unroll 1   // one copy of the unrolled factorisations, not four
      for (int b = warp; b < nb; b += nw) {
        warp_revchol_reg<D>(MA + b * dd, lane, &s_ok);            // MA = U
        warp_inv_upper_reg<D>(MA + b * dd, MB + b * dd, lane);    // MB = R = chol_upper(var)
      }
      __syncthreads();
    } else {
      blk_revchol((int) (MA - sm), nb, d, &s_ok);
      blk_inv_upper((int) (MA - sm), (int) (MB - sm), (int) (va - sm), nb, d);
    }
    if (!early) {
      for (int idx = tid; idx < nb * d; idx += nt) {   // va = lambda0 mu0 + Lambda_k s_k
        const int b = idx / d, i = idx - b * d, k = k0 + b;
        double s = a.lm0[i];
        for (int j = 0; j < d; j++) s += MAT(MC, b, i, j) * st[a.SL.sk + k * d + j];
        va[k * d + i] = s;
      }
      __syncthreads();
    }
    for (int idx = tid; idx < nb * d; idx += nt) {   // vb = R rhs + eps
      const int b = idx / d, i = idx - b * d, k = k0 + b;
      double s = early ? vb[k * d + i] : rng_normal(key, P_MU, (uint32_t) k, it, (uint32_t) i);
      for (int j = i; j < d; j++) s += MAT(MB, b, i, j) * va[k * d + j];
      vb[k * d + i] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < nb * d; idx += nt) {   // mu = R^T vb
      const int b = idx / d, j = idx - b * d, k = k0 + b;
      double s = 0.0;
      for (int i = 0; i <= j; i++) s += MAT(MB, b, i, j) * vb[k * d + i];
      mu_s[k * d + j] = s;
      P[a.PL.mu + k * d + j] = s;
    }
    __syncthreads();
  }
  BSB_T(3);

  // ---- (2) Lambda ~ Wishart(df, (beta I + S)^-1); factorise once for the sweep ------------------
  for (int k0 = 0; k0 < nl; k0 += batch) {
    const int nb = min(batch, nl - k0);
    for (int idx = tid; idx < nb * (int) dd; idx += nt) {
      const int b = idx / (int) dd, e = idx - b * (int) dd, ia = e % d, ib = e / d, kk = k0 + b;
      // V + S (upper triangle), S = sum_j w_j (y_j - mu_{z_j})(y_j - mu_{z_j})^T from the moment sums
      double s = 0.0;
      if (ia <= ib) {
        const double *T = a.SL.nlT ? st + a.SL.T + (size_t) (a.SL.nlT > 1 ? kk : 0) * np : a.T_fixed;
        s = T[pair_index(ia, ib, d)] + (ia == ib ? a.beta : 0.0);
        const int c0 = a.vvv ? kk : 0, c1 = a.vvv ? kk + 1 : q;
        for (int k = c0; k < c1; k++) {
          const double ma = mu_s[k * d + ia], mb = mu_s[k * d + ib];
          const double sa = st[a.SL.sk + k * d + ia], sb = st[a.SL.sk + k * d + ib];
          s -= ma * sb + sa * mb - st[a.SL.nk + k] * ma * mb;
        }
      }
      MA[idx] = s;
      MC[idx] = early ? MD[idx] : bartlett(kk, ia, ib);   // Bartlett factor A
    }
    __syncthreads();
    BSB_T(6);
    if constexpr (D > 0) {
#pragma unroll 1   // one copy of the unrolled factorisations, not four
      for (int b = warp; b < nb; b += nw) {
        warp_revchol_reg<D>(MA + b * dd, lane, &s_ok);            // V + S = U U^T
        warp_inv_upper_reg<D>(MA + b * dd, MB + b * dd, lane);    // MB = C = chol_upper((V + S)^-1)
      }
      __syncthreads();
    BSB_T(7);
    } else {
      blk_revchol((int) (MA - sm), nb, d, &s_ok);
      blk_inv_upper((int) (MA - sm), (int) (MB - sm), (int) (va - sm), nb, d);
    }
    for (int idx = tid; idx < nb * (int) dd; idx += nt) {   // MD = B = A^T C (upper triangular)
      const int b = idx / (int) dd, e = idx - b * (int) dd, i = e % d, j = e / d;
      double s = 0.0;
      for (int k = i; k <= j; k++) s += MAT(MC, b, k, i) * MAT(MB, b, k, j);
      MD[idx] = s;
    }
    __syncthreads();
    BSB_T(8);
    for (int idx = tid; idx < nb * (int) dd; idx += nt) {   // Lambda = B^T B -> MB (full), MA (upper triangle)
      const int b = idx / (int) dd, e = idx - b * (int) dd, i = e % d, j = e / d, kk = k0 + b;
      const int k1 = i < j ? i : j;
      double s = 0.0;
      for (int k = 0; k <= k1; k++) s += MAT(MD, b, k, i) * MAT(MD, b, k, j);
      P[a.PL.lam + (size_t) kk * dd + e] = s;
      if (i <= j) P[a.PL.PL + (size_t) kk * np + pair_index(i, j, d)] = (i == j ? 1.0 : 2.0) * s;
      MC[idx] = s;
      MA[idx] = i <= j ? s : 0.0;
    }
    __syncthreads();
    BSB_T(9);
    const bool q1 = (a.compat & BSB_COMPAT_Q1_BIT) && !a.precision_form;
    // The factor of Lambda itself is only needed for quirk Q1 (M = rooti^T rooti).  Otherwise M = Lambda and the
    // log-determinant term, sum log diag(rooti) = log det(Lambda) / 2, equals sum log diag(B) of Lambda = B^T B (MD).
    if (q1) {
      if constexpr (D > 0) {
#pragma unroll 1
        for (int b = warp; b < nb; b += nw) warp_revchol_reg<D>(MA + b * dd, lane, &s_ok);   // Lambda = rooti rooti^T
        __syncthreads();
      } else {
        blk_revchol((int) (MA - sm), nb, d, &s_ok);
      }
    } else {
      for (int b = warp; b < nb; b += nw)
        if (lane < d) MAT(MA, b, lane, lane) = MAT(MD, b, lane, lane);
      __syncthreads();
    }
    BSB_T(10);
    for (int idx = tid; idx < nb * (int) dd; idx += nt) {   // density matrix M -> MD and packed into P
      const int b = idx / (int) dd, e = idx - b * (int) dd, i = e % d, j = e / d, kk = k0 + b;
      double s;
      if (q1) {   // quirk Q1: M = rooti^T rooti (dmvnrm_arma_fast multiplies rooti * x, :96-100)
        const int k1 = i < j ? i : j;
        s = 0.0;
        for (int k = 0; k <= k1; k++) s += MAT(MA, b, k, i) * MAT(MA, b, k, j);
      } else {
        s = MC[idx];
      }
      MD[idx] = s;
      if (i <= j) P[a.PL.PM + (size_t) kk * np + pair_index(i, j, d)] = (i == j ? 1.0 : 2.0) * s;
    }
    for (int b = warp; b < nb; b += nw) {           // log-determinant term: sum log diag(rooti)
      double ld = lane < d ? log(MAT(MA, b, lane, lane)) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
      if (lane == 0) P[a.PL.logdet + k0 + b] = ld;
    }
    __syncthreads();
    BSB_T(11);
    // derived per-cluster vectors: g = M mu, gL = Lambda mu (then c = mu^T g, cL = mu^T gL)
    const int c0 = a.vvv ? k0 : 0, c1 = a.vvv ? k0 + nb : q;
    for (int idx = tid; idx < (c1 - c0) * d; idx += nt) {
      const int k = c0 + idx / d, i = idx % d, b = a.vvv ? k - k0 : 0;
      double s0 = 0.0, s1 = 0.0;
      for (int j = 0; j < d; j++) {
        const double m = mu_s[k * d + j];
        s0 += MAT(MD, b, i, j) * m;
        s1 += MAT(MC, b, i, j) * m;
      }
      va[k * d + i] = s0;
      vb[k * d + i] = s1;
      P[a.PL.g + k * d + i] = s0;
      P[a.PL.gL + k * d + i] = s1;
    }
    __syncthreads();
    BSB_T(12);
    for (int k = c0 + tid; k < c1; k += nt) {
      double s0 = 0.0, s1 = 0.0;
      for (int i = 0; i < d; i++) { s0 += mu_s[k * d + i] * va[k * d + i]; s1 += mu_s[k * d + i] * vb[k * d + i]; }
      P[a.PL.c + k] = s0;
      P[a.PL.cL + k] = s1;
    }
    __syncthreads();
    BSB_T(13);
  }
  BSB_T(4);
  if (tid == 0 && !s_ok) atomicExch(a.err, BSB_DEVERR_NUMERIC);

  // ---- (3) snapshots of mu / Lambda at thin boundaries (:309-314) and iteration counter ---------
  if ((it + 1u) % (uint32_t) a.thin == 0u) {
    const size_t r = (it + 1u) / (uint32_t) a.thin;
    if (r <= (size_t) (a.nrep / a.thin)) {
      if (a.snap_mu)
        for (int e = tid; e < q * d; e += nt) a.snap_mu[r * q * d + e] = mu_s[e] + a.ybar[e % d];
      if (a.snap_lambda)
        for (int e = tid; e < nl * (int) dd; e += nt) a.snap_lambda[r * nl * dd + e] = P[a.PL.lam + e];
    }
  }
  if (tid == 0) *a.iter = it;
  BSB_T(5);
}

static const size_t kParamSmemMax = 220 * 1024;

typedef void (*param_kernel_t)(ParamArgs, int, int);
static param_kernel_t *param_kernel_table() {
#ifdef BSB_PARAM_DEV   // development builds: only the dimensions of the benchmark configurations (compile time)
  static param_kernel_t t[kMaxD + 1] = {k_param<0>};
  for (int i = 1; i <= kMaxD; i++) t[i] = k_param<0>;
  t[15] = k_param<15>;
  t[20] = k_param<20>;
  return t;
#else
  static param_kernel_t t[kMaxD + 1] = {
      k_param<0>,  k_param<1>,  k_param<2>,  k_param<3>,  k_param<4>,  k_param<5>,  k_param<6>,  k_param<7>,  k_param<8>,
      k_param<9>,  k_param<10>, k_param<11>, k_param<12>, k_param<13>, k_param<14>, k_param<15>, k_param<16>, k_param<17>,
      k_param<18>, k_param<19>, k_param<20>, k_param<21>, k_param<22>, k_param<23>, k_param<24>, k_param<25>, k_param<26>,
      k_param<27>, k_param<28>, k_param<29>, k_param<30>, k_param<31>, k_param<32>};
  return t;
#endif
}

size_t param_smem_min(int q, int d) { return (3 * (size_t) q * d + 4 * (size_t) d * d) * sizeof(double); }

void launch_param(const ParamArgs &a, cudaStream_t stream) {
  const int d = a.PL.d, q = a.PL.q;
  const size_t per_mat = 4 * (size_t) d * d * sizeof(double);
  const size_t base = 3 * (size_t) q * d * sizeof(double), stats_bytes = (size_t) a.SL.E * sizeof(double);
  const int want = q;   // matrices of the largest phase
  int stats_in_smem = base + stats_bytes + per_mat <= kParamSmemMax && stats_bytes <= 96 * 1024;
  size_t avail = kParamSmemMax - base - (stats_in_smem ? stats_bytes : 0);
  int batch = (int) std::min<size_t>((size_t) want, avail / per_mat);
  if (batch < 1) batch = 1;
  const size_t smem = base + (stats_in_smem ? stats_bytes : 0) + (size_t) batch * per_mat;
  // factorisations with one register-resident column per lane, kernel instantiated for the exact d (k_param<0>,
  // the shared-memory version with run-time d, stays as the reference implementation: BSB200_PARAM_GENERIC=1)
  static const bool generic = std::getenv("BSB200_PARAM_GENERIC") != nullptr;
  // 16 warps for the register-resident version: one cluster's factorisations per warp up to q = 16 in one round
  const bool reg = !generic && d > 0;
  param_kernel_table()[reg ? d : 0]<<<1, reg ? 512 : 256, smem, stream>>>(a, stats_in_smem, batch);
}

#ifdef BSB_PARAM_TIMING
extern "C" int bsb200_debug_param_timing(long long *out16) {
  return (int) cudaMemcpyFromSymbol(out16, g_param_clock, sizeof(long long) * 16);
}
#endif

void param_kernel_init() {
  for (int i = 0; i <= kMaxD; i++)
    cudaFuncSetAttribute(param_kernel_table()[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kParamSmemMax);
}

}   // namespace bsb
