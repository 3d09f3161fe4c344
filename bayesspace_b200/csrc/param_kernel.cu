// param_kernel.cu -- k_param: the once-per-iteration "small" kernel of every sampler.
//
// One CTA.  (0) folds the per-segment statistics of the previous sweep in a fixed two-level order
// (bit-reproducible, independent of grid size and GPU count) and closes plogLik of that sweep;
// (1) draws mu_k | . for every cluster (one warp per cluster); (2) draws Lambda | . from the
// Wishart (Bartlett) and factorises it ONCE for the whole sweep; (3) publishes the parameter
// block (ParamLayout) the sweep CTAs stage into shared memory.
//
// Reference steps replaced (file:line under /root/reference/src/cluster.cpp):
//   mu update      :246-258  :359-372  :478-493  :609-625  :883-892
//   lambda update  :260-270  :374-390  :495-505  :627-644  :896-902
//   chol/inv that dmvnrm_arma_fast repeats per spot (:91) and dmvnrm_prec_arma_fast (:117)
//   plogLik[i] = sum(plogLikj) :307 :430 :553 :695, accu(log_likelihoods) :1056
// RcppDist::rmvnorm / rwish semantics: SURVEY.md App. C.
//
// Linear algebra.  The reference forms inverses and then factorises them (inv(P) then chol inside
// rmvnorm; inv(V+S) then chol inside rwish; inv(Lambda) then chol then inv inside dmvnrm_arma_fast).
// chol_upper(X^-1) is the inverse of the "reverse" Cholesky factor U of X (X = U U^T, U upper), and
// inv(chol_upper(Lambda^-1)) is that factor of Lambda itself, so each of those chains is one
// factorisation (+ one triangular inverse) here.  Same values up to rounding.
#include "chain.hpp"
#include "rng.cuh"
#include "smallmat.cuh"

namespace bsb {

__device__ __forceinline__ int pair_index(int a, int b, int d) {   // a <= b
  return a * d - a * (a - 1) / 2 + (b - a);
}

__device__ inline void pack_sym(const double *M, double *P, int d, int lane) {
  for (int a = lane; a < d; a += 32)
    for (int b = a; b < d; b++) P[pair_index(a, b, d)] = (a == b ? 1.0 : 2.0) * M[(size_t) b * d + a];
  __syncwarp();
}

// g = M mu_k, c = mu_k^T g, gL = Lambda mu_k, cL = mu_k^T gL  (the expanded quadratic forms of the sweep)
__device__ inline void derive_cluster_vectors(double *P, const ParamLayout &PL, const double *mu_s, const double *M,
                                              const double *Lam, int k, double *v0, double *v1, int d, int lane) {
  const double *muk = mu_s + k * d;
  warp_matvec(M, muk, v0, d, lane);
  warp_matvec(Lam, muk, v1, d, lane);
  double s0 = 0.0, s1 = 0.0;
  for (int c = lane; c < d; c += 32) {
    P[PL.g + k * d + c] = v0[c];
    P[PL.gL + k * d + c] = v1[c];
    s0 += muk[c] * v0[c];
    s1 += muk[c] * v1[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (lane == 0) {
    P[PL.c + k] = s0;
    P[PL.cL + k] = s1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(256) k_param(ParamArgs a, int stats_in_smem) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int d = a.PL.d, q = a.PL.q, nl = a.PL.nl, np = a.PL.np, E = a.SL.E;
  const size_t dd = (size_t) d * d;
  const Key key{a.key0, a.key1};
  const uint32_t it_prev = *a.iter, it = it_prev + 1u;
  double *P = a.params;
  double *mu_s = sm;                                      // q*d   new means
  double *eps_mu = mu_s + q * d;                          // q*d   N(0,1) draws of the mu update
  double *wish = eps_mu + q * d;                          // np    Bartlett draws (shared precision only)
  double *st = stats_in_smem ? wish + np : a.stats;       // reduced statistics
  double *scr0 = wish + np + (stats_in_smem ? E : 0);
  const double LOG2PI = 1.8378770664093454835606594728112;

  // ---- (0) deterministic two-level reduction of the segment partials ---------------------------
  for (int idx = tid; !a.gsum_ready && idx < E * a.ngroups; idx += blockDim.x) {
    const int e = idx % E, g = idx / E;
    double s = 0.0;
    for (int slot = a.group_start[g]; slot < a.group_start[g + 1]; slot++) s += a.partials[(size_t) slot * E + e];
    a.gsum[(size_t) g * E + e] = s;
  }
  __syncthreads();
  for (int e = tid; e < E; e += blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < a.ngroups; g++) s += a.gsum[(size_t) g * E + e];
    st[e] = s;
    if (stats_in_smem) a.stats[e] = s;
  }
  __syncthreads();

  // plogLik of the sweep that just finished (uses the parameters of that sweep, still in P)
  if (it_prev >= 1u && warp == 0) {
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
  }
  if (a.finalize_only) return;
  if (*a.err) return;
  __syncthreads();

  // every variate of this iteration that does not depend on the statistics is drawn up front, one per
  // thread (the transcendental chains of Box-Muller / Marsaglia-Tsang are long when run by a single warp)
  for (int e = tid; e < q * d; e += blockDim.x) eps_mu[e] = rng_normal(key, P_MU, (uint32_t) (e / d), it, (uint32_t) (e % d));
  if (nl == 1) {
    const int df = (int) ((double) a.n + a.alpha);         // Q4: int df
    for (int p = tid; p < np; p += blockDim.x) {
      if (p < np - d) {
        wish[d + p] = rng_normal(key, P_WISH_NORM, 0u, it, (uint32_t) p);   // A(i,j), i > j, index i(i-1)/2 + j
      } else {
        const int i = p - (np - d);
        wish[i] = rng_gamma(key, P_WISH_CHI, (uint32_t) i, it, (df - i) / 2.0, 2.0);
      }
    }
  }
  __syncthreads();

  double *scr = scr0 + (size_t) warp * (4 * dd + 4 * d);
  double *A0 = scr, *A1 = A0 + dd, *A2 = A1 + dd, *A3 = A2 + dd;
  double *v0 = A3 + dd, *v1 = v0 + d, *v2 = v1 + d, *v3 = v2 + d;
  bool ok = true;

  // ---- (1) mu_k ~ N(var (lambda0 mu0 + Lambda_k s_k), var), var = (lambda0 + n_k Lambda_k)^-1 ---
  // rmvnorm(1, mean, var) = eps chol(var) + mean with chol(var) = R = U^-1, lambda0 + n_k Lambda_k = U U^T,
  // and mean = var rhs = R^T (R rhs):   mu_k = R^T (eps + R rhs).
  for (int k = warp; k < q; k += nwarps) {
    const double *Lk = P + a.PL.lam + (size_t) (a.vvv ? k : 0) * dd;
    const double nk = st[a.SL.nk + k];
    for (int e = lane; e < (int) dd; e += 32) {
      const double l = Lk[e];
      A3[e] = l;
      A0[e] = a.lambda0[e] + nk * l;
    }
    for (int c = lane; c < d; c += 32) v3[c] = st[a.SL.sk + k * d + c];
    __syncwarp();
    ok &= warp_revchol_upper(A0, A1, d, lane);     // A1 = U
    warp_inv_upper(A1, A2, d, lane);               // A2 = R = chol_upper(var)
    warp_matvec(A3, v3, v0, d, lane);              // v0 = Lambda_k s_k
    for (int c = lane; c < d; c += 32) v0[c] += a.lm0[c];
    __syncwarp();
    warp_triu_matvec(A2, v0, v1, d, lane);         // v1 = R rhs
    for (int c = lane; c < d; c += 32) v1[c] += eps_mu[k * d + c];
    __syncwarp();
    warp_triu_matvec_t(A2, v1, v2, d, lane);       // v2 = R^T (eps + R rhs)
    for (int c = lane; c < d; c += 32) {
      mu_s[k * d + c] = v2[c];
      P[a.PL.mu + k * d + c] = v2[c];
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- (2) Lambda ~ Wishart(df, (beta I + S)^-1); factorise once for the sweep ------------------
  for (int kk = warp; kk < nl; kk += nwarps) {
    const double *T = a.SL.nlT ? st + a.SL.T + (size_t) (a.SL.nlT > 1 ? kk : 0) * np : a.T_fixed;
    // V + S, S = sum_j w_j (y_j - mu_{z_j})(y_j - mu_{z_j})^T from the moment sums (upper triangle)
    for (int e = lane; e < (int) dd; e += 32) {
      const int ia = e % d, ib = e / d;
      if (ia > ib) { A0[e] = 0.0; continue; }
      double s = T[pair_index(ia, ib, d)];
      const int k0 = a.vvv ? kk : 0, k1 = a.vvv ? kk + 1 : q;
      for (int k = k0; k < k1; k++) {
        const double ma = mu_s[k * d + ia], mb = mu_s[k * d + ib];
        const double sa = st[a.SL.sk + k * d + ia], sb = st[a.SL.sk + k * d + ib];
        s -= ma * sb + sa * mb - st[a.SL.nk + k] * ma * mb;
      }
      A0[e] = s + (ia == ib ? a.beta : 0.0);
    }
    // Bartlett factor (lower triangular)
    const int df = (int) ((a.vvv ? st[a.SL.cnt + kk] : (double) a.n) + a.alpha);   // Q4: int df
    for (int e = lane; e < (int) dd; e += 32) {
      const int i = e % d, j = e / d;
      if (i <= j) A3[e] = 0.0;
      else A3[e] = nl == 1 ? wish[d + i * (i - 1) / 2 + j]
                           : rng_normal(key, P_WISH_NORM, (uint32_t) kk, it, (uint32_t) (i * (i - 1) / 2 + j));
    }
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      const double chi = nl == 1 ? wish[i]
                                 : rng_gamma(key, P_WISH_CHI, (uint32_t) kk * 1024u + (uint32_t) i, it, (df - i) / 2.0, 2.0);
      if (!(chi > 0.0)) ok = false;                // Q9: df - i <= 0 (empty cluster)
      A3[(size_t) i * d + i] = sqrt(chi);
    }
    __syncwarp();
    ok &= warp_revchol_upper(A0, A1, d, lane);     // V + S = U U^T
    warp_inv_upper(A1, A2, d, lane);               // A2 = C = chol_upper((V + S)^-1)
    // B = A^T C (upper triangular), Lambda = B^T B
    for (int j = lane; j < d; j += 32)
      for (int i = 0; i < d; i++) {
        double s = 0.0;
        for (int k = i; k <= j; k++) s += A3[(size_t) i * d + k] * A2[(size_t) j * d + k];
        A0[(size_t) j * d + i] = s;
      }
    __syncwarp();
    warp_ttt_tn(A0, A1, d, lane);                  // A1 = Lambda
    double *Lnew = P + a.PL.lam + (size_t) kk * dd;
    for (int e = lane; e < (int) dd; e += 32) Lnew[e] = A1[e];
    pack_sym(A1, P + a.PL.PL + (size_t) kk * np, d, lane);
    ok &= warp_revchol_upper(A1, A2, d, lane);     // Lambda = rooti rooti^T, rooti = inv(chol_upper(Sigma))
    double ld = 0.0;
    for (int i = 0; i < d; i++) ld += log(A2[(size_t) i * d + i]);
    const double *M = A1;
    if ((a.compat & BSB_COMPAT_Q1_BIT) && !a.precision_form) {
      warp_ttt_tn(A2, A3, d, lane);                // quirk Q1: M = rooti^T rooti (dmvnrm_arma_fast, :96-100)
      M = A3;
    }
    pack_sym(M, P + a.PL.PM + (size_t) kk * np, d, lane);
    if (lane == 0) P[a.PL.logdet + kk] = ld;
    if (nl == 1) {   // shared precision: hand M / Lambda to the whole CTA (warp 0's scratch: A0 = M, A1 = Lambda)
      if (M != A1) for (int e = lane; e < (int) dd; e += 32) A0[e] = A3[e];
      else for (int e = lane; e < (int) dd; e += 32) A0[e] = A1[e];
      __syncwarp();
    } else {
      derive_cluster_vectors(P, a.PL, mu_s, M, A1, kk, v0, v1, d, lane);
    }
  }
  __syncthreads();
  if (nl == 1)   // derived vectors of every cluster, one warp per cluster
    for (int k = warp; k < q; k += nwarps) derive_cluster_vectors(P, a.PL, mu_s, scr0, scr0 + dd, k, v0, v1, d, lane);
  if (!ok) atomicExch(a.err, BSB_DEVERR_NUMERIC);
  __syncthreads();

  // ---- (3) snapshots of mu / Lambda at thin boundaries (:309-314) and iteration counter ---------
  if ((it + 1u) % (uint32_t) a.thin == 0u) {
    const size_t r = (it + 1u) / (uint32_t) a.thin;
    if (r <= (size_t) (a.nrep / a.thin)) {
      if (a.snap_mu)
        for (int e = tid; e < q * d; e += blockDim.x) a.snap_mu[r * q * d + e] = mu_s[e] + a.ybar[e % d];
      if (a.snap_lambda)
        for (int e = tid; e < nl * (int) dd; e += blockDim.x) a.snap_lambda[r * nl * dd + e] = P[a.PL.lam + e];
    }
  }
  if (tid == 0) *a.iter = it;
}

static const size_t kParamSmemMax = 220 * 1024;

void launch_param(const ParamArgs &a, cudaStream_t stream) {
  const int d = a.PL.d, q = a.PL.q;
  int nwarps = d <= 16 ? 8 : 4;
  const size_t per_warp = (4 * (size_t) d * d + 4 * d) * sizeof(double);
  const size_t base = (2 * (size_t) q * d + a.PL.np) * sizeof(double), stats_bytes = (size_t) a.SL.E * sizeof(double);
  const int stats_in_smem = base + stats_bytes + nwarps * per_warp <= kParamSmemMax;
  const size_t smem = base + (stats_in_smem ? stats_bytes : 0) + nwarps * per_warp;
  k_param<<<1, nwarps * 32, smem, stream>>>(a, stats_in_smem);
}

void param_kernel_init() {
  cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kParamSmemMax);
}

}   // namespace bsb
