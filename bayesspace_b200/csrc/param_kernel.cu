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

__global__ void __launch_bounds__(256) k_param(ParamArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int d = a.PL.d, q = a.PL.q, nl = a.PL.nl, np = a.PL.np, E = a.SL.E;
  const size_t dd = (size_t) d * d;
  const Key key{a.key0, a.key1};
  const uint32_t it_prev = *a.iter, it = it_prev + 1u;
  double *P = a.params;
  double *st = a.stats;
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

  double *scr = sm + (size_t) warp * (5 * dd + 4 * d);
  double *A0 = scr, *A1 = A0 + dd, *A2 = A1 + dd, *A3 = A2 + dd, *A4 = A3 + dd;
  double *v0 = A4 + dd, *v1 = v0 + d, *v2 = v1 + d, *v3 = v2 + d;
  bool ok = true;

  // ---- (1) mu_k ~ N(var (lambda0 mu0 + Lambda_k s_k), var), var = (lambda0 + n_k Lambda_k)^-1 ---
  for (int k = warp; k < q; k += nwarps) {
    const double *Lk = P + a.PL.lam + (size_t) (a.vvv ? k : 0) * dd;
    const double nk = st[a.SL.nk + k];
    const double *sk = st + a.SL.sk + (size_t) k * d;
    for (int e = lane; e < (int) dd; e += 32) A0[e] = a.lambda0[e] + nk * Lk[e];
    for (int c = lane; c < d; c += 32) v3[c] = sk[c];
    __syncwarp();
    ok &= warp_inv_spd(A0, A1, A2, d, lane);       // A1 = var
    warp_matvec(Lk, v3, v0, d, lane);              // v0 = Lambda_k s_k
    for (int c = lane; c < d; c += 32) v0[c] += a.lm0[c];
    __syncwarp();
    warp_matvec(A1, v0, v1, d, lane);              // v1 = posterior mean
    ok &= warp_chol_upper(A1, A2, d, lane);        // A2 = chol(var), upper
    for (int c = lane; c < d; c += 32) v2[c] = rng_normal(key, P_MU, (uint32_t) k, it, (uint32_t) c);
    __syncwarp();
    for (int j = lane; j < d; j += 32) {           // rmvnorm: z * chol(S) + mean
      double s = 0.0;
      for (int i = 0; i <= j; i++) s += v2[i] * A2[(size_t) j * d + i];
      P[a.PL.mu + k * d + j] = s + v1[j];
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- (2) Lambda ~ Wishart(df, (beta I + S)^-1); factorise once for the sweep ------------------
  const double *mu = P + a.PL.mu;
  for (int kk = warp; kk < nl; kk += nwarps) {
    const double *T = a.SL.nlT ? st + a.SL.T + (size_t) (a.SL.nlT > 1 ? kk : 0) * np : a.T_fixed;
    // S = sum_j w_j (y_j - mu_{z_j})(y_j - mu_{z_j})^T from the moment sums
    for (int e = lane; e < (int) dd; e += 32) {
      const int ia = e % d, ib = e / d;
      const int lo = ia < ib ? ia : ib, hi = ia < ib ? ib : ia;
      double s = T[pair_index(lo, hi, d)];
      const int k0 = a.vvv ? kk : 0, k1 = a.vvv ? kk + 1 : q;
      for (int k = k0; k < k1; k++) {
        const double ma = mu[k * d + ia], mb = mu[k * d + ib];
        const double sa = st[a.SL.sk + k * d + ia], sb = st[a.SL.sk + k * d + ib];
        s -= ma * sb + sa * mb - st[a.SL.nk + k] * ma * mb;
      }
      A0[e] = s + (ia == ib ? a.beta : 0.0);
    }
    __syncwarp();
    ok &= warp_inv_spd(A0, A1, A2, d, lane);       // A1 = (V + S)^-1
    ok &= warp_chol_upper(A1, A2, d, lane);        // A2 = chol(scale)
    const int df = (int) ((a.vvv ? st[a.SL.cnt + kk] : (double) a.n) + a.alpha);   // Q4: int df
    for (int e = lane; e < (int) dd; e += 32) {    // Bartlett factor, lower triangular
      const int i = e % d, j = e / d;
      A3[e] = i > j ? rng_normal(key, P_WISH_NORM, (uint32_t) kk, it, (uint32_t) (i * (i - 1) / 2 + j)) : 0.0;
    }
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      const double chi = rng_gamma(key, P_WISH_CHI, (uint32_t) kk * 1024u + (uint32_t) i, it, (df - i) / 2.0, 2.0);
      if (!(chi > 0.0)) ok = false;                // Q9: df - i <= 0 (empty cluster)
      A3[(size_t) i * d + i] = sqrt(chi);
    }
    __syncwarp();
    warp_matmul_tn(A3, A2, A0, d, lane);           // B = A^T chol(S)
    warp_matmul_tn(A0, A0, A4, d, lane);           // Lambda = B^T B
    double *Lnew = P + a.PL.lam + (size_t) kk * dd;
    for (int e = lane; e < (int) dd; e += 32) Lnew[e] = A4[e];
    pack_sym(A4, P + a.PL.PL + (size_t) kk * np, d, lane);
    ok &= warp_chol_upper(A4, A0, d, lane);        // A0 = chol(Lambda)
    const double *M = A4;
    double ld = 0.0;
    if ((a.compat & BSB_COMPAT_Q1_BIT) && !a.precision_form) {
      // dmvnrm_arma_fast: rooti = inv(chol(Sigma)), z = rooti * x  =>  M = rooti^T rooti
      warp_inv_upper(A0, A1, d, lane);
      warp_ttt(A1, A2, d, lane);                   // Sigma = Lambda^-1
      ok &= warp_chol_upper(A2, A3, d, lane);      // chol(Sigma)
      warp_inv_upper(A3, A0, d, lane);             // rooti
      warp_ttt_tn(A0, A1, d, lane);                // M = rooti^T rooti
      M = A1;
      for (int i = 0; i < d; i++) ld += log(A0[(size_t) i * d + i]);
    } else {
      for (int i = 0; i < d; i++) ld += log(A0[(size_t) i * d + i]);   // |Lambda|^(1/2)
    }
    pack_sym(M, P + a.PL.PM + (size_t) kk * np, d, lane);
    if (lane == 0) P[a.PL.logdet + kk] = ld;
    // derived per-cluster vectors: g = M mu, c = mu^T g, gL = Lambda mu, cL = mu^T gL
    const int k0 = a.vvv ? kk : 0, k1 = a.vvv ? kk + 1 : q;
    for (int k = k0; k < k1; k++) {
      for (int c = lane; c < d; c += 32) v3[c] = mu[k * d + c];
      __syncwarp();
      warp_matvec(M, v3, v0, d, lane);
      warp_matvec(A4, v3, v1, d, lane);
      for (int c = lane; c < d; c += 32) {
        P[a.PL.g + k * d + c] = v0[c];
        P[a.PL.gL + k * d + c] = v1[c];
      }
      if (lane == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int c = 0; c < d; c++) { s0 += v3[c] * v0[c]; s1 += v3[c] * v1[c]; }
        P[a.PL.c + k] = s0;
        P[a.PL.cL + k] = s1;
      }
      __syncwarp();
    }
  }
  if (!ok) atomicExch(a.err, BSB_DEVERR_NUMERIC);
  __syncthreads();

  // ---- (3) snapshots of mu / Lambda at thin boundaries (:309-314) and iteration counter ---------
  if ((it + 1u) % (uint32_t) a.thin == 0u) {
    const size_t r = (it + 1u) / (uint32_t) a.thin;
    if (r > (size_t) (a.nrep / a.thin)) { if (tid == 0) *a.iter = it; return; }
    if (a.snap_mu)
      for (int e = tid; e < q * d; e += blockDim.x) a.snap_mu[r * q * d + e] = P[a.PL.mu + e] + a.ybar[e % d];
    if (a.snap_lambda)
      for (int e = tid; e < nl * (int) dd; e += blockDim.x) a.snap_lambda[r * nl * dd + e] = P[a.PL.lam + e];
  }
  if (tid == 0) *a.iter = it;
}

void launch_param(const ParamArgs &a, cudaStream_t stream) {
  const int d = a.PL.d;
  const int nwarps = d <= 16 ? 8 : 4;
  const size_t smem = (size_t) nwarps * (5 * (size_t) d * d + 4 * d) * sizeof(double);
  k_param<<<1, nwarps * 32, smem, stream>>>(a);
}

void param_kernel_init() {
  cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

}   // namespace bsb
