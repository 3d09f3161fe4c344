// sweep_kernels.cu -- the z / w sweep of the four cluster samplers and the spot x cluster
// log-density probe, instantiated for ONE dimension d = BSB_D per translation unit (a spot's PCs
// live in registers, so d must be a compile-time constant).
//
// k_sweep<D, TDIST, VVV> replaces the sequential per-spot loops of
//   iterate        /root/reference/src/cluster.cpp:272-307
//   iterate_vvv    :392-430      iterate_t  :507-553      iterate_t_vvv  :646-695
// by a chromatic scan: one launch per colour class, one thread per spot, and fuses into the same
// pass the sufficient statistics the NEXT iteration's mu / lambda updates need
// (:248-253, :261-265, :482-486, :500), so Y is read once per sweep.
//
// Differences from the reference's arithmetic, all exact in real arithmetic:
//  * the covariance is factorised once per iteration in k_param, not twice per spot (:91);
//  * shared precision: the quadratic form is expanded, q(y, mu_k) = y^T M y - 2 y^T (M mu_k)
//    + mu_k^T M mu_k.  y^T M y cancels in the MH ratio; its sum over spots (needed only by the
//    plogLik diagnostic, :305-307) is tr(M sum_j w_j y_j y_j^T), taken from the statistics;
//  * sum_j log w_j (plogLik only) is accumulated as a product of mantissas, one log per segment;
//  * Y is centred by its column means at upload (translation invariance), which keeps the
//    expansion well conditioned.
#include "kernels.hpp"
#include "rng.cuh"
#include "tile_stats.cuh"

#ifndef BSB_D
#error "compile with -DBSB_D=<d>"
#endif

namespace bsb {
namespace {

// Shared-memory carve-up of k_sweep (offsets in doubles).
struct SweepSmem {
  int g, c, gL, cL, PL, PM, logdet, tile, sws, wvs, acc, accP, ct, red, ybuf, ellbuf, origbuf, zs_bytes_off, P;
  size_t bytes;
};
__host__ __device__ inline SweepSmem sweep_smem_layout(int D, int q, int nl, int np, int E, bool tdist, bool vvv,
                                                       bool vvv_pairs, int maxdeg) {
  SweepSmem s;
  const int DP = (D + 1 + 7) / 8 * 8, NFRAG = stats_frags(D, q > 16 ? 0 : (q + 7) / 8);
  int o = 0;
  s.g = o; o += q * D;
  s.c = o; o += q;
  s.gL = o; o += tdist ? q * D : 0;
  s.cL = o; o += tdist ? q : 0;
  s.PL = o; o += tdist ? nl * np : 0;
  s.PM = o; o += vvv ? nl * np : 0;
  s.logdet = o; o += vvv ? nl : 0;
  o = (o + 1) & ~1;   // 16 B alignment of the tile
  s.tile = o; o += 2 * DP * kTileLd;   // two buffers: cp.async landing zone of y, restaged in place as sqrt(w) y
  s.ct = o; o += (kBlock / 32) * NFRAG * 64;   // DMMA output tiles, warp-private, resident across the segment
  s.sws = 0;
  s.wvs = o; o += kBlock;
  s.acc = o; o += E;
  const bool scalar_sums = vvv_pairs || q > 16;
  s.P = scalar_sums ? cluster_sum_parts(D, q) : 0;
  s.accP = o; o += s.P * q * (D + 2);
  s.red = o; o += 3 * (kBlock / 32);
  s.ybuf = s.tile;
  s.ellbuf = o; o += (2 * maxdeg * kBlock + 1) / 2;       // neighbour ids (int32),
  s.origbuf = o; o += kBlock;                             // original spot ids (int32, 2 buffers)
  s.zs_bytes_off = o * 8;
  s.bytes = (size_t) o * 8 + kBlock;
  return s;
}

template <int D, bool TDIST, bool VVV>
__global__ void __launch_bounds__(kBlock) k_sweep(const SweepArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NP = D * (D + 1) / 2, DP = TileGeom<D>::DP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = a.PL.q, nl = a.PL.nl, E = a.SL.E;
  const SweepSmem L = sweep_smem_layout(D, q, nl, NP, E, TDIST, VVV, a.SL.nlT > 1, a.maxdeg);
  double *s_g = sm + L.g, *s_c = sm + L.c, *s_gL = sm + L.gL, *s_cL = sm + L.cL, *s_PL = sm + L.PL;
  double *s_PM = sm + L.PM, *s_logdet = sm + L.logdet, *tile0 = sm + L.tile;
  double *wvs = sm + L.wvs, *acc = sm + L.acc, *accP = sm + L.accP, *ct = sm + L.ct, *red = sm + L.red;
  uint8_t *zs = reinterpret_cast<uint8_t *>(sm) + L.zs_bytes_off;
  const int P = L.P;
  int32_t *ellbuf = reinterpret_cast<int32_t *>(sm + L.ellbuf), *origbuf = reinterpret_cast<int32_t *>(sm + L.origbuf);

  const bool stats_only = a.flags & SW_STATS_ONLY, dry = a.flags & SW_DRY;
  if (*a.err) return;
  if (!stats_only) {
    const double *Pm = a.params;
    for (int e = tid; e < q * D; e += kBlock) s_g[e] = Pm[a.PL.g + e];
    for (int e = tid; e < q; e += kBlock) s_c[e] = Pm[a.PL.c + e];
    if (TDIST) {
      for (int e = tid; e < q * D; e += kBlock) s_gL[e] = Pm[a.PL.gL + e];
      for (int e = tid; e < q; e += kBlock) s_cL[e] = Pm[a.PL.cL + e];
      for (int e = tid; e < nl * NP; e += kBlock) s_PL[e] = Pm[a.PL.PL + e];
    }
    if (VVV) {
      for (int e = tid; e < nl * NP; e += kBlock) s_PM[e] = Pm[a.PL.PM + e];
      for (int e = tid; e < nl; e += kBlock) s_logdet[e] = Pm[a.PL.logdet + e];
    }
  }
  for (int e = tid; e < (DP - D - 1) * kTileLd; e += kBlock)   // padding rows of both buffers
    tile0[(D + 1) * kTileLd + e] = tile0[(DP + D + 1) * kTileLd + e] = 0.0;
  const Key key{a.key0, a.key1};
  const uint32_t it = *a.iter + a.iter_add;
  const double w_alpha = (double) ((D + 4) / 2);   // quirk Q2: integer division (src/cluster.cpp:508)
  const double halfD = 0.5 * D;
  const bool do_pairs = !(a.flags & SW_NO_PAIRS) && a.SL.nlT > 0;
  const bool per_cluster_pairs = a.SL.nlT > 1;
  const bool scalar_sums = per_cluster_pairs || q > 16;   // VVV samplers / many clusters: scalar reductions
  const int nh = scalar_sums ? 0 : (q + 7) / 8;
  __syncthreads();

  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int begin = a.seg_begin[seg], count = a.seg_count[seg];
    for (int e = tid; e < E; e += kBlock) acc[e] = 0.0;
    for (int e = tid; e < P * q * (D + 2); e += kBlock) accP[e] = 0.0;
    for (int e = tid; e < (kBlock / 32) * stats_frags(D, nh) * 64; e += kBlock) ct[e] = 0.0;
    double hp_acc = 0.0, nacc = 0.0;
    LogProd lw;
    // Each thread streams the inputs of ITS spot of the next tile into shared memory with cp.async while
    // the current tile is being processed (it is the only reader of what it copied: no barrier needed).
    auto prefetch = [&](int buf, int pbase) {
      if (pbase + tid < count) {
        const int64_t pj = (int64_t) begin + pbase + tid;
        const double *src = a.Y + pj;
        double *dst = tile0 + buf * DP * kTileLd + tid;
#pragma unroll
        for (int c = 0; c < D; c++, src += a.ld, dst += kTileLd) cp_async8(dst, src);
        const int32_t *esrc = a.ell + pj;
        int32_t *edst = ellbuf + buf * a.maxdeg * kBlock + tid;
#pragma unroll 4
        for (int e = 0; e < a.maxdeg; e++, esrc += a.ld, edst += kBlock) cp_async4(edst, esrc);
        cp_async4(origbuf + buf * kBlock + tid, a.orig + pj);
      }
      cp_async_commit();
    };
    prefetch(0, 0);

    for (int base = 0, tix = 0; base < count; base += kBlock, tix++) {
      const int cur = tix & 1;
      double *tile = tile0 + cur * DP * kTileLd, *sws = tile + D * kTileLd;
      const bool active = base + tid < count;
      const int64_t j = (int64_t) begin + base + tid;
      if (base + kBlock < count) { prefetch(cur ^ 1, base + kBlock); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      double y[D];
      double wj = 1.0;
      int zfinal = 255;
      if (active) {
#pragma unroll
        for (int c = 0; c < D; c++) y[c] = tile[c * kTileLd + tid];
        const int zk = a.z[j];
        zfinal = zk;
        if (stats_only) {
          if (TDIST) wj = a.w[j];
        } else {
          const uint32_t og = (uint32_t) origbuf[cur * kBlock + tid];
          uint32_t prop_bits;
          double u_acc, u_sq;
          rng_spot(key, og, it, prop_bits, u_acc, u_sq);   // one Philox block: proposal, accept, first gamma attempt
          const int zn = propose_label_bits(zk, q, prop_bits);
          int deg = 0, cp = 0, cn = 0;
          for (int e = 0; e < a.maxdeg; e++) {
            const int32_t nb = ellbuf[(cur * a.maxdeg + e) * kBlock + tid];
            if (nb >= 0) {
              deg++;
              const int zu = a.z[nb];
              cn += (zu == zn);
              if (a.flags & SW_Q3) cp += (a.orig[nb] == zk + 1);   // quirk Q3 (:404)
              else cp += (zu == zk);
            }
          }
          double pot_p = 0.0, pot_n = 0.0;
          if (deg) {
            pot_p = a.gamma / deg * 2 * cp;
            pot_n = a.gamma / deg * 2 * cn;
          }
          if (TDIST) {
            // w_j ~ Gamma((d+4)/2, scale 2 / ((y-mu)^T Lambda (y-mu) + 4))   (:513-518, :655-660)
            const double aL = quadform<D>(y, s_PL + (VVV ? zk * NP : 0));
            const double bL = dotv<D>(y, s_gL + zk * D);
            const double quad = aL - 2.0 * bL + s_cL[zk];
            wj = rng_gamma_spot(key, og, it, w_alpha, 2.0 / (quad + 4.0), u_sq);
          }
          const double lin_p = s_c[zk] - 2.0 * dotv<D>(y, s_g + zk * D);
          const double lin_n = s_c[zn] - 2.0 * dotv<D>(y, s_g + zn * D);
          // the (d/2) log w_j term is common to both labels: it only enters plogLik (via lw)
          double delta, hp_part, aP = 0.0, aN = 0.0, ldp = 0.0, ldn = 0.0;
          if (VVV) {
            aP = quadform<D>(y, s_PM + zk * NP);
            aN = quadform<D>(y, s_PM + zn * NP);
            ldp = s_logdet[zk];
            ldn = s_logdet[zn];
            const double hp = pot_p + (ldp - 0.5 * wj * (aP + lin_p));
            const double hn = pot_n + (ldn - 0.5 * wj * (aN + lin_n));
            delta = hn - hp;
            hp_part = hp;
          } else {
            delta = (pot_n - pot_p) - 0.5 * wj * (lin_n - lin_p);
            hp_part = pot_p - 0.5 * wj * lin_p;
          }
          double prob = exp(delta);
          if (prob > 1.0) prob = 1.0;
          if (prob != prob) atomicExch(a.err, BSB_DEVERR_NUMERIC);   // NA probability (Rcpp sample() stops)
          const bool accept = sample_two(prob, u_acc);
          if (dry) {
            const double LOG2PI = 1.8378770664093454835606594728112;
            const double logw = TDIST ? log(wj) : 0.0;
            if (!VVV) {
              aP = aN = quadform<D>(y, a.params + a.PL.PM);
              ldp = ldn = a.params[a.PL.logdet];
            }
            a.dry_znew[og] = zn + 1;
            a.dry_w[og] = wj;
            a.dry_hp[og] = pot_p + ((ldp - halfD * LOG2PI) + halfD * logw - 0.5 * wj * (aP + lin_p));
            a.dry_hn[og] = pot_n + ((ldn - halfD * LOG2PI) + halfD * logw - 0.5 * wj * (aN + lin_n));
            a.dry_prob[og] = prob;
            a.dry_acc[og] = accept ? 1 : 0;
          } else {
            if (accept) {
              a.z[j] = (uint8_t) zn;
              zfinal = zn;
              nacc += 1.0;
            }
            if (TDIST) {
              a.w[j] = wj;
              lw.mul(wj);
            }
            hp_acc += hp_part;
          }
        }
      }
      if (dry) continue;
      // ---- stage sqrt(w) y; fold the tile into the segment's statistics -------------------------
      {
        const double sw = active ? sqrt(wj) : 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) tile[c * kTileLd + tid] = active ? sw * y[c] : 0.0;
        sws[tid] = sw;
        wvs[tid] = active ? wj : 0.0;
        zs[tid] = (uint8_t) zfinal;
      }
      __syncthreads();
      if (do_pairs && per_cluster_pairs) tile_pairs_per_cluster<D>(acc + a.SL.T, tile, zs, tid);
      if (!per_cluster_pairs && (do_pairs || nh)) tile_stats_mma<D>(ct, tile, zs, do_pairs, nh, warp, lane);
      if (scalar_sums) tile_cluster_sums<D>(accP, P, q, tile, sws, wvs, zs, tid);
      __syncthreads();
    }
    if (dry) continue;
    // ---- fold warps / parts in a fixed order and publish the segment's record --------------------
    double lws = TDIST ? lw.value() : 0.0;
    block_reduce3(hp_acc, nacc, lws, red, tid);
    if (tid == 0) {
      acc[a.SL.hp] = hp_acc + halfD * lws;
      acc[a.SL.nacc] = nacc;
    }
    if (!per_cluster_pairs) tile_stats_fold<D>(ct, acc, a.SL, do_pairs, nh ? q : 0, nh, tid);
    if (scalar_sums) tile_cluster_fold<D>(accP, P, q, acc, a.SL, tid);
    __syncthreads();
    double *out = a.partials + (size_t) a.seg_slot[seg] * E;
    for (int e = tid; e < E; e += kBlock) out[e] = acc[e];
    __syncthreads();
  }
}

template <bool TDIST, bool VVV>
void launch_sweep_t(const SweepArgs &a, int grid, cudaStream_t s) {
  constexpr int D = BSB_D;
  const SweepSmem L = sweep_smem_layout(D, a.PL.q, a.PL.nl, D * (D + 1) / 2, a.SL.E, TDIST, VVV, a.SL.nlT > 1, a.maxdeg);
  static size_t granted[64] = {};
  if (ensure_dynamic_smem(k_sweep<D, TDIST, VVV>, L.bytes, granted) != cudaSuccess) return;   // the launch check reports it
  k_sweep<D, TDIST, VVV><<<grid, kBlock, L.bytes, s>>>(a);
}

void launch_sweep(int tdist, int vvv, const SweepArgs &a, int grid, cudaStream_t s) {
  if (tdist) { if (vvv) launch_sweep_t<true, true>(a, grid, s); else launch_sweep_t<true, false>(a, grid, s); }
  else { if (vvv) launch_sweep_t<false, true>(a, grid, s); else launch_sweep_t<false, false>(a, grid, s); }
}

size_t sweep_smem(int tdist, int vvv, const SweepArgs &a) {
  return sweep_smem_layout(BSB_D, a.PL.q, a.PL.nl, BSB_D * (BSB_D + 1) / 2, a.SL.E, tdist, vvv, a.SL.nlT > 1, a.maxdeg).bytes;
}

int sweep_max_blocks(int tdist, int vvv, size_t smem) {
  int nb = 0;
  constexpr int D = BSB_D;
  if (tdist) {
    if (vvv) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sweep<D, true, true>, kBlock, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sweep<D, true, false>, kBlock, smem);
  } else {
    if (vvv) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sweep<D, false, true>, kBlock, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sweep<D, false, false>, kBlock, smem);
  }
  return nb;
}

// ---- K3: spot x cluster log-density probe (dmvnrm_arma_fast / dmvnrm_prec_arma_fast applied to
// every (spot, cluster) pair), direct triangular form as at src/cluster.cpp:96-101, :122-127 ------
template <int D>
__global__ void __launch_bounds__(kBlock) k_loglik(const LoglikArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int q = a.q, nl = a.nl;
  double *s_R = sm, *s_mu = s_R + nl * D * D, *s_ld = s_mu + q * D;
  for (int e = tid; e < nl * D * D; e += kBlock) s_R[e] = a.rooti[e];
  for (int e = tid; e < q * D; e += kBlock) s_mu[e] = a.mu[e];
  for (int e = tid; e < nl; e += kBlock) s_ld[e] = a.logdet[e];
  __syncthreads();
  for (int64_t j = (int64_t) blockIdx.x * kBlock + tid; j < a.n; j += (int64_t) gridDim.x * kBlock) {
    double y[D];
#pragma unroll
    for (int c = 0; c < D; c++) y[c] = a.Y[(size_t) c * a.n + j];
    const double wj = a.w ? a.w[j] : 1.0;
    const double lw = a.w ? 0.5 * D * log(wj) : 0.0;
    for (int k = 0; k < q; k++) {
      const double *R = s_R + (nl > 1 ? k : 0) * D * D;
      double x[D];
#pragma unroll
      for (int c = 0; c < D; c++) x[c] = y[c] - s_mu[k * D + c];
      double dot = 0.0;
      if (!a.transposed) {
#pragma unroll
        for (int i = 0; i < D; i++) {   // inplace_tri_mat_mult_t (src/cluster.cpp:68-78)
          double t = 0.0;
#pragma unroll
          for (int jj = i; jj < D; jj++) t += R[jj * D + i] * x[jj];
          dot += t * t;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < D; jj++) {
          double t = 0.0;
#pragma unroll
          for (int i = 0; i <= jj; i++) t += x[i] * R[jj * D + i];
          dot += t * t;
        }
      }
      a.out[(size_t) k * a.n + j] =
          (s_ld[nl > 1 ? k : 0] + lw - 0.5 * D * 1.8378770664093454835606594728112) - 0.5 * wj * dot;
    }
  }
}

void launch_loglik(const LoglikArgs &a, cudaStream_t s) {
  constexpr int D = BSB_D;
  const size_t smem = sizeof(double) * ((size_t) a.nl * D * D + (size_t) a.q * D + a.nl);
  cudaFuncSetAttribute(k_loglik<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const int grid = (int) std::min<int64_t>((a.n + kBlock - 1) / kBlock, 148 * 8);
  k_loglik<D><<<grid > 0 ? grid : 1, kBlock, smem, s>>>(a);
}

KernelRegistrar reg(BSB_D, KernelsForD{launch_sweep, sweep_smem, sweep_max_blocks, launch_loglik, nullptr, nullptr});

}   // namespace
}   // namespace bsb
