// deconv_kernels.cu -- k_deconv<D, ADAPT>: one sweep of iterate_deconv (spatialEnhance) for the
// parents of ONE parent-colour class, instantiated for d = BSB_D.
//
// Reference loop replaced: /root/reference/src/cluster.cpp:904-1056
//   :905-908   jitter draws                    :913-962  per-parent proposal and log-ratio
//   :969-982   joint accept of the S subspots  :987-995  Robust Adaptive Metropolis update
//   :998-1008  t-model weights                 :1010-1051 label MH per subspot
//   :1055-1056 Ychange / plogLik
// Mapping: one lane per subspot, the S subspots of a parent on S adjacent lanes of one warp, so
// the zero-sum constraint and the joint accept stay inside a warp (shuffles), the S label updates
// of a parent run sequentially (r = 0..S-1, live neighbour labels) exactly as in the reference,
// and parents of one colour class -- never adjacent -- run in parallel.  The same pass
// accumulates the sufficient statistics of the next iteration, so Y is read and written once.
#include "deconv.hpp"
#include "kernels.hpp"
#include "peer_sync.cuh"
#include "rng.cuh"
#include "tile_stats.cuh"

#ifndef BSB_D
#error "compile with -DBSB_D=<d>"
#endif

namespace bsb {
namespace {

struct DeconvSmem {
  int gL, cL, PL, tile, sws, wvs, acc, accP, ct, red, logtab, zs_bytes_off, P;
  size_t bytes;
};
__host__ __device__ inline DeconvSmem deconv_smem_layout(int D, int q, int np, int E) {
  DeconvSmem s;
  const int DP = (D + 1 + 7) / 8 * 8, NFRAG = stats_frags(D, q > 16 ? 0 : (q + 7) / 8);
  int o = 0;
  s.gL = o; o += q * D;
  s.cL = o; o += q;
  s.PL = o; o += np;
  o = (o + 1) & ~1;
  s.tile = o; o += DP * kTileLd;
  s.ct = o; o += (kBlock / 32) * NFRAG * 64;   // DMMA output tiles, warp-private, resident across the segment
  s.sws = s.tile + D * kTileLd;   // sqrt(w) is row D of the tile
  s.wvs = o; o += kBlock;
  s.acc = o; o += E;
  s.P = q > 16 ? cluster_sum_parts(D, q) : 0;
  s.accP = o; o += s.P * q * (D + 2);
  s.red = o; o += 3 * (kBlock / 32);
  o = (o + 1) & ~1;
  s.logtab = o; o += 256;   // 128 (log, reciprocal) pairs of the table logarithm (rng.cuh)
  s.zs_bytes_off = o * 8;
  s.bytes = (size_t) o * 8 + kBlock;
  return s;
}

// subspot position nb = rr * ld0 + ipn -> (rr, ipn) without the integer division: rr < S <= 32, so the single-precision
// quotient is within one of the true one.
__device__ __forceinline__ void split_subspot(int32_t nb, int32_t ld0, float inv_ld0, int32_t &rr, int32_t &ipn) {
  rr = (int32_t) ((float) nb * inv_ld0);
  ipn = nb - rr * ld0;
  if (ipn < 0) { rr--; ipn += ld0; }
  else if (ipn >= ld0) { rr++; ipn -= ld0; }
}

template <int D, bool ADAPT>
__global__ void __launch_bounds__(kBlock) k_deconv(const DeconvArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int NP = D * (D + 1) / 2, DP = TileGeom<D>::DP, NTILES = TileGeom<D>::NTILES, NT = TileGeom<D>::NT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = a.PL.q, E = a.SL.E, S = a.S;
  const DeconvSmem L = deconv_smem_layout(D, q, NP, E);
  double *s_gL = sm + L.gL, *s_cL = sm + L.cL, *s_PL = sm + L.PL, *tile = sm + L.tile, *sws = sm + L.sws;
  double *wvs = sm + L.wvs, *acc = sm + L.acc, *accP = sm + L.accP, *ct = sm + L.ct, *red = sm + L.red;
  uint8_t *zs = reinterpret_cast<uint8_t *>(sm) + L.zs_bytes_off;
  double2 *s_logtab = reinterpret_cast<double2 *>(sm + L.logtab);
  const int P = L.P;
  if (*a.err) return;
  const bool scalar_sums = q > 16;
  const int nh = scalar_sums ? 0 : (q + 7) / 8;
  for (int e = tid; e < (DP - D - 1) * kTileLd; e += kBlock) tile[(D + 1) * kTileLd + e] = 0.0;   // padding rows
  if (!a.stats_only) {
    for (int e = tid; e < 128; e += kBlock) log_table_entry(e, s_logtab[e]);
    const double *P = a.params;
#pragma unroll 1
    for (int e = tid; e < q * D; e += kBlock) s_gL[e] = P[a.PL.gL + e];
#pragma unroll 1
    for (int e = tid; e < q; e += kBlock) s_cL[e] = P[a.PL.cL + e];
#pragma unroll 1
    for (int e = tid; e < NP; e += kBlock) s_PL[e] = P[a.PL.PL + e];
  }
  const Key key{a.key0, a.key1};
  const uint32_t it = *a.iter;
  const double w_alpha = (double) ((D + 4) / 2);   // quirk Q2 (src/cluster.cpp:809)
  const double halfD = 0.5 * D;
  const int gpw = 32 / S;                          // parents per warp
  const int ppb = (kBlock / 32) * gpw;             // parents per CTA pass
  const int g = lane / S, r = lane - g * S;
  const bool lane_valid = g < gpw;
  const int glane0 = g * S;                        // first lane of this parent's group
  const double inv_S = 1.0 / S;
  const float inv_ld0 = 1.0f / (float) a.ld0;
  __syncthreads();

  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int begin = a.seg_begin[seg], count = a.seg_count[seg];
    int my_colour = 0;
    if (a.fuse)
      while (my_colour + 1 < a.ncol && seg >= a.colour_seg_start[my_colour + 1]) my_colour++;
#pragma unroll 1
    for (int e = tid; e < E; e += kBlock) acc[e] = 0.0;
    for (int e = tid; e < P * q * (D + 2); e += kBlock) accP[e] = 0.0;
    for (int e = tid; e < (kBlock / 32) * stats_frags(D, nh) * 64; e += kBlock) ct[e] = 0.0;
    double hp_acc = 0.0, nacc_y = 0.0;
    LogProd lw;

    for (int base = 0; base < count; base += ppb) {
      const int pidx = base + warp * gpw + g;
      const bool active = lane_valid && pidx < count;
      const int64_t ip = (int64_t) begin + pidx;
      const int64_t si = (int64_t) r * a.ld0 + ip;
      double y[D], e[D];
      double wj = 1.0;
      int zk = 0, zfinal = 255;
      uint32_t j0 = 0, sorig = 0;
#pragma unroll
      for (int c = 0; c < D; c++) { y[c] = 0.0; e[c] = 0.0; }
      if (active) {
#pragma unroll
        for (int c = 0; c < D; c++) y[c] = a.Y[((size_t) c * S + r) * a.ld0 + ip];
        zk = a.z[si];
        zfinal = zk;
        if (a.tdist) wj = a.w[si];
        j0 = (uint32_t) a.orig0[ip];
        sorig = (uint32_t) r * (uint32_t) a.n0 + j0;
      }
      double dl_z = 0.0, u_acc_z = 0.0, hp_stay = 0.0, hp_move = 0.0;   // inputs of the label scan
      int zn = 0, nacc_new = 0;
      if (!a.stats_only) {   // uniform branch
        double v[ADAPT ? D : 1];
#pragma unroll
        for (int c = 0; c < (ADAPT ? D : 1); c++) v[c] = 0.0;
        // ---- jitter: rmvnorm(n, 0, I * jitter_scale) (:905-908), transformed by the RAM factor (:922-929)
        if (active) {
          // One copy of the Box-Muller code instead of (D + 1) / 2: the loop is rolled and hands its pairs over through
          // this thread's column of the staging tile, which is free until sqrt(w) y is staged below (row D, hit by the
          // last pair when D is odd, is rewritten there).  The kernel runs a handful of warps per SM once per colour
          // class: its time is instruction fetch as much as arithmetic, so code size is what counts here.
#pragma unroll 1
          for (int c2 = 0; c2 < (D + 1) / 2; c2++) {
            double n0v, n1v;
            rng_normal2_tab(key, P_ERR, sorig, it, (uint32_t) c2, s_logtab, n0v, n1v);
            tile[(2 * c2) * kTileLd + tid] = a.err_sd * n0v;
            tile[(2 * c2 + 1) * kTileLd + tid] = a.err_sd * n1v;
          }
#pragma unroll
          for (int c = 0; c < D; c++) e[c] = tile[c * kTileLd + tid];
          if (ADAPT) {
#pragma unroll
            for (int i = 0; i < D; i++) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j <= i; j++) s += a.A[((size_t) (i * (i + 1) / 2 + j) * S + r) * a.ld0 + ip] * e[j];
              v[i] = s;
            }
          }
        }
        double usq = 0.0;   // |u|^2 of the raw draw, needed by the RAM update
        if (ADAPT) {
#pragma unroll
          for (int c = 0; c < D; c++) { usq += e[c] * e[c]; e[c] = v[c]; }
        }
        // ---- zero-sum constraint: subtract the mean over the parent's S subspots (:932-935)
        {
          double tot[D];
#pragma unroll
          for (int c = 0; c < D; c++) tot[c] = 0.0;
#pragma unroll 1
          for (int rr = 0; rr < S; rr++) {
            const int from = glane0 + rr;
#pragma unroll
            for (int c = 0; c < D; c++) tot[c] += __shfl_sync(0xffffffffu, e[c], from);
          }
#pragma unroll
          for (int c = 0; c < D; c++) e[c] = fma(-tot[c], inv_S, e[c]);
        }
        // ---- log-ratio of the joint proposal (:941-956)
        double dl = 0.0, quad_old = 0.0, quad_new = 0.0, b_old = 0.0, b_new = 0.0;
        if (active) {
          double pen_old = 0.0, pen_new = 0.0;
#pragma unroll
          for (int c = 0; c < D; c++) {
            const double y0 = a.Y0[(size_t) c * a.ld0 + ip];
            pen_old += (y[c] - y0) * (y[c] - y0);
            e[c] = y[c] + e[c];   // e now holds the proposed Y
            pen_new += (e[c] - y0) * (e[c] - y0);
          }
          b_old = dotv<D>(y, s_gL + zk * D);
          b_new = dotv<D>(e, s_gL + zk * D);
          quad_old = quadform<D>(y, s_PL) - 2.0 * b_old + s_cL[zk];
          quad_new = quadform<D>(e, s_PL) - 2.0 * b_new + s_cL[zk];
          dl = -0.5 * wj * (quad_new - quad_old) - a.c * (pen_new - pen_old);
        }
        double tot = 0.0;
#pragma unroll 1
        for (int rr = 0; rr < S; rr++) tot += __shfl_sync(0xffffffffu, dl, glane0 + rr);
        bool accY = false;
        if (active) {
          double probY = exp(tot);
          if (probY > 1.0) probY = 1.0;
          if (probY != probY) atomicExch(a.err, BSB_DEVERR_NUMERIC);
          double u_y, u_unused;
          rng_uniform2(key, P_SPOT_YACC, j0, it, 0u, u_y, u_unused);
          accY = sample_two(probY, u_y);   // one draw per parent (:975)
          if (accY) {
#pragma unroll
            for (int c = 0; c < D; c++) {
              y[c] = e[c];
              a.Y[((size_t) c * S + r) * a.ld0 + ip] = e[c];
            }
            quad_old = quad_new;
            b_old = b_new;
            if (r == 0) nacc_y += 1.0;
          }
          // ---- Robust Adaptive Metropolis factor update (adaptive_mcmc, :49-65, :987-995)
          if (ADAPT) {
            nacc_new = a.nacc[ip] + (accY ? 1 : 0);
            const bool do_adapt = it > 10u && (a.adapt_before == 0 || it <= (uint32_t) a.adapt_before);
            if (do_adapt) {
              const double rate = (double) nacc_new / (double) it;   // cumulative acceptance (quirk Q11)
              const double step = fmin(1.0, D * pow((double) it, -2.0 / 3));
              const double kappa = step * (rate - 0.234) / usq;
              const double sk = sqrt(fabs(kappa));
              const bool up = kappa >= 0.0;
              double x[D];
#pragma unroll
              for (int c = 0; c < D; c++) x[c] = sk * v[c];
#pragma unroll
              for (int k = 0; k < D; k++) {
                double *Lkk = a.A + ((size_t) (k * (k + 1) / 2 + k) * S + r) * a.ld0 + ip;
                const double lkk = *Lkk;
                const double r2 = up ? lkk * lkk + x[k] * x[k] : lkk * lkk - x[k] * x[k];
                if (!(r2 > 0.0)) atomicExch(a.err, BSB_DEVERR_NUMERIC);
                const double rr = sqrt(r2), cc = rr / lkk, ss = x[k] / lkk;
                *Lkk = rr;
#pragma unroll
                for (int i = k + 1; i < D; i++) {
                  double *Lik = a.A + ((size_t) (i * (i + 1) / 2 + k) * S + r) * a.ld0 + ip;
                  const double lik = up ? (*Lik + ss * x[i]) / cc : (*Lik - ss * x[i]) / cc;
                  x[i] = cc * x[i] - ss * lik;
                  *Lik = lik;
                }
              }
            }
          }
          // ---- t-model weight (:998-1008)
          if (a.tdist) {
            wj = rng_gamma(key, P_SPOT_GAMMA, sorig, it, w_alpha, 2.0 / (quad_old + 4.0));
            a.w[si] = wj;
            lw.mul(wj);   // (d/2) log w_j enters plogLik only: accumulated as a product, one log per segment
          }
          // ---- label proposal; the density part of the MH ratio is known before the scan (:1011-1027)
          double u_prop;
          rng_uniform2(key, P_SPOT_U, sorig, it, 0u, u_prop, u_acc_z);
          zn = propose_label(zk, q, u_prop);
          const double lin_p = s_cL[zk] - 2.0 * b_old;
          const double lin_n = s_cL[zn] - 2.0 * dotv<D>(y, s_gL + zn * D);
          dl_z = -0.5 * wj * (lin_n - lin_p);
          hp_stay = -0.5 * wj * lin_p;   // per-spot part of plogLik if the label stays
          hp_move = -0.5 * wj * lin_n;   // ... if it moves (:1035, :1048)
        }
        if (ADAPT && active && r == 0) a.nacc[ip] = nacc_new;
      }
      // ---- stage sqrt(w) y; the second moments do not depend on the labels: they are folded BEFORE the label scan,
      //      i.e. (one launch for all colour classes) while this CTA would otherwise wait for the previous class
      {
        const double sw = active ? sqrt(wj) : 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) tile[c * kTileLd + tid] = active ? sw * y[c] : 0.0;
        sws[tid] = sw;
        wvs[tid] = active ? wj : 0.0;
        zs[tid] = 255;
      }
      __syncthreads();
      tile_stats_mma<D>(ct, tile, zs, true, 0, warp, lane, nh);
      if (!a.stats_only) {
        // ---- sequential label updates inside each parent: r = 0..S-1, neighbour labels live (:984,:1029-1051)
        // Neighbours in OTHER parents belong to another colour class and do not change during this launch: their
        // labels are gathered once, before the turns.  Neighbours inside the parent sit on adjacent lanes: their live
        // label is the lane's register, fetched by shuffle at every turn -- no memory traffic inside the scan.
        // Everything of the scan that does not read a label happens BEFORE the wait for the previous colour class (one
        // launch for all classes): the neighbour ids, which of them sit in the own parent, and the Potts scale.
        constexpr int kNbReg = 8;
        int32_t nbv[kNbReg];
        int src[kNbReg];
        int deg = 0;
#pragma unroll
        for (int ee = 0; ee < kNbReg; ee++) {
          nbv[ee] = -1;
          src[ee] = lane;
          if (active && ee < a.maxdeg) {
            const int32_t nb = a.ell[((size_t) ee * S + r) * a.ld0 + ip];
            if (nb >= 0) {
              deg++;
              int32_t rr, ipn;
              split_subspot(nb, (int32_t) a.ld0, inv_ld0, rr, ipn);
              if (ipn == (int32_t) ip) src[ee] = glane0 + rr;   // same parent: live label of that lane
              else nbv[ee] = nb;                                  // another parent: gathered after the wait
            }
          }
        }
        // lanes of the own parent this subspot reads at every turn, compacted to the front (static indices only); the
        // turns shuffle nsp_max labels instead of kNbReg
        int sps[kNbReg], nsp = 0;
#pragma unroll
        for (int ee = 0; ee < kNbReg; ee++) sps[ee] = lane;
#pragma unroll
        for (int ee = 0; ee < kNbReg; ee++) {
          const bool same = src[ee] != lane;
#pragma unroll
          for (int k = 0; k < kNbReg; k++)
            if (same && k == nsp) sps[k] = src[ee];
          nsp += same;
        }
        const int nsp_max = __reduce_max_sync(0xffffffffu, nsp);
        if (a.fuse && my_colour > 0) {   // labels of the previous colour classes must be final and visible
          if (tid == 0) {
            const uint32_t target = it * (uint32_t) (a.colour_seg_start[my_colour] - a.colour_seg_start[my_colour - 1]);
            const uint32_t *flag = a.done + (my_colour - 1);
            const unsigned long long t0 = global_ns();
            uint32_t v;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
              if ((int32_t) (v - target) >= 0 || *(volatile int *) a.err) break;
              __nanosleep(20);
              if (global_ns() - t0 > 20ull * 1000 * 1000 * 1000) { atomicCAS(a.err, 0, BSB_DEVERR_PEER_TIMEOUT); break; }
            } while (true);
          }
          __syncthreads();
        }
        int zu[kNbReg];
#pragma unroll
        for (int ee = 0; ee < kNbReg; ee++)
          zu[ee] = nbv[ee] >= 0 ? (a.fuse ? (int) __ldcg(a.z + nbv[ee]) : (int) a.z[nbv[ee]]) : -1;
        int deg_more = 0, cp_more = 0, cn_more = 0;   // neighbours beyond kNbReg (general graphs): other parents only counted here
        bool more_same_parent = false;
        if (active)
          for (int ee = kNbReg; ee < a.maxdeg; ee++) {
            const int32_t nb = a.ell[((size_t) ee * S + r) * a.ld0 + ip];
            if (nb >= 0) {
              int32_t rr, ipn;
              split_subspot(nb, (int32_t) a.ld0, inv_ld0, rr, ipn);
              if (ipn == (int32_t) ip) { more_same_parent = true; continue; }
              deg_more++;
              const int z2 = a.fuse ? (int) __ldcg(a.z + nb) : (int) a.z[nb];
              cp_more += z2 == zk;
              cn_more += z2 == zn;
            }
          }
        double pscale = (deg + deg_more) ? a.gamma / (deg + deg_more) * 2 : 0.0;
        int cp_fix = cp_more, cn_fix = cn_more;   // neighbours in other parents: fixed during this class's scan
#pragma unroll
        for (int ee = 0; ee < kNbReg; ee++) {
          cp_fix += zu[ee] == zk;   // -1 (no such neighbour, or one of the own parent) matches no label
          cn_fix += zu[ee] == zn;
        }
        for (int turn = 0; turn < S; turn++) {
          int cp = cp_fix, cn = cn_fix, dg = deg + deg_more;
#pragma unroll
          for (int k = 0; k < kNbReg; k++)
            if (k < nsp_max) {   // warp-uniform
              const int live = __shfl_sync(0xffffffffu, zfinal, sps[k]);
              cp += (k < nsp) & (live == zk);
              cn += (k < nsp) & (live == zn);
            }
          if (active && r == turn) {
            if (more_same_parent)   // more than kNbReg neighbours, some of them in the own parent: read those live
              for (int ee = kNbReg; ee < a.maxdeg; ee++) {
                const int32_t nb = a.ell[((size_t) ee * S + r) * a.ld0 + ip];
                if (nb >= 0 && nb % (int32_t) a.ld0 == (int32_t) ip) {
                  dg++;
                  const int z2 = *((volatile uint8_t *) (a.z + nb));
                  cp += z2 == zk;
                  cn += z2 == zn;
                }
              }
            double pot_p = 0.0, pot_n = 0.0;
            if (dg) {
              if (more_same_parent) pscale = a.gamma / dg * 2;
              pot_p = pscale * cp;
              pot_n = pscale * cn;
            }
            // accept with probability min(1, exp(delta)) (:1043-1046), decided in the log domain like the cluster samplers'
            // sweep (rng.cuh: mh_accept): no exponential inside the turns unless the decision is within its rounding
            const double delta = (pot_n - pot_p) + dl_z;
            if (delta != delta) atomicExch(a.err, BSB_DEVERR_NUMERIC);
            const bool upd = mh_accept(delta, u_acc_z);
            if (upd) {
              a.z[si] = (uint8_t) zn;
              zfinal = zn;
              hp_acc += hp_move;
            } else {
              hp_acc += hp_stay;
            }
            if (more_same_parent) __threadfence_block();
          }
          __syncwarp();
        }
      }
      // ---- per-cluster sums with the final labels ---------------------------------------------------
      zs[tid] = (uint8_t) (active ? zfinal : 255);
      if (a.fuse && !a.stats_only && base + ppb >= count) {
        // the labels of this segment are final (last pass): let the next colour class go before folding the statistics
        // (the barrier orders every thread's label stores before thread 0's fence, which is cumulative)
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          atomicAdd(a.done + my_colour, 1u);
        }
      } else {
        __syncthreads();
      }
      if (nh) tile_stats_mma<D>(ct, tile, zs, false, nh, warp, lane, nh);
      if (scalar_sums) tile_cluster_sums<D>(accP, P, q, tile, sws, wvs, zs, tid);
      __syncthreads();
    }
    double lws = lw.value();
    block_reduce3(hp_acc, nacc_y, lws, red, tid);
    if (tid == 0) { acc[a.SL.hp] = hp_acc + halfD * lws; acc[a.SL.nacc] = nacc_y; }
    tile_stats_fold<D>(ct, acc, a.SL, true, nh ? q : 0, nh, tid);
    if (scalar_sums) tile_cluster_fold<D>(accP, P, q, acc, a.SL, tid);
    __syncthreads();
    double *out = a.partials + (size_t) a.seg_slot[seg] * E;
#pragma unroll 1
    for (int e2 = tid; e2 < E; e2 += kBlock) out[e2] = acc[e2];
    __syncthreads();
  }
}

size_t deconv_smem(const DeconvArgs &a) {
  return deconv_smem_layout(BSB_D, a.PL.q, BSB_D * (BSB_D + 1) / 2, a.SL.E).bytes;
}

void launch_deconv(const DeconvArgs &a, int grid, cudaStream_t s) {
  constexpr int D = BSB_D;
  const size_t smem = deconv_smem(a);
  static size_t granted0[64] = {}, granted1[64] = {};
  if (a.adaptive) {
    if (ensure_dynamic_smem(k_deconv<D, true>, smem, granted1) != cudaSuccess) return;   // the launch check reports it
    k_deconv<D, true><<<grid, kBlock, smem, s>>>(a);
  } else {
    if (ensure_dynamic_smem(k_deconv<D, false>, smem, granted0) != cudaSuccess) return;
    k_deconv<D, false><<<grid, kBlock, smem, s>>>(a);
  }
}

KernelRegistrar reg(BSB_D, KernelsForD{nullptr, nullptr, nullptr, nullptr, launch_deconv, deconv_smem});

}   // namespace
}   // namespace bsb
