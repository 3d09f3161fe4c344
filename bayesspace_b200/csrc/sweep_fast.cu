// sweep_fast.cu -- k_sweep_fast<D, TDIST, NH>: the z / w sweep of the shared-precision samplers
//   iterate    /root/reference/src/cluster.cpp:272-307      iterate_t  :507-553
// for q <= 16 clusters (NH = ceil(q / 8) label tiles), one instantiation per dimension d = BSB_D.  It is the
// kernel BASELINE.json's C2 / C4 / C5 configurations run; per-cluster precisions (iterate_vvv, iterate_t_vvv)
// and q > 16 stay on k_sweep (sweep_kernels.cu).
//
// Same chromatic scan and the same arithmetic identities as k_sweep (factorise once per iteration, expanded
// quadratic form, statistics of the next iteration fused into the sweep).  What is different is the machine
// mapping, chosen from the round-1 ncu profile (the sweep was instruction-issue bound at 53 warp-instructions
// per spot-update, not memory bound):
//
//  * every WARP is an independent worker with a private two-stage pipeline: 32 spots x (d PC columns,
//    neighbour ids, original ids) arrive in its own shared-memory buffers through four TMA tensor copies
//    (cp.async.bulk.tensor, tma_maps.hpp) issued by one lane and are awaited on the warp's own mbarrier.  There is
//    no CTA barrier in the tile loop any more, and no per-thread cp.async address arithmetic;
//  * the tile is never restaged: the pooled second moments sum_j w_j y_j y_j^T run on the FP64 tensor pipe
//    (mma.sync.m8n8k4.f64) straight from the landing buffer as A = y, B = w . y (the weight is applied to the B
//    fragment), the per-cluster sums as A = y, B = w . [z = k]; each warp folds its own 32 spots and keeps the
//    output tiles in registers for the whole segment;
//  * one Philox block per spot for proposal / accept / first gamma attempt (rng_spot), the MH accept and the
//    second Marsaglia-Tsang test decided by a single-precision screen with an exact FP64 fallback (rng.cuh);
//  * parameters are read from shared memory as 16-byte broadcasts; the Potts scale 2 gamma / |N_j| is a table;
//    the neighbour loop is unrolled for the lattice degrees (4 and 6).
#include <cstdlib>
#include <type_traits>

#include "kernels.hpp"
#include "peer_sync.cuh"
#include "pipeline.cuh"
#include "rng.cuh"
#include "tile_stats.cuh"
#include "tma_maps.hpp"

#ifndef BSB_D
#error "compile with -DBSB_D=<d>"
#endif

namespace bsb {
namespace {

#ifndef BSB_FAST_WARPS
#define BSB_FAST_WARPS 4
#endif
constexpr int kWarps = BSB_FAST_WARPS;        // warps (independent workers) per CTA of k_sweep_fast
constexpr int kFBlock = 32 * kWarps;          // threads per CTA
constexpr int kRowB = 128;   // bytes of one row of a Y box (16 spots) and of an ell / orig row (32 spots)

// Shared-memory carve-up (byte offsets; the per-warp areas start at a 1024-byte boundary, which SWIZZLE_128B needs
// for its pattern to be a function of the row index alone).
struct FastSmem {
  int g, c, gL, cL, PL, pot, logtab, acc, red, mbar, shared_bytes, warp_stride, ybox, ell, orig, wst, zst, skacc;
  size_t bytes;
};
__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline FastSmem fast_smem_layout(int D, int q, int E, bool tdist, int nh, int maxdeg) {
  FastSmem s;
  const int RS = D + (D & 1) + 2, NP = D * (D + 1) / 2, DP = (D + 1 + 7) / 8 * 8;   // RS: row stride of g / gL
  int o = 0;
  s.g = o; o += q * RS * 8;
  s.gL = o; o += tdist ? q * RS * 8 : 0;
  s.PL = o; o += tdist ? align16(NP * 8) : 0;
  s.c = o; o += align16(q * 8);
  s.cL = o; o += tdist ? align16(q * 8) : 0;
  s.pot = o; o += align16((kMaxDeg + 1) * 8);
  s.logtab = o; o += tdist ? 128 * 16 : 0;
  s.acc = o; o += align16(E * 8);
  s.red = o; o += align16(3 * kWarps * 8);
  s.mbar = o; o += kWarps * 2 * 8;
  s.shared_bytes = (o + 1023) & ~1023;
  int w = 0;
  s.ybox = DP * kRowB;                 // one box: DP rows (d PCs, the constant-1 row, zero rows) of 16 spots
  w += 4 * s.ybox;                     // [buffer][half]; also holds the warp's output tiles at the end of a segment
  s.ell = w; w += 2 * maxdeg * kRowB;
  s.orig = w; w += 2 * kRowB;
  s.wst = w; w += 32 * 8;
  s.zst = w; w += 32;
  s.skacc = w; w += (D <= 16 ? q * DP * 8 : 0);   // per-cluster sums of the label-uniform sub-tiles (rows 0..d-1: s_k, row d: n_k)
  s.warp_stride = (w + 1023) & ~1023;
  s.bytes = (size_t) s.shared_bytes + (size_t) kWarps * s.warp_stride + 1024;   // + slack to align the base
  return s;
}

// policy: L2 eviction hint.  Everything a sweep streams (PCs, neighbour ids, original ids) is read exactly once per
// launch: evict-first keeps it from flushing what IS re-read -- the label field (neighbour gathers), the statistics
// records and the code of k_param, which runs between two sweeps.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

#ifdef BSB_FAST_MINB   // tuning experiments
constexpr int fast_min_blocks(int D) { return D <= 16 ? BSB_FAST_MINB : (D <= 24 ? BSB_FAST_MINB - 1 : 1); }
#else
constexpr int fast_min_blocks(int D) { return (D <= 16 ? 16 : (D <= 24 ? 12 : 8)) / kWarps; }   // 16 / 12 / 8 warps per SM
#endif
// The cluster sums of label-uniform sub-tiles come from the second moments' constant-row column (below) when the
// per-warp accumulators fit next to four resident CTAs; larger d keep the shared memory for occupancy instead.
__host__ __device__ constexpr bool fast_uniform_shortcut(int D) { return D <= 16; }

// x^T P x with P = packed upper triangle, COLUMN by column (b outer, a <= b inner, off-diagonals doubled), read as
// 16-byte broadcasts.  Consecutive FMAs of a column go to different accumulators s[a], so the FP64 latency of one
// is covered by the others (the row-wise order is a serial chain of d FMAs per row).
template <int D>
__device__ __forceinline__ double quadform_cols(const double (&y)[D], const double *__restrict__ P) {
  constexpr int NP = D * (D + 1) / 2;
  const double2 *P2 = reinterpret_cast<const double2 *>(P);
  double s[D];
#pragma unroll
  for (int a = 0; a < D; a++) s[a] = 0.0;
  int a = 0, b = 0;   // (a, b) of packed index p, advanced at compile time by the unroller
#pragma unroll
  for (int p2 = 0; p2 < (NP + 1) / 2; p2++) {
    const double2 v = P2[p2];
    s[a] = fma(v.x, y[b], s[a]);
    if (++a > b) { b++; a = 0; }
    if (2 * p2 + 1 < NP) {
      s[a] = fma(v.y, y[b], s[a]);
      if (++a > b) { b++; a = 0; }
    }
  }
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int r = 0; r < D; r += 2) {
    acc0 = fma(y[r], s[r], acc0);
    if (r + 1 < D) acc1 = fma(y[r + 1], s[r + 1], acc1);
  }
  return acc0 + acc1;
}
__host__ __device__ inline int colpacked_index(int a, int b) { return b * (b + 1) / 2 + a; }   // a <= b

// Three dot products of y at once (rows padded to an even length, 16-byte aligned): six independent FMA chains.
template <int D, bool WITH_L>
__device__ __forceinline__ void dot3(const double (&y)[D], const double *__restrict__ gl, const double *__restrict__ gp,
                                     const double *__restrict__ gn, double &dl, double &dp, double &dn) {
  const double2 *l2 = reinterpret_cast<const double2 *>(gl), *p2 = reinterpret_cast<const double2 *>(gp),
                *n2 = reinterpret_cast<const double2 *>(gn);
  double l0 = 0.0, l1 = 0.0, p0 = 0.0, p1 = 0.0, n0 = 0.0, n1 = 0.0;
#pragma unroll
  for (int c2 = 0; c2 < (D + 1) / 2; c2++) {
    const double2 vp = p2[c2], vn = n2[c2];
    p0 = fma(y[2 * c2], vp.x, p0);
    n0 = fma(y[2 * c2], vn.x, n0);
    if (WITH_L) {
      const double2 vl = l2[c2];
      l0 = fma(y[2 * c2], vl.x, l0);
      if (2 * c2 + 1 < D) l1 = fma(y[2 * c2 + 1], vl.y, l1);
    }
    if (2 * c2 + 1 < D) {
      p1 = fma(y[2 * c2 + 1], vp.y, p1);
      n1 = fma(y[2 * c2 + 1], vn.y, n1);
    }
  }
  dl = l0 + l1; dp = p0 + p1; dn = n0 + n1;
}

// Fixed-order CTA reduction of three per-thread scalars over kWarps warps; results valid on thread 0.
__device__ __forceinline__ void block_reduce3_fast(double &x, double &y, double &z, double *red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
    z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  if ((tid & 31) == 0) { red[3 * (tid >> 5)] = x; red[3 * (tid >> 5) + 1] = y; red[3 * (tid >> 5) + 2] = z; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int wv = 0; wv < kWarps; wv++) { a += red[3 * wv]; b += red[3 * wv + 1]; c += red[3 * wv + 2]; }
    x = a; y = b; z = c;
  }
}

constexpr int kNbInline = 6;   // neighbour slots handled in the straight-line part (hex lattice degree); the rest in a loop

template <int D, bool TDIST, int NH, bool LEAN, bool FUSEC>
__global__ void __launch_bounds__(kFBlock, fast_min_blocks(D))
    k_sweep_fast(const __grid_constant__ SweepArgs a, const __grid_constant__ SweepMaps maps) {
  extern __shared__ __align__(1024) unsigned char smdyn[];
  constexpr int NP = D * (D + 1) / 2, RS = D + (D & 1) + 2;   // RS = 2 (mod 16) doubles: rows of different clusters start in different banks
  constexpr int NT = TileGeom<D>::NT, NTILES = TileGeom<D>::NTILES, DP = TileGeom<D>::DP;
  constexpr int YBOX = DP * kRowB;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: TMA operands stay in uniform registers
  const int q = a.PL.q, E = a.SL.E;
  const FastSmem L = fast_smem_layout(D, q, E, TDIST, NH, a.maxdeg);
  unsigned char *smraw = smdyn + ((1024u - (smem_u32(smdyn) & 1023u)) & 1023u);   // 1024-byte aligned base
  double *s_g = reinterpret_cast<double *>(smraw + L.g), *s_c = reinterpret_cast<double *>(smraw + L.c);
  double *s_gL = reinterpret_cast<double *>(smraw + L.gL), *s_cL = reinterpret_cast<double *>(smraw + L.cL);
  double *s_PL = reinterpret_cast<double *>(smraw + L.PL), *s_pot = reinterpret_cast<double *>(smraw + L.pot);
  double2 *s_logtab = reinterpret_cast<double2 *>(smraw + L.logtab);
  double *acc = reinterpret_cast<double *>(smraw + L.acc), *red = reinterpret_cast<double *>(smraw + L.red);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smraw + L.mbar) + 2 * warp;
  unsigned char *wbase = smraw + L.shared_bytes + (size_t) warp * L.warp_stride;
  unsigned char *ybuf = wbase;   // [buffer][half] boxes of DP rows x 128 bytes, SWIZZLE_128B
  int32_t *well = reinterpret_cast<int32_t *>(wbase + L.ell), *worig = reinterpret_cast<int32_t *>(wbase + L.orig);
  double *wst = reinterpret_cast<double *>(wbase + L.wst);
  uint8_t *zst = wbase + L.zst;
  double *skacc = reinterpret_cast<double *>(wbase + L.skacc);
  double *cfold = reinterpret_cast<double *>(smraw + L.shared_bytes);   // + warp * warp_stride / 8
  const int fold_stride = L.warp_stride / 8;

  // LEAN: the instantiation of the sampling loop proper (no probe flags): the statistics-only and dry-run paths are
  // compiled out, which keeps the hot loop's code -- and its instruction-cache footprint -- small
  const bool stats_only = !LEAN && (a.flags & SW_STATS_ONLY), dry = !LEAN && (a.flags & SW_DRY);
  if (*a.err) return;
  if (!stats_only) {
    const double *Pm = a.params;
    for (int e = tid; e < q * RS; e += kFBlock) {
      const int k = e / RS, cc = e - k * RS;
      s_g[e] = cc < D ? Pm[a.PL.g + k * D + cc] : 0.0;
      if (TDIST) s_gL[e] = cc < D ? Pm[a.PL.gL + k * D + cc] : 0.0;
    }
    for (int e = tid; e < q; e += kFBlock) {
      s_c[e] = Pm[a.PL.c + e];
      if (TDIST) s_cL[e] = Pm[a.PL.cL + e];
    }
    if (TDIST) {   // packed Lambda, permuted from k_param's row-major order to column-major (quadform_cols)
      for (int e = tid; e < NP; e += kFBlock) {
        int pa, pb;
        pair_from_index(e, D, pa, pb);
        s_PL[colpacked_index(pa, pb)] = Pm[a.PL.PL + e];
      }
      if (tid == 0 && (NP & 1)) s_PL[NP] = 0.0;
    }
    for (int e = tid; e <= kMaxDeg; e += kFBlock) s_pot[e] = e ? a.gamma / e * 2 : 0.0;   // (gamma / |N_j|) * 2 as at :289
    if (TDIST)
      for (int e = tid; e < 128; e += kFBlock) log_table_entry(e, s_logtab[e]);
  }
  // constant rows of the four boxes: row D = 1 (its products give n_k and the sum of weights), rows beyond = 0
  auto init_const_rows = [&]() {
    for (int e = lane; e < 4 * (DP - D) * 16; e += 32) {
      const int box = e / ((DP - D) * 16), r = e % ((DP - D) * 16);
      reinterpret_cast<double *>(ybuf + box * YBOX + D * kRowB)[r] = r < 16 ? 1.0 : 0.0;
    }
  };
  init_const_rows();
  if (lane == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); }
  mbar_fence_init();
  fence_proxy_async();
  __syncthreads();

  const Key key{a.key0, a.key1};
  const uint32_t it = *a.iter + a.iter_add;
  constexpr double w_alpha = (double) ((D + 4) / 2);   // quirk Q2: integer division (src/cluster.cpp:508)
  const double halfD = 0.5 * D;
  const bool do_pairs = !(a.flags & SW_NO_PAIRS) && a.SL.nlT > 0;
  const int ell_rows = stats_only ? 0 : a.maxdeg;
  const uint32_t tx_bytes = (uint32_t) (2 * D * kRowB + (stats_only ? 0 : (ell_rows + 1) * kRowB));
  const uint64_t stream_policy = l2_evict_first_policy();
  uint32_t ntile = 0;   // sub-tiles this warp has consumed: buffer = ntile & 1, barrier parity = (ntile >> 1) & 1

  // Swizzled addressing (SWIZZLE_128B: 16-byte chunk index XOR row mod 8).
  //  per-spot reads: this lane's spot is column (lane & 15) of half (lane >> 4)
  const int my_half = (lane >> 4) * YBOX, my_chunk = ((lane & 15) >> 1) << 4, my_lo = (lane & 1) << 3;
  //  DMMA fragments: row 8 i + kc, spot 4 g + lq of the warp's 32
  const int kc = lane >> 2, lq = lane & 3;
  const int frag_row = kc * kRowB, frag_sw = ((lq >> 1) ^ kc) << 4, frag_lo = (lq & 1) << 3;

  __shared__ int s_last_boundary;
  for (int seg = blockIdx.x; seg < a.nseg; seg += gridDim.x) {
    const int begin = a.seg_begin[seg], count = a.seg_count[seg];
    const int nsub = (count + 31) >> 5;
    // fused halo exchange (HaloFuse, layout.hpp): a boundary segment first makes sure the neighbours' labels of the
    // previous colour phase have landed in this rank's ghost slots
    int my_colour = 0;
    if (FUSEC)
      while (my_colour + 1 < a.cfuse.ncol && seg >= a.cfuse.colour_seg_start[my_colour + 1]) my_colour++;
    const bool bseg = a.halo.nlegs > 0 && seg < a.halo.nbound && !stats_only;
    if (bseg) {
      if (tid == 0) {
        const uint32_t target = (it - 1u) * (uint32_t) a.halo.ncol + (dry ? 0u : (uint32_t) a.halo.colour);
        for (int l = 0; l < a.halo.nlegs; l++) wait_flag(a.halo.legs[l].my_flag, target, a.err, a.halo.timeout_ns);
      }
      __syncthreads();
    }
    for (int e = tid; e < E; e += kFBlock) acc[e] = 0.0;
    if (fast_uniform_shortcut(D))
      for (int e = lane; e < q * DP; e += 32) skacc[e] = 0.0;
    double hp_acc = 0.0, nacc = 0.0, lw_m = 1.0;
    int lw_e = 0;
    double C[NTILES][2], CH[NT][NH][2];
#pragma unroll
    for (int t = 0; t < NTILES; t++) C[t][0] = C[t][1] = 0.0;
#pragma unroll
    for (int i = 0; i < NT; i++)
#pragma unroll
      for (int h = 0; h < NH; h++) CH[i][h][0] = CH[i][h][1] = 0.0;

    auto issue = [&](int sub, int buf) {
      if (lane == 0) {
        const int gi = begin + 32 * sub;
        mbar_arrive_expect_tx(mbar + buf, tx_bytes);
        tma_load_2d(ybuf + (2 * buf) * YBOX, &maps.y, gi, 0, mbar + buf, stream_policy);
        tma_load_2d(ybuf + (2 * buf + 1) * YBOX, &maps.y, gi + 16, 0, mbar + buf, stream_policy);
        if (!stats_only) {
          if (ell_rows) tma_load_2d(well + buf * ell_rows * 32, &maps.ell, gi, 0, mbar + buf, stream_policy);
          tma_load_2d(worig + buf * 32, &maps.orig, gi, 0, mbar + buf, stream_policy);
        }
      }
    };
    if (warp < nsub) issue(warp, ntile & 1);

    for (int sub = warp; sub < nsub; sub += kWarps, ntile++) {
      const int cur = ntile & 1;
      if (sub + kWarps < nsub) issue(sub + kWarps, cur ^ 1);
      const int64_t j = (int64_t) begin + 32 * sub + lane;
      const bool active = 32 * sub + lane < count;
      const unsigned char *ycur = ybuf + 2 * cur * YBOX;
      // Every lane runs the straight-line code below; lanes beyond the segment's end sit on holes of the internal
      // order (zero PCs, no neighbours) and are masked out of every store and statistic.
      const int zk = a.z[j];
      mbar_wait(mbar + cur, (ntile >> 1) & 1);

      double wj = active ? 1.0 : 0.0;
      int zfinal = active ? zk : 255;
      if (stats_only) {
        if (TDIST && active) wj = a.w[j];
      } else {
        const uint32_t og = (uint32_t) worig[cur * 32 + lane];
        // neighbour labels: issue the gathers first, they are the longest latency of the spot
        // Padding entries (-1) read z[ld] = 255, a label no cluster has: the loads carry no select after them, so the
        // in-order issue does not stall on them before the independent work below has been issued
        int zu[kNbInline], deg = 0;
        if (!FUSEC) {
          const int32_t *ellrow = well + cur * ell_rows * 32 + lane;
#pragma unroll
          for (int e = 0; e < kNbInline; e++) {
            const int32_t nb = e < ell_rows ? ellrow[e * 32] : -1;
            deg += nb >= 0;
            zu[e] = a.z[nb >= 0 ? (int64_t) nb : a.ld];
          }
        }
        // randomness: depends on nothing but (spot, iteration)
        const uint4 oa = philox4x32_10(og, 0u, it, P_SPOT_U, key);
        const double u_acc = u53(oa.y, oa.z), u_sq = u32c(oa.w);
        double x0 = 0.0;
        if (TDIST) x0 = normal_from_block(philox4x32_10(og, 0u, it, P_SPOT_GAMMA, key), s_logtab);
        const int zn = propose_label_bits(zk, q, oa.x);
        double y[D];
        {
          const unsigned char *yb = ycur + my_half + my_lo;
#pragma unroll
          for (int c = 0; c < D; c++)
            y[c] = *reinterpret_cast<const double *>(yb + c * kRowB + (my_chunk ^ ((c & 7) << 4)));
        }
        double scale = 0.0, bL, dp, dn;
        dot3<D, TDIST>(y, s_gL + zk * RS, s_g + zk * RS, s_g + zn * RS, bL, dp, dn);
        if (TDIST) {
          // w_j ~ Gamma((d+4)/2, scale 2 / ((y-mu)^T Lambda (y-mu) + 4))   (:513-518)
          const double aL = quadform_cols<D>(y, s_PL);
          // 2 / (quadratic form + 4).  Measured on one box: the written-out reciprocal is 2.7 % faster at d = 20 and 2 %
          // slower at d = 15, where its longer live ranges add spills at 128 registers
          const double qf = aL - 2.0 * bL + s_cL[zk];
          scale = D > 16 ? rcp_moderate(fma(0.5, qf, 2.0)) : 2.0 / (qf + 4.0);
        }
        const double lin_p = s_c[zk] - 2.0 * dp;
        const double lin_n = s_c[zn] - 2.0 * dn;
        if (TDIST && FUSEC) wj = rng_gamma_spot(key, og, it, w_alpha, scale, x0, u_sq, s_logtab);
        if (FUSEC) {
          // one launch for all colour classes: everything above is independent of the neighbours; their labels are
          // final once every segment of the previous class has published (counter reset by k_param each iteration)
          if (my_colour > 0) {
            if (lane == 0) {
              const uint32_t target = (uint32_t) (a.cfuse.colour_seg_start[my_colour] - a.cfuse.colour_seg_start[my_colour - 1]);
              const uint32_t *flag = a.cfuse.done + (my_colour - 1);
              const unsigned long long t0 = global_ns();
              uint32_t v;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v >= target || *(volatile int *) a.err) break;
                __nanosleep(32);
                if (global_ns() - t0 > 20ull * 1000 * 1000 * 1000) { atomicCAS(a.err, 0, BSB_DEVERR_PEER_TIMEOUT); break; }
              } while (true);
            }
            __syncwarp();
          }
          const int32_t *ellrow = well + cur * ell_rows * 32 + lane;
#pragma unroll
          for (int e = 0; e < kNbInline; e++) {
            const int32_t nb = e < ell_rows ? ellrow[e * 32] : -1;
            deg += nb >= 0;
            zu[e] = __ldcg(a.z + (nb >= 0 ? (int64_t) nb : a.ld));
          }
        }
        int cp = 0, cn = 0;
#pragma unroll
        for (int e = 0; e < kNbInline; e++) {
          cp += zu[e] == zk;
          cn += zu[e] == zn;
        }
        for (int e = kNbInline; e < ell_rows; e++) {   // general graphs with more than 6 neighbours
          const int32_t nb = well[(cur * ell_rows + e) * 32 + lane];
          if (nb >= 0) {
            const int z2 = FUSEC ? (int) __ldcg(a.z + nb) : (int) a.z[nb];
            deg++;
            cp += z2 == zk;
            cn += z2 == zn;
          }
        }
        const double ps = s_pot[deg];
        const double pot_p = ps * cp, pot_n = ps * cn;
        if (TDIST && !FUSEC) wj = rng_gamma_spot(key, og, it, w_alpha, scale, x0, u_sq, s_logtab);
        // y^T M y and (d/2) log w_j are common to both labels: they only enter plogLik (statistics, lw)
        const double delta = (pot_n - pot_p) - 0.5 * wj * (lin_n - lin_p);
        const double hp_part = pot_p - 0.5 * wj * lin_p;
        const bool accept = mh_accept(delta, u_acc);
        if (active && delta != delta) atomicExch(a.err, BSB_DEVERR_NUMERIC);   // NA probability (Rcpp sample() stops)
        if (dry) {
          if (active) {
            const double LOG2PI = 1.8378770664093454835606594728112;
            const double logw = TDIST ? log(wj) : 0.0;
            const double aM = quadform<D>(y, a.params + a.PL.PM), ld = a.params[a.PL.logdet];
            double prob = exp(delta);
            if (prob > 1.0) prob = 1.0;
            a.dry_znew[og] = zn + 1;
            a.dry_w[og] = wj;
            a.dry_hp[og] = pot_p + ((ld - halfD * LOG2PI) + halfD * logw - 0.5 * wj * (aM + lin_p));
            a.dry_hn[og] = pot_n + ((ld - halfD * LOG2PI) + halfD * logw - 0.5 * wj * (aM + lin_n));
            a.dry_prob[og] = prob;
            a.dry_acc[og] = accept ? 1 : 0;
          }
        } else if (active) {
          if (accept) {
            a.z[j] = (uint8_t) zn;
            zfinal = zn;
            nacc += 1.0;
          }
          if (TDIST) {
            __stcs(a.w + j, wj);   // streaming store: written once per sweep, read only by snapshots
            // sum_j log w_j as a running product, renormalised only when it leaves [1e-150, 1e150]
            double m2 = lw_m * wj;
            if (!(m2 >= 1e-150 && m2 <= 1e150)) {
              int e1, e2;
              const double f1 = frexp(lw_m, &e1), f2 = frexp(wj, &e2);
              m2 = f1 * f2;
              lw_e += e1 + e2;
            }
            lw_m = m2;
          }
          hp_acc += hp_part;
        } else {
          wj = 0.0;
        }
      }
      if (!dry) {
        // ---- fold this warp's 32 spots into its output tiles (FP64 tensor pipe) ---------------------
        // DMMA runs on the same pipe as DFMA at the same FMA rate (measured), so every DMMA saved is time saved.
        // The label tiles spend 8x the useful work (one label per spot, eight columns per tile); when all spots of
        // the sub-tile carry ONE label k -- the common case on spatially coherent data -- their result is already in
        // the second moments: column d of U (w U)^T (the constant-1 row) is sum_j w_j y_j and sum_j w_j of the
        // sub-tile.  Its increment over this sub-tile is added to cluster k's sums; mixed sub-tiles use the label tiles.
        wst[lane] = wj;
        zst[lane] = (uint8_t) zfinal;
        const int zf0 = __shfl_sync(0xffffffffu, zfinal, 0);
        const bool uniform = fast_uniform_shortcut(D) && do_pairs && __all_sync(0xffffffffu, !active || zfinal == zf0);
        __syncwarp();
        // tile (i, NT-1) has index i*NT - i(i-1)/2 + NT-1-i in the upper-triangular enumeration
        double snap[NT];
#pragma unroll
        for (int i = 0; i < NT; i++) snap[i] = C[i * NT - i * (i - 1) / 2 + (NT - 1 - i)][D & 1];
        if (do_pairs) {
#pragma unroll
          for (int g = 0; g < 8; g++) {
            // spot 4 g + lq: half g >> 2, column 4 (g & 3) + lq, i.e. chunk 2 (g & 3) + (lq >> 1)
            const unsigned char *fb = ycur + (g >> 2) * YBOX + frag_row + (frag_sw ^ ((g & 3) << 5)) + frag_lo;
            const double wt = wst[4 * g + lq];
            double f[NT], b[NT];
#pragma unroll
            for (int i = 0; i < NT; i++) {
              f[i] = *reinterpret_cast<const double *>(fb + i * 8 * kRowB);
              b[i] = f[i] * wt;
            }
            int tix = 0;
#pragma unroll
            for (int i = 0; i < NT; i++)
#pragma unroll
              for (int jj = i; jj < NT; jj++) dmma_m8n8k4(C[tix++], f[i], b[jj]);
          }
        }
        if (uniform) {
          if (lq == ((D & 7) >> 1)) {   // the lanes that hold column d of the tiles (i, NT-1)
#pragma unroll
            for (int i = 0; i < NT; i++) {
              const int row = 8 * i + kc;
              if (row <= D) skacc[zf0 * DP + row] += C[i * NT - i * (i - 1) / 2 + (NT - 1 - i)][D & 1] - snap[i];
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; g++) {
            const unsigned char *fb = ycur + (g >> 2) * YBOX + frag_row + (frag_sw ^ ((g & 3) << 5)) + frag_lo;
            const double wt = wst[4 * g + lq];
            const int zt = zst[4 * g + lq];
#pragma unroll
            for (int h = 0; h < NH; h++) {
              const double hv = zt == kc + 8 * h ? wt : 0.0;
#pragma unroll
              for (int i = 0; i < NT; i++) dmma_m8n8k4(CH[i][h], *reinterpret_cast<const double *>(fb + i * 8 * kRowB), hv);
            }
          }
        }
      }
      __syncwarp();   // the buffer and the staging rows may be refilled from here on
    }
    if (dry) continue;
    // ---- fold warps in a fixed order and publish the segment's record ------------------------------------
    double lws = TDIST ? log(lw_m) + (double) lw_e * 0.69314718055994530942 : 0.0;
    {
      // this warp's buffers are idle (all of its copies have been consumed): they now carry its output tiles
      double2 *mine = reinterpret_cast<double2 *>(cfold + (size_t) warp * fold_stride + kc * 8 + 2 * lq);
#pragma unroll
      for (int t = 0; t < NTILES; t++) mine[t * 32] = make_double2(C[t][0], C[t][1]);
#pragma unroll
      for (int i = 0; i < NT; i++)
#pragma unroll
        for (int h = 0; h < NH; h++) mine[(NTILES + h * NT + i) * 32] = make_double2(CH[i][h][0], CH[i][h][1]);
    }
    if (FUSEC) __threadfence();                   // this segment's labels, before the barrier below
    block_reduce3_fast(hp_acc, nacc, lws, red, tid);   // contains a __syncthreads()
    if (FUSEC && tid == 0) atomicAdd(a.cfuse.done + my_colour, 1u);   // ... are final: the next colour class may read them
    if (tid == 0) {
      acc[a.SL.hp] = hp_acc + halfD * lws;
      acc[a.SL.nacc] = nacc;
    }
    if (do_pairs)
      for (int p = tid; p < NP; p += kFBlock) {
        int pa, pb;
        pair_from_index(p, D, pa, pb);
        const int i = pa >> 3, jj = pb >> 3;
        const int t = i * NT - i * (i - 1) / 2 + (jj - i);
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) s += cfold[(size_t) w * fold_stride + t * 64 + (pa & 7) * 8 + (pb & 7)];
        acc[a.SL.T + p] = s;
      }
    for (int e = tid; e < q * (D + 1); e += kFBlock) {   // (cluster k, row r): r < D -> s_k[r], r == D -> n_k
      const int k = e / (D + 1), r = e - k * (D + 1);
      const int t = NTILES + (k >> 3) * NT + (r >> 3);
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; w++)
        s += cfold[(size_t) w * fold_stride + t * 64 + (r & 7) * 8 + (k & 7)] +
             (fast_uniform_shortcut(D)
                  ? reinterpret_cast<const double *>(smraw + L.shared_bytes + (size_t) w * L.warp_stride + L.skacc)[k * DP + r]
                  : 0.0);
      if (r < D) acc[a.SL.sk + k * D + r] = s;
      else acc[a.SL.nk + k] = s;
    }
    __syncthreads();
    double *out = a.partials + (size_t) a.seg_slot[seg] * E;
    for (int e = tid; e < E; e += kFBlock) out[e] = acc[e];
    if (bseg && !dry) {
      // the CTA that completes the LAST boundary segment of this launch pushes the boundary labels of this colour to
      // the neighbours and releases their flags; the interior segments of the launch are still running
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last_boundary = atomicAdd(a.halo.done, 1u) == (uint32_t) (a.halo.nbound - 1);
      __syncthreads();
      if (s_last_boundary) {
        __threadfence();
        for (int l = 0; l < a.halo.nlegs; l++) {
          const HaloLeg leg = a.halo.legs[l];
          for (int k = leg.s0 + tid; k < leg.s1; k += kFBlock) leg.peer_z[a.halo.send_dst[k]] = __ldcg(a.z + a.halo.send_pos[k]);
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
          const uint32_t epoch = (it - 1u) * (uint32_t) a.halo.ncol + (uint32_t) a.halo.colour + 1u;
          for (int l = 0; l < a.halo.nlegs; l++) st_release_sys(a.halo.legs[l].peer_flag, epoch);
          *a.halo.done = 0u;
        }
      }
    }
    init_const_rows();     // the output tiles overwrote them
    fence_proxy_async();   // generic-proxy writes to the buffers before the next segment's TMA copies
    __syncthreads();
  }
}

template <bool TDIST, int NH, bool LEAN, bool FUSEC>
int launch_fast_t(const SweepArgs &a, int grid, cudaStream_t s) {
  constexpr int D = BSB_D;
  const FastSmem L = fast_smem_layout(D, a.PL.q, a.SL.E, TDIST, NH, a.maxdeg);
  static size_t granted[64] = {};
  const cudaError_t e = ensure_dynamic_smem(k_sweep_fast<D, TDIST, NH, LEAN, FUSEC>, L.bytes, granted);
  if (e != cudaSuccess) return (int) e;
  // Persistent launch: at most as many CTAs as the device holds at once, each walking its segments (seg = blockIdx.x,
  // + gridDim.x, ...).  The per-CTA set-up (parameter staging, logarithm table, barrier initialisation) is then paid
  // once per CTA instead of once per segment; the segments, their records and the fold order do not change.
  static int slots[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int resident = dev >= 0 && dev < 64 ? slots[dev] : 0;
  if (!resident) {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<D, TDIST, NH, LEAN, FUSEC>, kFBlock, L.bytes);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = per_sm > 0 && sms > 0 ? per_sm * sms : grid;
    if (dev >= 0 && dev < 64) slots[dev] = resident;   // for the largest q seen first; smaller q only fit better
  }
  static const bool persistent = std::getenv("BSB200_NO_PERSISTENT_SWEEP") == nullptr;
  if (persistent && grid > resident && !FUSEC) grid = resident;   // FUSEC: the caller made sure every CTA is resident
  k_sweep_fast<D, TDIST, NH, LEAN, FUSEC><<<grid, kFBlock, L.bytes, s>>>(a, *a.maps);
  return 0;
}

// shared precision, q <= kFastMaxQ only (the caller checks)
void launch_fast(int tdist, const SweepArgs &a, int grid, cudaStream_t s) {
  const int nh = (a.PL.q + 7) / 8;
  const bool lean = !(a.flags & (SW_STATS_ONLY | SW_DRY));
  if (lean && a.cfuse.on) {
    if (tdist) { if (nh == 1) launch_fast_t<true, 1, true, true>(a, grid, s); else launch_fast_t<true, 2, true, true>(a, grid, s); }
    else { if (nh == 1) launch_fast_t<false, 1, true, true>(a, grid, s); else launch_fast_t<false, 2, true, true>(a, grid, s); }
  } else if (lean) {
    if (tdist) { if (nh == 1) launch_fast_t<true, 1, true, false>(a, grid, s); else launch_fast_t<true, 2, true, false>(a, grid, s); }
    else { if (nh == 1) launch_fast_t<false, 1, true, false>(a, grid, s); else launch_fast_t<false, 2, true, false>(a, grid, s); }
  } else {
    if (tdist) { if (nh == 1) launch_fast_t<true, 1, false, false>(a, grid, s); else launch_fast_t<true, 2, false, false>(a, grid, s); }
    else { if (nh == 1) launch_fast_t<false, 1, false, false>(a, grid, s); else launch_fast_t<false, 2, false, false>(a, grid, s); }
  }
}

size_t fast_smem(int tdist, const SweepArgs &a) {
  return fast_smem_layout(BSB_D, a.PL.q, a.SL.E, tdist, (a.PL.q + 7) / 8, a.maxdeg).bytes;
}

FastRegistrar reg(BSB_D, launch_fast, fast_smem);

}   // namespace
}   // namespace bsb
