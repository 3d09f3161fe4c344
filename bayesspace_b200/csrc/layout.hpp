// layout.hpp -- layouts of the per-iteration parameter block, the sufficient-statistics records
// and the kernel argument structs, shared by host code and kernels.
//
// HBM data layout of a chain (all arrays in "internal" spot order = sorted by (shard, colour,
// original index), so that every colour phase streams contiguous memory):
//   Y     double [d][ld]     centred PCs, one coalesced column per PC (R's column-major layout)
//   z     uint8  [ld]        labels, 0-based
//   w     double [ld]        t-model scale weights
//   ell   int32  [maxdeg][ld] neighbour ids (internal order), -1 padded (ELL format)
//   orig  int32  [ld]        original spot id of each slot (keys the Philox counter)
#pragma once
#include <cstdint>

namespace bsb {

constexpr int kBlock = 128;       // threads per sweep CTA = spots per tile
constexpr int kMaxD = 32;
constexpr int kMaxDeg = 16;
constexpr int kGroups = 64;       // fixed virtual shards of the deterministic two-level reduction
constexpr int kTileLd = kBlock + 4;   // tile row stride = 4 (mod 16) doubles: DMMA fragment loads conflict free
constexpr uint32_t BSB_COMPAT_Q1_BIT = 1u;
constexpr uint32_t BSB_COMPAT_Q3_BIT = 4u;
constexpr int BSB_DEVERR_NUMERIC = 1;   // NaN probability / Wishart df <= 0 / not positive definite

// Offsets (in doubles) into the parameter block written once per iteration by k_param and read by
// every sweep CTA.  "Packed" = upper triangle, row-major (a <= b), off-diagonals doubled, so that
// x^T M x = sum_a x_a * sum_{b>=a} P_ab x_b.
struct ParamLayout {
  int q, d, nl, np;
  int mu;      // q*d        cluster means (centred coordinates)
  int PM;      // nl*np      packed density matrix: (R R^T)^-1 under Q1, else Lambda
  int g;       // q*d        M_k mu_k
  int c;       // q          mu_k^T M_k mu_k
  int logdet;  // nl         sum log diag(rooti)
  int PL;      // nl*np      packed Lambda (t model: quadratic form of the w update)
  int gL;      // q*d        Lambda_k mu_k
  int cL;      // q          mu_k^T Lambda_k mu_k
  int lam;     // nl*d*d     Lambda, full column-major
  int total;
};

inline ParamLayout make_param_layout(int q, int d, int nl) {
  ParamLayout L;
  L.q = q; L.d = d; L.nl = nl; L.np = d * (d + 1) / 2;
  int o = 0;
  L.mu = o; o += q * d;
  L.PM = o; o += nl * L.np;
  L.g = o; o += q * d;
  L.c = o; o += q;
  L.logdet = o; o += nl;
  L.PL = o; o += nl * L.np;
  L.gL = o; o += q * d;
  L.cL = o; o += q;
  L.lam = o; o += nl * d * d;
  L.total = o;
  return L;
}

// One sufficient-statistics record (per segment partial, per group sum, and the reduced total).
struct StatLayout {
  int q, d, nlT, np;
  int T;      // nlT*np   sum_j w_j y_ja y_jb (upper pairs, NOT doubled); nlT = 0 when not tracked
  int sk;     // q*d      sum_j w_j y_j  [z_j = k]
  int nk;     // q        sum_j w_j      [z_j = k]
  int cnt;    // q        #{j: z_j = k}
  int hp;     // 1        sum_j of the per-spot part of plogLik
  int nacc;   // 1        accepted label moves
  int E;
};

inline StatLayout make_stat_layout(int q, int d, int nlT) {
  StatLayout S;
  S.q = q; S.d = d; S.nlT = nlT; S.np = d * (d + 1) / 2;
  int o = 0;
  S.T = o; o += nlT * S.np;
  S.sk = o; o += q * d;
  S.nk = o; o += q;
  S.cnt = o; o += q;
  S.hp = o; o += 1;
  S.nacc = o; o += 1;
  S.E = o;
  return S;
}

enum SweepFlags : uint32_t {
  SW_Q3 = 1u,          // Potts count of the current label compares neighbour ids (quirk Q3)
  SW_STATS_ONLY = 2u,  // accumulate statistics of the current state, change nothing
  SW_DRY = 4u,         // evaluate proposals/ratios into the dry_* arrays, change nothing
  SW_NO_PAIRS = 8u     // skip the y y^T accumulation (normal model, shared precision)
};

struct SweepMaps;   // tma_maps.hpp
struct HaloLeg;     // comm.hpp

// Fused halo exchange of k_sweep_fast on the peer transport (one chain over several GPUs): the segments that hold
// boundary spots come FIRST in a colour's launch list; each waits, before it starts, for the neighbours' labels of the
// previous colour phase, and the CTA that finishes the last of them stores the boundary labels of this colour into the
// neighbours' ghost slots and releases their flags -- while the interior segments are still being swept.
// Epoch of (iteration it >= 1, colour c) = (it - 1) * ncol + c + 1; a phase waits for the epoch before it.
struct HaloFuse {
  const HaloLeg *legs;       // this colour's legs: every neighbouring rank, empty send lists included (device)
  int nlegs;                 // 0: no fused exchange
  int nbound;                // boundary segments = the first nbound entries of the launch list
  int ncol, colour;
  const int32_t *send_pos;   // own positions of the boundary labels, indexed by the legs' [s0, s1)
  const int32_t *send_dst;   // their positions in the peer's label array
  uint32_t *done;            // finished boundary segments of this launch (reset by the last one)
  unsigned long long timeout_ns;
};

// All colour classes of a SMALL chain in one launch (every CTA resident at once, one segment per CTA): everything
// that does not depend on the neighbours' labels -- random numbers, quadratic form, t-weight -- runs at once; a CTA
// of colour class c then waits until all segments of class c - 1 have published their labels (done[c - 1], reset by
// k_param every iteration), gathers its neighbours' labels and finishes.  Saves ncol - 1 launch latencies per iteration
// where the launches are latency bound (Visium: 13 CTAs per colour class).
struct ColourFuse {
  int on, ncol;
  const int32_t *colour_seg_start;   // ncol + 1 offsets into the (colour-major) launch list
  uint32_t *done;                    // ncol counters
};

struct SweepArgs {
  const SweepMaps *maps;      // host pointer: TMA descriptors of Y / ell / orig (k_sweep_fast), may be null
  const double *Y;
  int64_t ld;
  uint8_t *z;
  double *w;
  const int32_t *ell;
  int maxdeg;
  const int32_t *orig;
  const int32_t *seg_begin;   // segments of this launch
  const int32_t *seg_count;
  const int32_t *seg_slot;    // partial record written by each segment
  int nseg;
  const double *params;
  ParamLayout PL;
  double *partials;
  StatLayout SL;
  const uint32_t *iter;       // device counter: iterations completed
  uint32_t iter_add;          // iteration being run = *iter + iter_add
  uint32_t key0, key1;
  double gamma;
  uint32_t flags;
  int *err;
  ColourFuse cfuse;           // all colour classes in one launch (small chains, k_sweep_fast)
  HaloFuse halo;              // fused halo exchange (k_sweep_fast, peer transport); nlegs == 0 otherwise
  int32_t *dry_znew;          // dry-run outputs, original spot order
  double *dry_w, *dry_hp, *dry_hn, *dry_prob;
  int32_t *dry_acc;
};

struct ParamArgs {
  ParamLayout PL;
  StatLayout SL;
  double *params;
  const double *partials;
  const int32_t *group_start;   // kGroups+1 offsets into the slot range
  int ngroups;
  double *gsum;                 // ngroups * E scratch
  int gsum_ready;               // level 1 already done by k_reduce_groups
  const uint32_t *gepoch;       // peer transport: shard sums are double buffered, parity = *gepoch & 1 (else null)
  size_t gsum_stride;           // doubles between the two parities
  double *stats;                // E reduced statistics (kept for probes)
  const double *T_fixed;        // np: sum y y^T (normal model, shared precision) or nullptr
  const double *mu0c;           // d   mu0 - ybar
  const double *lambda0;        // d*d
  const double *lm0;            // d   lambda0 * mu0c
  const double *ybar;           // d
  double alpha, beta;
  int64_t n;
  int tdist, vvv;
  uint32_t compat;
  uint32_t key0, key1;
  uint32_t *iter;
  int nrep, thin;
  double *plogLik;              // nrep
  double *ychange;              // nrep or nullptr: accepted Y moves / n0 (deconv)
  double n0;                    // parents (deconv)
  double *snap_mu;              // nsnap * q*d
  double *snap_lambda;          // nsnap * nl*d*d
  uint32_t *zero_counters;      // counters k_param resets for the coming sweep (ColourFuse::done), or nullptr
  int n_zero;
  int finalize_only;            // reduce + plogLik only (after the last iteration)
  int precision_form;           // 1: density in precision form (deconv): M = Lambda, no Sigma chain
  int *err;
};

}   // namespace bsb
