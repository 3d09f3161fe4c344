// tile_stats.cuh -- fused sufficient statistics of one tile of kBlock spots (device).
//
// After a CTA has updated a tile, every thread has staged u_j = sqrt(w_j) y_j (one shared-memory row
// per PC, kTileLd doubles apart), sqrt(w_j), w_j and the final label of its spot.  Two reductions
// then run over the tile, both in a fixed order so that records are bit-reproducible:
//
//  * pooled second moments  T = sum_j u_j u_j^T  (the only GEMM-shaped piece of the path: a
//    d x 128 by 128 x d FP64 SYRK per tile).  It runs on the FP64 tensor pipe with
//    mma.sync.m8n8k4.f64 (DMMA): each warp owns every 4th group of 4 spots, keeps the 8x8 output
//    tiles of the upper triangle in registers across the whole segment, and the 4 warps are
//    folded once per segment.  Measured on B200: DMMA sustains the full FP64 rate (37 TFLOP/s) at
//    1/8 of the issue slots of DFMA, which is what matters here (the sweep is issue-bound).
//  * per-cluster sums  s_k = sum w y, n_k = sum w, counts: thread (column c, part p) walks spots
//    p, p+P, p+2P, ... of the tile with a run-length flush (labels are piecewise constant along a
//    tile), into its private accumulator row; parts are folded once per segment.
//
// Replaces find(z == k) / sum(Yrows, 0) / (Y - mu)^T diag(w) (Y - mu) of the reference's mu and
// lambda updates (src/cluster.cpp:249-253, :265, :482-486, :500, :884-887, :901), which re-read
// all of Y in separate passes.
#pragma once
#include "layout.hpp"

namespace bsb {

// Ampere-style asynchronous global -> shared copies (LDGSTS): no register staging, no scoreboard stall.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// DMMA output tiles kept per warp: the upper 8x8 tiles of U U^T plus nh label tiles per row tile.
__host__ __device__ inline int stats_frags(int D, int nh) {
  const int NT = (D + 1 + 7) / 8;
  return NT * (NT + 1) / 2 + nh * NT;
}

template <int D>
struct TileGeom {
  static constexpr int NT = (D + 1 + 7) / 8;      // 8-row tiles of the staged rows: d PCs, then sqrt(w)
  static constexpr int DP = NT * 8;               // padded rows of the shared-memory tile
  static constexpr int NTILES = NT * (NT + 1) / 2;   // upper-triangular 8x8 output tiles
  static constexpr int NP = D * (D + 1) / 2;
};

__device__ __forceinline__ void pair_from_index(int p, int d, int &a, int &b) {
  a = 0;
  int rem = p, len = d;
  while (rem >= len) { rem -= len; len--; a++; }
  b = a + rem;
}

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// This warp's share of one tile, everything on the FP64 tensor pipe:
//   C  += U U^T          (U = staged rows: sqrt(w) y_1..d, then sqrt(w); upper 8x8 tiles)  -> T
//   CH += U H^T          (H_k[t] = sqrt(w_t) [z_t = k], built on the fly, k < 8*nh <= 16) -> s_k (rows < d), n_k (row d)
// A fragment: lane holds U[8i + (lane>>2)][t0 + (lane&3)]; the B fragment of row-tile j is the same
// register as the A fragment of row-tile j, so NT loads feed all the DMMAs of a group of 4 spots.
template <int D>
__device__ __forceinline__ void tile_stats_mma(double *cacc, const double *tile, const uint8_t *zs, bool do_pairs, int nh,
                                               int warp, int lane, int nh_layout = -1) {
  // nh: label tiles folded by THIS call; nh_layout: label tiles the accumulator area was laid out for (a caller may
  // fold the second moments and the label tiles in separate calls)
  constexpr int NT = TileGeom<D>::NT, NTILES = TileGeom<D>::NTILES, NW = kBlock / 32;
  const int NFRAG = stats_frags(D, nh_layout < 0 ? nh : nh_layout);
  // The output tiles live in shared memory between tiles (warp-private, 2 doubles per lane and tile): they are
  // only in registers while the DMMAs run, which keeps the per-spot phase of the kernel below 128 registers.
  double2 *mine = reinterpret_cast<double2 *>(cacc + (size_t) warp * NFRAG * 64 + (lane >> 2) * 8 + 2 * (lane & 3));
  double C[NTILES][2], CH[NT][2][2];
  if (do_pairs) {
#pragma unroll
    for (int t = 0; t < NTILES; t++) { const double2 v = mine[t * 32]; C[t][0] = v.x; C[t][1] = v.y; }
  }
#pragma unroll
  for (int i = 0; i < NT; i++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (h < nh) { const double2 v = mine[(NTILES + h * NT + i) * 32]; CH[i][h][0] = v.x; CH[i][h][1] = v.y; }
      else CH[i][h][0] = CH[i][h][1] = 0.0;
    }
  const double *base = tile + (lane >> 2) * kTileLd + (lane & 3);
  const double *swrow = tile + D * kTileLd + (lane & 3);
  const int kc = lane >> 2;
#pragma unroll
  for (int s = 0; s < kBlock / 4 / NW; s++) {
    const int t0 = (warp + s * NW) * 4;
    double f[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) f[i] = base[i * 8 * kTileLd + t0];
    if (do_pairs) {
      int tix = 0;
#pragma unroll
      for (int i = 0; i < NT; i++)
#pragma unroll
        for (int j = i; j < NT; j++) dmma_m8n8k4(C[tix++], f[i], f[j]);
    }
    if (nh == 0) continue;
    const int zt = zs[t0 + (lane & 3)];
    const double swt = swrow[t0];
    const double h0 = zt == kc ? swt : 0.0;
#pragma unroll
    for (int i = 0; i < NT; i++) dmma_m8n8k4(CH[i][0], f[i], h0);
    if (nh > 1) {
      const double h1 = zt == kc + 8 ? swt : 0.0;
#pragma unroll
      for (int i = 0; i < NT; i++) dmma_m8n8k4(CH[i][1], f[i], h1);
    }
  }
  if (do_pairs) {
#pragma unroll
    for (int t = 0; t < NTILES; t++) mine[t * 32] = make_double2(C[t][0], C[t][1]);
  }
#pragma unroll
  for (int i = 0; i < NT; i++)
#pragma unroll
    for (int h = 0; h < 2; h++)
      if (h < nh) mine[(NTILES + h * NT + i) * 32] = make_double2(CH[i][h][0], CH[i][h][1]);
}

// Segment end: fold the warps' output tiles (fixed order) into the record.  Contains a __syncthreads().
template <int D>
__device__ __forceinline__ void tile_stats_fold(const double *cacc, double *acc, const StatLayout &SL, bool do_pairs,
                                                int q, int nh, int tid) {
  constexpr int NT = TileGeom<D>::NT, NTILES = TileGeom<D>::NTILES, NP = TileGeom<D>::NP, NW = kBlock / 32;
  const int NFRAG = stats_frags(D, nh);
  __syncthreads();
  if (do_pairs)
    for (int p = tid; p < NP; p += kBlock) {
      int a, b;
      pair_from_index(p, D, a, b);
      const int i = a >> 3, j = b >> 3;
      const int t = i * NT - i * (i - 1) / 2 + (j - i);
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) s += cacc[((size_t) w * NFRAG + t) * 64 + (a & 7) * 8 + (b & 7)];
      acc[SL.T + p] = s;
    }
  for (int e = tid; e < q * (D + 1); e += kBlock) {   // (cluster k, row a): a < D -> s_k[a], a == D -> n_k
    const int k = e / (D + 1), a = e - k * (D + 1);
    const int t = NTILES + (k >> 3) * NT + (a >> 3);
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w++) s += cacc[((size_t) w * NFRAG + t) * 64 + (a & 7) * 8 + (k & 7)];
    if (a < D) acc[SL.sk + k * D + a] = s;
    else acc[SL.nk + k] = s;
  }
}

// Per-cluster second moments (VVV samplers): scalar run-length path, one (a,b) pair per thread.
template <int D>
__device__ __forceinline__ void tile_pairs_per_cluster(double *accT, const double *tile, const uint8_t *zs, int tid) {
  constexpr int NP = TileGeom<D>::NP;
  for (int role = tid; role < NP; role += kBlock) {
    int pa, pb;
    pair_from_index(role, D, pa, pb);
    const double *ua = tile + pa * kTileLd, *ub = tile + pb * kTileLd;
    int cur = 255;
    double s = 0.0;
    for (int t = 0; t < kBlock; t++) {
      const int k = zs[t];
      if (k == 255) continue;
      if (k != cur) {
        if (cur != 255) accT[cur * NP + role] += s;
        s = 0.0;
        cur = k;
      }
      s = fma(ua[t], ub[t], s);
    }
    if (cur != 255) accT[cur * NP + role] += s;
  }
}

// Per-cluster sums.  Role (c, p): column c in [0, D+2) (D PCs, then w, then the count), part p in [0, P).
// accP: P * q * (D+2) doubles, accP[(p*q + k)*(D+2) + c].  zs[t] == 255 marks an empty slot.
template <int D>
__device__ __forceinline__ void tile_cluster_sums(double *accP, int P, int q, const double *tile, const double *sws,
                                                  const double *wvs, const uint8_t *zs, int tid) {
  constexpr int NC = D + 2;
  if (tid >= NC * P) return;
  const int c = tid / P, p = tid - c * P;
  double *mine = accP + (size_t) p * q * NC + c;
  int cur = 255;
  double s = 0.0;
  for (int t = p; t < kBlock; t += P) {
    const int k = zs[t];
    if (k == 255) continue;
    if (k != cur) {
      if (cur != 255) mine[cur * NC] += s;
      s = 0.0;
      cur = k;
    }
    s += c < D ? sws[t] * tile[c * kTileLd + t] : (c == D ? wvs[t] : 1.0);
  }
  if (cur != 255) mine[cur * NC] += s;
}

// Segment end: fold the P parts (fixed order) into the record's s_k / n_k / count entries.
template <int D>
__device__ __forceinline__ void tile_cluster_fold(const double *accP, int P, int q, double *acc, const StatLayout &SL,
                                                  int tid) {
  constexpr int NC = D + 2;
  for (int e = tid; e < q * NC; e += kBlock) {
    const int k = e / NC, c = e - k * NC;
    double s = 0.0;
    for (int p = 0; p < P; p++) s += accP[((size_t) p * q + k) * NC + c];
    if (c < D) acc[SL.sk + k * D + c] = s;
    else if (c == D) acc[SL.nk + k] = s;
    else acc[SL.cnt + k] = s;
  }
}

// Number of parts of tile_cluster_sums: as many as fit the CTA and a 48 KB accumulator budget.
__host__ __device__ inline int cluster_sum_parts(int D, int q) {
  int P = kBlock / (D + 2);
  const int cap = (48 * 1024) / (q * (D + 2) * 8);
  if (P > cap) P = cap;
  return P < 1 ? 1 : P;
}

// Fixed-order CTA reduction of three per-thread scalars; results valid on thread 0.  red: 3 * warps doubles.
__device__ __forceinline__ void block_reduce3(double &x, double &y, double &z, double *red, int tid) {
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
    for (int wv = 0; wv < kBlock / 32; wv++) { a += red[3 * wv]; b += red[3 * wv + 1]; c += red[3 * wv + 2]; }
    x = a; y = b; z = c;
  }
}

// sum_j log w_j without one log per spot: running product of mantissas + integer exponents.
struct LogProd {
  double m = 1.0;
  int e = 0;
  __device__ __forceinline__ void mul(double w) {
    int ex;
    m *= frexp(w, &ex);
    e += ex;
    if (m < 1e-280) { m = frexp(m, &ex); e += ex; }
  }
  __device__ __forceinline__ double value() const { return log(m) + (double) e * 0.69314718055994530942; }
};

template <int D>
__device__ __forceinline__ double quadform(const double (&y)[D], const double *__restrict__ P) {
  double acc = 0.0;
  int p = 0;
#pragma unroll
  for (int a = 0; a < D; a++) {
    double s = 0.0;
#pragma unroll
    for (int b = a; b < D; b++) s = fma(P[p++], y[b], s);
    acc = fma(y[a], s, acc);
  }
  return acc;
}

template <int D>
__device__ __forceinline__ double dotv(const double (&y)[D], const double *__restrict__ g) {
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < D; c++) s = fma(y[c], g[c], s);
  return s;
}

}   // namespace bsb
