// tile_stats.cuh -- fused sufficient statistics of one tile of kBlock spots (device).
//
// After a CTA has updated a tile, every thread has staged sqrt(w_j) y_j, sqrt(w_j), w_j and the
// final label of its spot in shared memory.  Each thread then owns a few "roles" -- one (a,b)
// pair of sum w y y^T, one column of the per-cluster sums s_k, n_k or the counts -- and folds the
// whole tile into its roles in spot order.  The order is fixed, so records are bit-reproducible.
// Labels are piecewise constant along a tile (spatial domains), hence the run-length flushes.
//
// Replaces find(z == k) / sum(Yrows, 0) / (Y - mu)^T diag(w) (Y - mu) of the reference's mu and
// lambda updates (src/cluster.cpp:249-253, :265, :482-486, :500, :884-887, :901), which re-read
// all of Y in separate passes.
#pragma once
#include "layout.hpp"

namespace bsb {

__device__ __forceinline__ void pair_from_index(int p, int d, int &a, int &b) {
  a = 0;
  int rem = p, len = d;
  while (rem >= len) { rem -= len; len--; a++; }
  b = a + rem;
}

// zs[t] == 255 marks an empty slot of the tile (its staged values are zero).
template <int D>
__device__ __forceinline__ void tile_accumulate(double *acc, const StatLayout &SL, const double *tile,
                                                const double *sws, const double *wvs, const uint8_t *zs,
                                                bool do_pairs, bool per_cluster_pairs, int tid) {
  constexpr int NP = D * (D + 1) / 2;
  for (int role = tid; role < NP + D + 2; role += kBlock) {
    if (role < NP) {
      if (!do_pairs) continue;
      int pa, pb;
      pair_from_index(role, D, pa, pb);
      const double *ua = tile + pa * kTileLd, *ub = tile + pb * kTileLd;
      if (!per_cluster_pairs) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int t = 0; t < kBlock; t += 2) {
          const double2 va = *reinterpret_cast<const double2 *>(ua + t);
          const double2 vb = *reinterpret_cast<const double2 *>(ub + t);
          s0 = fma(va.x, vb.x, s0);
          s1 = fma(va.y, vb.y, s1);
        }
        acc[SL.T + role] += s0 + s1;
      } else {
        int cur = 255;
        double s = 0.0;
        for (int t = 0; t < kBlock; t++) {
          const int k = zs[t];
          if (k == 255) continue;
          if (k != cur) {
            if (cur != 255) acc[SL.T + cur * NP + role] += s;
            s = 0.0;
            cur = k;
          }
          s = fma(ua[t], ub[t], s);
        }
        if (cur != 255) acc[SL.T + cur * NP + role] += s;
      }
    } else {
      const int c = role - NP;   // c < D: column of s_k; c == D: n_k; c == D+1: counts
      const int off = c < D ? SL.sk : (c == D ? SL.nk : SL.cnt);
      const int stride = c < D ? D : 1, col = c < D ? c : 0;
      int cur = 255;
      double s = 0.0;
      for (int t = 0; t < kBlock; t++) {
        const int k = zs[t];
        if (k == 255) continue;
        if (k != cur) {
          if (cur != 255) acc[off + cur * stride + col] += s;
          s = 0.0;
          cur = k;
        }
        s += c < D ? sws[t] * tile[c * kTileLd + t] : (c == D ? wvs[t] : 1.0);
      }
      if (cur != 255) acc[off + cur * stride + col] += s;
    }
  }
}

// Fixed-order CTA reduction of two per-thread scalars; result valid on thread 0.
__device__ __forceinline__ void block_reduce2(double &x, double &y, double *red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  if ((tid & 31) == 0) { red[2 * (tid >> 5)] = x; red[2 * (tid >> 5) + 1] = y; }
  __syncthreads();
  if (tid == 0) {
    double h = 0.0, na = 0.0;
    for (int wv = 0; wv < kBlock / 32; wv++) { h += red[2 * wv]; na += red[2 * wv + 1]; }
    x = h; y = na;
  }
}

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
