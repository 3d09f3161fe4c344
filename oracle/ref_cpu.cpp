// oracle/ref_cpu.cpp -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// CPU restatement of the BayesSpace MCMC sampler hot path, written from the reference's
// behaviour (reference = /root/reference, R package BayesSpace v1.15.3).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// PARITY STATUS: the reference (R + Rcpp + RcppArmadillo + RcppDist) cannot be built or run in
// this image, and its own tests pin no sampler numerics (only tests/testthat/test-find_neighbors.R
// holds a known answer, reproduced in tests/).  The sampler arithmetic below is therefore
// "parity unpinned": it follows src/cluster.cpp line by line, and App. C of SURVEY.md for the
// un-vendored RcppDist / Rcpp-sugar / Armadillo semantics.  Its reading of the reference is cross-checked by a
// second, independent restatement (tests/spec_numpy.py, numpy.linalg in the reference's operation order): the two
// must produce the same chains for all five samplers (tests/test_oracle.py).
//
// What is restated (file:line = /root/reference/...):
//   dmvnrm_arma_fast            src/cluster.cpp:83-106   (incl. quirk Q1: T*x instead of x^T*T)
//   dmvnrm_prec_arma_fast       src/cluster.cpp:109-132
//   inplace_tri_mat_mult_t      src/cluster.cpp:68-78
//   adaptive_mcmc               src/cluster.cpp:49-65
//   convert_neighbors/str_split src/cluster.cpp:134-145, src/utils.h:10-62, src/neighbor.h:67-80
//   convert_neighbors2mtx       src/cluster.cpp:147-159
//   find_subspot_neighbors      src/cluster.cpp:161-215
//   iterate                     src/cluster.cpp:217-321
//   iterate_vvv                 src/cluster.cpp:323-444  (incl. quirk Q3 at :404)
//   iterate_t                   src/cluster.cpp:446-568  (incl. quirk Q2 at :508)
//   iterate_t_vvv               src/cluster.cpp:570-710
//   iterate_deconv              src/cluster.cpp:742-1141
//   .find_neighbors             R/spatialCluster.R:227-302
//   find_neighbors              R/utils.R:13-26
//   rmvnorm / rwish (RcppDist), sample() (Rcpp sugar): SURVEY.md App. C
//
// Two deliberate departures, both required by BASELINE.json's north_star:
//  (1) Randomness.  R's Mersenne-Twister stream is not a parity goal.  Every variate is drawn from
//      a counter-based Philox4x32-10 stream keyed on (seed; purpose, index, iteration, slot), so a
//      variate belongs to "spot j at iteration i", not to a position in a serial stream.
//  (2) Visiting order.  The reference visits spots j = 0..n-1.  Here the z/w scan takes an explicit
//      visiting order; order == NULL reproduces the reference's identity order.  Passing the spots
//      sorted by (colour, index) makes this sequential scan coincide with the chromatic (parallel
//      per colour class) scan the CUDA path performs, which is what the bit-level chain-parity
//      tests rely on.  All arithmetic (per-spot re-factorisation included) stays the reference's.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> vecd;
static const double LOG2PI = std::log(2.0 * M_PI);

// ----------------------------------------------------------------------------------------------
// Dense helpers (column-major d x d), stand-ins for the Armadillo calls the reference makes.
// ----------------------------------------------------------------------------------------------
inline double &at(vecd &m, int d, int i, int j) { return m[(size_t) j * d + i]; }
inline double at(const vecd &m, int d, int i, int j) { return m[(size_t) j * d + i]; }

// arma::chol(X): upper R with R^T R = X; reads the upper triangle only (LAPACK dpotrf 'U').
bool chol_upper(const vecd &X, int d, vecd &R) {
  R.assign((size_t) d * d, 0.0);
  for (int j = 0; j < d; j++) {
    double s = at(X, d, j, j);
    for (int k = 0; k < j; k++) s -= at(R, d, k, j) * at(R, d, k, j);
    if (!(s > 0.0)) return false;
    double rjj = std::sqrt(s);
    at(R, d, j, j) = rjj;
    for (int i = j + 1; i < d; i++) {
      double t = at(X, d, j, i);
      for (int k = 0; k < j; k++) t -= at(R, d, k, j) * at(R, d, k, i);
      at(R, d, j, i) = t / rjj;
    }
  }
  return true;
}

// arma::inv(trimatu(R)): inverse of an upper-triangular matrix (back substitution).
void inv_upper(const vecd &R, int d, vecd &T) {
  T.assign((size_t) d * d, 0.0);
  for (int j = 0; j < d; j++) {
    at(T, d, j, j) = 1.0 / at(R, d, j, j);
    for (int i = j - 1; i >= 0; i--) {
      double s = 0.0;
      for (int k = i + 1; k <= j; k++) s += at(R, d, i, k) * at(T, d, k, j);
      at(T, d, i, j) = -s / at(R, d, i, i);
    }
  }
}

// arma::inv(X): general inverse (LAPACK getrf/getri) -> Gauss-Jordan with partial pivoting.
bool inv_general(const vecd &X, int d, vecd &Xi) {
  vecd a(X);
  Xi.assign((size_t) d * d, 0.0);
  for (int i = 0; i < d; i++) at(Xi, d, i, i) = 1.0;
  for (int c = 0; c < d; c++) {
    int p = c;
    double best = std::fabs(at(a, d, c, c));
    for (int r = c + 1; r < d; r++)
      if (std::fabs(at(a, d, r, c)) > best) { best = std::fabs(at(a, d, r, c)); p = r; }
    if (!(best > 0.0)) return false;
    if (p != c)
      for (int j = 0; j < d; j++) {
        std::swap(at(a, d, p, j), at(a, d, c, j));
        std::swap(at(Xi, d, p, j), at(Xi, d, c, j));
      }
    double piv = at(a, d, c, c);
    for (int j = 0; j < d; j++) { at(a, d, c, j) /= piv; at(Xi, d, c, j) /= piv; }
    for (int r = 0; r < d; r++) {
      if (r == c) continue;
      double f = at(a, d, r, c);
      if (f == 0.0) continue;
      for (int j = 0; j < d; j++) {
        at(a, d, r, j) -= f * at(a, d, c, j);
        at(Xi, d, r, j) -= f * at(Xi, d, c, j);
      }
    }
  }
  return true;
}

void matmul(const vecd &A, const vecd &B, int d, vecd &C) {   // C = A*B
  C.assign((size_t) d * d, 0.0);
  for (int j = 0; j < d; j++)
    for (int k = 0; k < d; k++) {
      double b = at(B, d, k, j);
      for (int i = 0; i < d; i++) at(C, d, i, j) += at(A, d, i, k) * b;
    }
}
void matmul_tn(const vecd &A, const vecd &B, int d, vecd &C) {   // C = A^T*B
  C.assign((size_t) d * d, 0.0);
  for (int j = 0; j < d; j++)
    for (int i = 0; i < d; i++) {
      double s = 0.0;
      for (int k = 0; k < d; k++) s += at(A, d, k, i) * at(B, d, k, j);
      at(C, d, i, j) = s;
    }
}
void matvec(const vecd &A, const double *x, int d, double *y) {
  for (int i = 0; i < d; i++) {
    double s = 0.0;
    for (int j = 0; j < d; j++) s += at(A, d, i, j) * x[j];
    y[i] = s;
  }
}

// ----------------------------------------------------------------------------------------------
// Counter-based randomness: Philox4x32-10 (Salmon et al., SC'11; Random123 known answers in tests).
// counter = (index, slot, iteration, purpose), key = seed.
// ----------------------------------------------------------------------------------------------
enum Purpose : uint32_t {
  P_MU = 1,          // index = cluster k, normals m = 0..d-1
  P_WISH_NORM = 2,   // index = cluster k (0 for shared precision), normal m = i(i-1)/2 + j
  P_WISH_CHI = 3,    // index = k*1024 + i, gamma attempts
  P_SPOT_U = 4,      // index = spot, slot 0: (u_proposal, u_accept)
  P_SPOT_GAMMA = 5,  // index = spot, gamma attempts for w
  P_ERR = 6,         // index = subspot, normal m = column
  P_SPOT_YACC = 7    // index = parent spot, slot 0: u_accept for the joint Y proposal
};

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
    uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
  uint32_t key[2];
  explicit Rng(uint64_t seed) { key[0] = (uint32_t) seed; key[1] = (uint32_t) (seed >> 32); }
  static inline double u53(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t) hi << 32) | lo;
    return ((double) (x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  inline void uniform2(uint32_t purpose, uint32_t index, uint32_t iter, uint32_t slot, double &u0,
                       double &u1) const {
    uint32_t c[4] = {index, slot, iter, purpose}, o[4];
    philox4x32_10(c, key, o);
    u0 = u53(o[0], o[1]);
    u1 = u53(o[2], o[3]);
  }
  inline void normal2(uint32_t purpose, uint32_t index, uint32_t iter, uint32_t slot, double &n0,
                      double &n1) const {   // Box-Muller
    double u0, u1;
    uniform2(purpose, index, iter, slot, u0, u1);
    double r = std::sqrt(-2.0 * std::log(u0));
    double a = 2.0 * M_PI * u1;
    n0 = r * std::cos(a);
    n1 = r * std::sin(a);
  }
  inline double normal(uint32_t purpose, uint32_t index, uint32_t iter, uint32_t m) const {
    double a, b;
    normal2(purpose, index, iter, m >> 1, a, b);
    return (m & 1u) ? b : a;
  }
  // R::rgamma(shape, scale) stand-in: Marsaglia-Tsang (2000); shape<1 boosted by u^(1/shape).
  double gamma(uint32_t purpose, uint32_t index, uint32_t iter, double shape, double scale) const {
    if (!(shape > 0.0) || !(scale > 0.0)) return std::numeric_limits<double>::quiet_NaN();
    double boost = 1.0, a = shape;
    if (a < 1.0) {
      double u0, u1;
      uniform2(purpose, index, iter, 0x80000000u, u0, u1);
      boost = std::pow(u0, 1.0 / a);
      a += 1.0;
    }
    const double dd = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * dd);
    for (uint32_t t = 0; t < 0x3fffffffu; t++) {
      double x, x_unused, u, u_unused;
      normal2(purpose, index, iter, 2u * t, x, x_unused);
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      uniform2(purpose, index, iter, 2u * t + 1u, u, u_unused);
      double x2 = x * x;
      if (u < 1.0 - 0.0331 * x2 * x2) return dd * v * scale * boost;
      if (std::log(u) < 0.5 * x2 + dd * (1.0 - v + std::log(v))) return dd * v * scale * boost;
    }
    return std::numeric_limits<double>::quiet_NaN();
  }
};

// ---- per-spot draws of the cluster samplers' z / w scan (mirrors bayesspace_b200/csrc/rng.cuh) ----
// One Philox block (spot, 0, iteration, P_SPOT_U) serves the scan's three uniforms: word 0 -> proposal
// (sample(qlessk, 1), src/cluster.cpp:278), words 1..2 -> the 53-bit uniform of the accept step (:304),
// word 3 -> the uniform of the first Marsaglia-Tsang attempt of R::rgamma (:518).  The normal of attempt t
// comes from block (spot, t, iteration, P_SPOT_GAMMA); attempts t >= 1 take their uniform from word 0 of
// block (spot, 0x40000000 + t, iteration, P_SPOT_GAMMA).
struct SpotDraw {
  uint32_t prop_bits;
  double u_acc, u_first;
};
inline double u32c(uint32_t x) { return ((double) x + 0.5) * (1.0 / 4294967296.0); }
inline SpotDraw spot_draw(const Rng &rng, uint32_t index, uint32_t iter) {
  uint32_t c[4] = {index, 0u, iter, P_SPOT_U}, o[4];
  philox4x32_10(c, rng.key, o);
  return SpotDraw{o[0], Rng::u53(o[1], o[2]), u32c(o[3])};
}
// sample(qlessk, 1): element floor((q-1) * bits / 2^32) of the labels 1..q without z_prev, ascending.
inline int propose_label_bits(int z_prev, int q, uint32_t bits) {
  const int t = (int) (((uint64_t) bits * (uint64_t) (q - 1)) >> 32);
  return (t + 1 < z_prev) ? t + 1 : t + 2;
}
// R::rgamma(shape >= 1, scale) stand-in (Marsaglia-Tsang) with the draw layout above.
inline double gamma_spot(const Rng &rng, uint32_t index, uint32_t iter, double shape, double scale, double u_first) {
  if (!(shape >= 1.0) || !(scale > 0.0)) return std::numeric_limits<double>::quiet_NaN();
  const double dd = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * dd);
  for (uint32_t t = 0; t < 0x3fffffffu; t++) {
    double x, x_unused;
    rng.normal2(P_SPOT_GAMMA, index, iter, t, x, x_unused);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = u_first;
    if (t) {
      uint32_t cc[4] = {index, 0x40000000u + t, iter, P_SPOT_GAMMA}, o[4];
      philox4x32_10(cc, rng.key, o);
      u = u32c(o[0]);
    }
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return dd * v * scale;
    if (std::log(u) < 0.5 * x2 + dd * (1.0 - v + std::log(v))) return dd * v * scale;
  }
  return std::numeric_limits<double>::quiet_NaN();
}

// Rcpp sugar sample(x,1,TRUE,{1-p,p}): probabilities sorted descending, one
// uniform, cumulative inversion.  Returns true when the second element ("new") is picked.
inline bool sample_two(double p_new, double u) {
  if (p_new >= 0.5) return u <= p_new;
  return u > 1.0 - p_new;
}
// Rcpp sugar sample(qlessk,1): x[floor(len*u)] over the labels 1..q without z_prev, ascending.
inline int propose_label(int z_prev, int q, double u) {
  int t = (int) std::floor((q - 1) * u);
  if (t > q - 2) t = q - 2;
  return (t + 1 < z_prev) ? t + 1 : t + 2;
}

// ----------------------------------------------------------------------------------------------
// Densities (src/cluster.cpp:68-132)
// ----------------------------------------------------------------------------------------------
// inplace_tri_mat_mult_t: x <- T x for upper-triangular T (row i uses columns j >= i).
inline void inplace_tri_mat_mult_t(double *x, const vecd &T, int d) {
  for (int i = 0; i < d; i++) {
    double tmp = 0.0;
    for (int j = i; j < d; j++) tmp += at(T, d, i, j) * x[j];
    x[i] = tmp;
  }
}

// dmvnrm_arma_fast(x(1 x d), mean, sigma, log=true): rooti = inv(trimatu(chol(sigma)));
// quirk Q1: z = rooti * (x-mean) rather than (x-mean) * rooti.
double dmvnrm_arma_fast(const double *x, const double *mean, const vecd &sigma, int d,
                        bool q1_compat) {
  vecd R, T;
  if (!chol_upper(sigma, d, R)) return std::numeric_limits<double>::quiet_NaN();
  inv_upper(R, d, T);
  double rootisum = 0.0;
  for (int i = 0; i < d; i++) rootisum += std::log(at(T, d, i, i));
  const double other = rootisum - (double) d / 2.0 * LOG2PI;
  vecd z(d);
  for (int i = 0; i < d; i++) z[i] = x[i] - mean[i];
  if (q1_compat) {
    inplace_tri_mat_mult_t(z.data(), T, d);
  } else {   // the Rcpp-gallery original: z <- z * rooti  (z_j = sum_{i<=j} z_i T_ij)
    vecd y(d);
    for (int j = 0; j < d; j++) {
      double s = 0.0;
      for (int i = 0; i <= j; i++) s += z[i] * at(T, d, i, j);
      y[j] = s;
    }
    z = y;
  }
  double dot = 0.0;
  for (int i = 0; i < d; i++) dot += z[i] * z[i];
  return other - 0.5 * dot;
}

// dmvnrm_prec_arma_fast: rooti = trimatu(chol(lambda)); z = rooti * (x-mean).
double dmvnrm_prec_arma_fast(const double *x, const double *mean, const vecd &lambda, int d) {
  vecd U;
  if (!chol_upper(lambda, d, U)) return std::numeric_limits<double>::quiet_NaN();
  double rootisum = 0.0;
  for (int i = 0; i < d; i++) rootisum += std::log(at(U, d, i, i));
  const double other = rootisum - (double) d / 2.0 * LOG2PI;
  vecd z(d);
  for (int i = 0; i < d; i++) z[i] = x[i] - mean[i];
  inplace_tri_mat_mult_t(z.data(), U, d);
  double dot = 0.0;
  for (int i = 0; i < d; i++) dot += z[i] * z[i];
  return other - 0.5 * dot;
}

// ----------------------------------------------------------------------------------------------
// RcppDist stand-ins
// ----------------------------------------------------------------------------------------------
// rmvnorm(1, mean, S): z * chol(S) + mean, z a row of d standard normals.
bool rmvnorm1(const Rng &rng, uint32_t purpose, uint32_t index, uint32_t iter, const double *mean,
              const vecd &S, int d, double *out) {
  vecd R;
  if (!chol_upper(S, d, R)) return false;
  vecd z(d);
  for (int i = 0; i < d; i++) z[i] = rng.normal(purpose, index, iter, (uint32_t) i);
  for (int j = 0; j < d; j++) {
    double s = 0.0;
    for (int i = 0; i <= j; i++) s += z[i] * at(R, d, i, j);
    out[j] = s + mean[j];
  }
  return true;
}

// rwish(int df, S): Bartlett. A lower-tri, A(i,j)~N(0,1) (i>j), A(i,i)=sqrt(rchisq(df-i));
// B = A^T chol(S); returns B^T B.   rchisq(v) = rgamma(v/2, scale 2).
bool rwish(const Rng &rng, uint32_t index, uint32_t iter, int df, const vecd &S, int d, vecd &W) {
  vecd A((size_t) d * d, 0.0), C;
  for (int i = 1; i < d; i++)
    for (int j = 0; j < i; j++)
      at(A, d, i, j) = rng.normal(P_WISH_NORM, index, iter, (uint32_t) (i * (i - 1) / 2 + j));
  for (int i = 0; i < d; i++) {
    double chi = rng.gamma(P_WISH_CHI, index * 1024u + (uint32_t) i, iter, (df - i) / 2.0, 2.0);
    if (!(chi > 0.0)) return false;   // Q9: df - i <= 0 -> NaN in the reference
    at(A, d, i, i) = std::sqrt(chi);
  }
  if (!chol_upper(S, d, C)) return false;
  vecd B;
  matmul_tn(A, C, d, B);
  matmul_tn(B, B, d, W);
  return true;
}

// ----------------------------------------------------------------------------------------------
// Neighbour graph (CSR, 0-based)
// ----------------------------------------------------------------------------------------------
struct Csr {
  const int64_t *ptr;
  const int32_t *idx;
  inline int deg(int64_t j) const { return (int) (ptr[j + 1] - ptr[j]); }
  inline const int32_t *nb(int64_t j) const { return idx + ptr[j]; }
};

enum Compat : uint32_t { COMPAT_Q1 = 1u, COMPAT_Q3 = 4u };

struct ClusterIO {
  int variant;   // 0 iterate, 1 iterate_vvv, 2 iterate_t, 3 iterate_t_vvv
  const double *Y;
  int64_t n;
  int d, q;
  Csr g;
  int nrep, thin;
  double gamma;
  const int32_t *init;
  const double *mu0, *lambda0;
  double alpha, beta;
  uint64_t seed;
  const int32_t *order;
  uint32_t compat;
  int32_t *z_out;
  double *mu_out, *lambda_out, *w_out, *plogLik;
};

inline double yat(const double *Y, int64_t n, int64_t j, int c) { return Y[(size_t) c * n + j]; }

// One z (and w) scan of the cluster samplers (:272-306, :392-429, :507-552, :646-694) with the parameters of
// iteration i: the body of the reference's `for (j = 0; j < n; j++)` loop, over io.order when given.
// Shared by run_cluster() and by the fixed-parameter scan used to pin the target distribution (orc_scan_fixed).
int scan_zw(const ClusterIO &io, int i, const Rng &rng, std::vector<int32_t> &z, vecd &w, const vecd &mu,
            const vecd &mu_long, const std::vector<vecd> &lambda, const std::vector<vecd> &sigma, vecd &plogLikj) {
  const bool vvv = io.variant == 1 || io.variant == 3, tdist = io.variant >= 2;
  const int64_t n = io.n;
  const int d = io.d, q = io.q;
  const size_t dd = (size_t) d * d;
  const bool q1 = io.compat & COMPAT_Q1, q3 = (io.compat & COMPAT_Q3) && io.variant == 1;
  vecd sigw(dd);
  const double w_alpha = (double) ((d + 4) / 2);   // Q2: integer division
  for (int64_t t = 0; t < n; t++) {
    const int64_t j = io.order ? io.order[t] : t;
    const int z_prev = z[j];
    vecd yj(d);
    for (int c = 0; c < d; c++) yj[c] = yat(io.Y, n, j, c);
    if (tdist) {
      const vecd &lam = lambda[vvv ? z_prev - 1 : 0];
      double quad = 0.0;
      for (int a = 0; a < d; a++) {
        double s = 0.0;
        for (int b = 0; b < d; b++) s += at(lam, d, a, b) * (yj[b] - mu_long[(size_t) j * d + b]);
        quad += (yj[a] - mu_long[(size_t) j * d + a]) * s;
      }
      const double w_beta = 2.0 / (quad + 4.0);
      w[j] = gamma_spot(rng, (uint32_t) j, (uint32_t) i, w_alpha, w_beta, spot_draw(rng, (uint32_t) j, (uint32_t) i).u_first);
    }
    const SpotDraw dr = spot_draw(rng, (uint32_t) j, (uint32_t) i);
    const double u_acc = dr.u_acc;
    const int z_new = propose_label_bits(z_prev, q, dr.prop_bits);
    const int dg = io.g.deg(j);
    const int32_t *nb = io.g.nb(j);
    double h_prev = 0.0, h_new = 0.0;
    if (dg != 0) {
      int c_prev = 0, c_new = 0;
      for (int e = 0; e < dg; e++) {
        if (q3) c_prev += (nb[e] == z_prev);   // Q3 (:404): neighbour INDEX compared to label
        else c_prev += (z[nb[e]] == z_prev);
        c_new += (z[nb[e]] == z_new);
      }
      h_prev = io.gamma / dg * 2 * c_prev;
      h_new = io.gamma / dg * 2 * c_new;
    }
    const vecd *sp = &sigma[vvv ? z_prev - 1 : 0], *sn = &sigma[vvv ? z_new - 1 : 0];
    if (tdist) {
      for (size_t e = 0; e < dd; e++) sigw[e] = (*sp)[e] / w[j];
      h_prev += dmvnrm_arma_fast(yj.data(), &mu[(size_t) (z_prev - 1) * d], sigw, d, q1);
      if (vvv)
        for (size_t e = 0; e < dd; e++) sigw[e] = (*sn)[e] / w[j];
      h_new += dmvnrm_arma_fast(yj.data(), &mu[(size_t) (z_new - 1) * d], sigw, d, q1);
    } else {
      h_prev += dmvnrm_arma_fast(yj.data(), &mu[(size_t) (z_prev - 1) * d], *sp, d, q1);
      h_new += dmvnrm_arma_fast(yj.data(), &mu[(size_t) (z_new - 1) * d], *sn, d, q1);
    }
    double prob = std::exp(h_new - h_prev);
    if (prob > 1) prob = 1;
    if (std::isnan(prob)) return 6;   // Rcpp sample() errors on NA probabilities
    if (sample_two(prob, u_acc)) z[j] = z_new;
    plogLikj[j] = h_prev;
  }
  return 0;
}

// The four cluster samplers share one body; `vvv` and `tdist` select the reference function.
int run_cluster(const ClusterIO &io) {
  const bool vvv = io.variant == 1 || io.variant == 3, tdist = io.variant >= 2;
  const int64_t n = io.n;
  const int d = io.d, q = io.q, nrep = io.nrep, thin = io.thin;
  const int nsnap = nrep / thin + 1, nl = vvv ? q : 1;
  const size_t dd = (size_t) d * d;
  const Rng rng(io.seed);
  const double nan = std::numeric_limits<double>::quiet_NaN();

  vecd lambda0(io.lambda0, io.lambda0 + dd), sigma0;
  if (!inv_general(lambda0, d, sigma0)) return 2;
  std::vector<vecd> lambda(nl, lambda0), sigma(nl, sigma0);
  std::vector<int32_t> z(io.init, io.init + n);
  vecd w((size_t) n, 1.0), mu((size_t) q * d), mu_long((size_t) n * d), plogLikj((size_t) n);

  // row 0 of every output (src/cluster.cpp:232-237, 462-469)
  for (int64_t j = 0; j < n; j++) io.z_out[j] = z[j];
  for (int k = 0; k < q; k++)
    for (int c = 0; c < d; c++) io.mu_out[(size_t) k * d + c] = io.mu0[c];
  for (int k = 0; k < nl; k++) std::memcpy(io.lambda_out + k * dd, io.lambda0, dd * sizeof(double));
  if (tdist && io.w_out)
    for (int64_t j = 0; j < n; j++) io.w_out[j] = 1.0;
  for (int i = 0; i < nrep; i++) io.plogLik[i] = nan;

  vecd lm0(d);
  matvec(lambda0, io.mu0, d, lm0.data());
  vecd P(dd), var, rhs(d), mean(d), ysum(d), S(dd), Sc, VS(dd);

  for (int i = 1; i < nrep; i++) {
    // ---- update mu (:246-258, :359-372, :478-493, :609-625)
    for (int k = 1; k <= q; k++) {
      const vecd &lam_k = lambda[vvv ? k - 1 : 0];
      double n_i = 0.0;
      std::fill(ysum.begin(), ysum.end(), 0.0);
      for (int64_t j = 0; j < n; j++)
        if (z[j] == k) {
          n_i += tdist ? w[j] : 1.0;
          for (int c = 0; c < d; c++) ysum[c] += (tdist ? w[j] : 1.0) * yat(io.Y, n, j, c);
        }
      for (size_t e = 0; e < dd; e++) P[e] = lambda0[e] + n_i * lam_k[e];
      if (!inv_general(P, d, var)) return 3;
      matvec(lam_k, ysum.data(), d, rhs.data());
      for (int c = 0; c < d; c++) rhs[c] += lm0[c];
      matvec(var, rhs.data(), d, mean.data());
      if (!rmvnorm1(rng, P_MU, (uint32_t) (k - 1), (uint32_t) i, mean.data(), var, d,
                    &mu[(size_t) (k - 1) * d]))
        return 3;
    }
    // ---- update lambda (:260-270, :374-390, :495-505, :627-644)
    for (int64_t j = 0; j < n; j++)
      for (int c = 0; c < d; c++) mu_long[(size_t) j * d + c] = mu[(size_t) (z[j] - 1) * d + c];
    for (int kk = 0; kk < nl; kk++) {
      std::fill(S.begin(), S.end(), 0.0);
      int64_t n_k = 0;
      for (int64_t j = 0; j < n; j++) {
        if (vvv && z[j] != kk + 1) continue;
        n_k++;
        const double wj = tdist ? w[j] : 1.0;
        for (int b = 0; b < d; b++) {
          const double rb = wj * (yat(io.Y, n, j, b) - mu_long[(size_t) j * d + b]);
          for (int a = 0; a < d; a++)
            S[(size_t) b * d + a] += (yat(io.Y, n, j, a) - mu_long[(size_t) j * d + a]) * rb;
        }
      }
      for (size_t e = 0; e < dd; e++) VS[e] = S[e];
      for (int c = 0; c < d; c++) at(VS, d, c, c) += io.beta;
      if (!inv_general(VS, d, Sc)) return 4;
      const int df = (int) ((vvv ? (double) n_k : (double) n) + io.alpha);   // Q4: int df
      if (!rwish(rng, (uint32_t) kk, (uint32_t) i, df, Sc, d, lambda[kk])) return 5;
      if (!inv_general(lambda[kk], d, sigma[kk])) return 5;
    }
    // ---- update z (and w) (:272-306, :392-429, :507-552, :646-694)
    {
      const int rc = scan_zw(io, i, rng, z, w, mu, mu_long, lambda, sigma, plogLikj);
      if (rc) return rc;
    }
    double pl = 0.0;
    for (int64_t j = 0; j < n; j++) pl += plogLikj[j];
    io.plogLik[i] = pl;
    // ---- snapshots (:309-314 etc.)
    if ((i + 1) % thin == 0) {
      const size_t r = (size_t) (i + 1) / thin;
      if (r < (size_t) nsnap) {
        std::memcpy(io.mu_out + r * q * d, mu.data(), sizeof(double) * q * d);
        for (int k = 0; k < nl; k++)
          std::memcpy(io.lambda_out + (r * nl + k) * dd, lambda[k].data(), dd * sizeof(double));
        for (int64_t j = 0; j < n; j++) io.z_out[r * n + j] = z[j];
        if (tdist && io.w_out) std::memcpy(io.w_out + r * n, w.data(), sizeof(double) * n);
      }
    }
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Neighbour construction
// ----------------------------------------------------------------------------------------------
// str_split<uword>(str, ",") + Neighbor(string, ",", false): 1-based tokens -> 0-based ids; empty
// tokens skipped (src/utils.h:10-62, src/neighbor.h:67-80).  Returns false on unconvertible text.
bool parse_neighbor_string(const char *s, std::vector<int32_t> &out) {
  out.clear();
  if (!s) return true;   // NA -> empty
  const char *p = s;
  while (*p) {
    if (*p == ',') { p++; continue; }
    if (*p < '0' || *p > '9') return false;
    uint64_t v = 0;
    while (*p >= '0' && *p <= '9') { v = v * 10 + (uint64_t) (*p - '0'); p++; }
    if (*p && *p != ',') return false;
    out.push_back((int32_t) ((int64_t) v - 1));
  }
  return true;
}

// find_subspot_neighbors (src/cluster.cpp:161-215); positions n x 2 column-major.
int find_subspot_neighbors(const Csr &spot, int64_t n0, const double *pos, int64_t n, double dist,
                           std::vector<std::vector<int32_t>> &out) {
  if (n0 <= 0) return 1;
  const int64_t S = n / n0;
  if (std::fabs((double) n / (double) n0 - (double) S) > 1e-6) return 1;   // "Invalid arguments!"
  out.assign((size_t) n, std::vector<int32_t>());
  std::vector<int32_t> spots;
  for (int64_t s = 0; s < n; s++) {
    const int64_t j0 = s % n0;
    spots.assign(spot.nb(j0), spot.nb(j0) + spot.deg(j0));
    spots.push_back((int32_t) j0);
    // candidates: column-major flatten of the S x (|N|+1) matrix r*n0 + p  -> p outer, r inner
    for (size_t pi = 0; pi < spots.size(); pi++)
      for (int64_t r = 0; r < S; r++) {
        const int64_t cnd = r * n0 + spots[pi];
        const double m = std::fabs(pos[cnd] - pos[s]) + std::fabs(pos[n + cnd] - pos[n + s]);
        if (m < dist && m > 1e-6) out[(size_t) s].push_back((int32_t) cnd);
      }
    if (out[(size_t) s].size() < 2) return 2;   // "Error in finding neighbors of subspots!"
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// iterate_deconv (src/cluster.cpp:742-1141)
// ----------------------------------------------------------------------------------------------
// adaptive_mcmc (:49-65): chol(A (I + eta (acc-target) u^T u / |u|^2) A^T, "lower")
bool adaptive_mcmc(double acc, double target, size_t iter, const vecd &A, const double *u, int d,
                   vecd &out) {
  const double step = std::min(1.0, d * std::pow((double) iter, -2.0 / 3));
  double uu = 0.0;
  for (int i = 0; i < d; i++) uu += u[i] * u[i];
  vecd M((size_t) d * d), AM, AMAt((size_t) d * d, 0.0), R;
  for (int j = 0; j < d; j++)
    for (int i = 0; i < d; i++)
      at(M, d, i, j) = (i == j ? 1.0 : 0.0) + step * (acc - target) * u[i] * u[j] / uu;
  matmul(A, M, d, AM);
  for (int j = 0; j < d; j++)
    for (int i = 0; i < d; i++) {
      double s = 0.0;
      for (int k = 0; k < d; k++) s += at(AM, d, i, k) * at(A, d, j, k);
      at(AMAt, d, i, j) = s;
    }
  // lower Cholesky = transpose of the upper factor of the (symmetrised via upper-triangle) matrix
  if (!chol_upper(AMAt, d, R)) return false;
  out.assign((size_t) d * d, 0.0);
  for (int j = 0; j < d; j++)
    for (int i = j; i < d; i++) at(out, d, i, j) = at(R, d, j, i);
  return true;
}

struct DeconvIO {
  const double *pos;   // n x 2 col-major
  double dist;
  Csr spot;            // parsed spot neighbours (n0)
  double *Y;           // n x d col-major, mutated in place (Q10)
  int tdist, nrep, thin;
  int64_t n, n0;
  int d;
  double gamma;
  int q;
  const int32_t *init;
  int subspots;
  double jitter_scale;
  int adapt_before;
  double c;
  const double *mu0, *lambda0;
  double alpha, beta;
  int thread_num;
  uint64_t seed;
  const int32_t *order;   // visiting order over parent spots (n0) or NULL
  int32_t *z_out;
  double *mu_out, *lambda_out, *w_out, *Y_out, *Ychange, *plogLik;
  int64_t *dfj_out;   // n x 4 col-major, 1-based zero padded
};

int run_deconv(const DeconvIO &io) {
  const int64_t n = io.n, n0 = io.n0;
  const int d = io.d, q = io.q, nrep = io.nrep, thin = io.thin, S = io.subspots;
  const size_t dd = (size_t) d * d;
  const int nsnap = nrep / thin + 1;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  const Rng rng(io.seed);
#ifdef _OPENMP
  omp_set_num_threads(io.thread_num > 0 ? io.thread_num : 1);
#endif
  std::vector<std::vector<int32_t>> sub_nb;
  int rc = find_subspot_neighbors(io.spot, n0, io.pos, n, io.dist, sub_nb);
  if (rc) return rc;
  if ((int64_t) S * n0 != n) return 1;

  vecd Y0((size_t) n0 * d);   // Y.rows(0, n0-1) (:777)
  for (int64_t j = 0; j < n0; j++)
    for (int c = 0; c < d; c++) Y0[(size_t) c * n0 + j] = yat(io.Y, n, j, c);
  vecd Y_new((size_t) n * d), acceptance_prob((size_t) n0), ll_prev((size_t) n), ll_new((size_t) n);
  std::vector<uint8_t> ll_sel((size_t) n, 0);
  vecd lambda_i(io.lambda0, io.lambda0 + dd), lambda0(lambda_i);
  std::vector<int32_t> z(io.init, io.init + n);
  vecd w((size_t) n, 1.0), mu((size_t) q * d), mu_long((size_t) n * d), error((size_t) n * d);

  for (int64_t j = 0; j < n; j++) io.z_out[j] = z[j];
  for (int k = 0; k < q; k++)
    for (int c = 0; c < d; c++) io.mu_out[(size_t) k * d + c] = io.mu0[c];
  std::memcpy(io.lambda_out, io.lambda0, dd * sizeof(double));
  if (io.Y_out) std::memcpy(io.Y_out, io.Y, sizeof(double) * n * d);
  for (int64_t j = 0; j < n; j++) io.w_out[j] = 1.0;
  for (int i = 0; i < nrep; i++) io.Ychange[i] = io.plogLik[i] = nan;

  const double w_alpha = (double) ((d + 4) / 2);   // Q2
  const double jitter_scale = io.jitter_scale / d;   // :814
  const bool adaptive = jitter_scale == 0.0;
  std::vector<size_t> num_accepts((size_t) n0, 0), num_rejects((size_t) n0, 0);
  std::vector<vecd> adaptive_mtx;
  if (adaptive) {
    adaptive_mtx.assign((size_t) n, vecd(dd, 0.0));
    for (int64_t s = 0; s < n; s++)
      for (int c = 0; c < d; c++) at(adaptive_mtx[(size_t) s], d, c, c) = 1.0;
  }
  vecd lm0(d);
  matvec(lambda0, io.mu0, d, lm0.data());
  vecd P(dd), var, rhs(d), mean(d), ysum(d), Ssum(dd), VS(dd), Sc, lw(dd);
  const double err_sd = adaptive ? 1.0 : std::sqrt(jitter_scale);

  for (int i = 1; i < nrep; i++) {
    // ---- mu (:883-892)
    for (int k = 1; k <= q; k++) {
      double n_i = 0.0;
      std::fill(ysum.begin(), ysum.end(), 0.0);
      for (int64_t s = 0; s < n; s++)
        if (z[s] == k) {
          n_i += w[s];
          for (int c = 0; c < d; c++) ysum[c] += w[s] * yat(io.Y, n, s, c);
        }
      for (size_t e = 0; e < dd; e++) P[e] = lambda0[e] + n_i * lambda_i[e];
      if (!inv_general(P, d, var)) return 3;
      matvec(lambda_i, ysum.data(), d, rhs.data());
      for (int c = 0; c < d; c++) rhs[c] += lm0[c];
      matvec(var, rhs.data(), d, mean.data());
      if (!rmvnorm1(rng, P_MU, (uint32_t) (k - 1), (uint32_t) i, mean.data(), var, d,
                    &mu[(size_t) (k - 1) * d]))
        return 3;
    }
    // ---- lambda (:896-902)
    for (int64_t s = 0; s < n; s++)
      for (int c = 0; c < d; c++) mu_long[(size_t) s * d + c] = mu[(size_t) (z[s] - 1) * d + c];
    std::fill(Ssum.begin(), Ssum.end(), 0.0);
    for (int64_t s = 0; s < n; s++)
      for (int b = 0; b < d; b++) {
        const double rb = w[s] * (yat(io.Y, n, s, b) - mu_long[(size_t) s * d + b]);
        for (int a = 0; a < d; a++)
          Ssum[(size_t) b * d + a] += (yat(io.Y, n, s, a) - mu_long[(size_t) s * d + a]) * rb;
      }
    for (size_t e = 0; e < dd; e++) VS[e] = Ssum[e];
    for (int c = 0; c < d; c++) at(VS, d, c, c) += io.beta;
    if (!inv_general(VS, d, Sc)) return 4;
    if (!rwish(rng, 0u, (uint32_t) i, (int) ((double) n + io.alpha), Sc, d, lambda_i)) return 5;
    // ---- jitter draws (:905-908): rmvnorm(n, 0, I*jitter_scale) = Z * sqrt(jitter_scale)
    for (int64_t s = 0; s < n; s++)
      for (int c = 0; c < d; c++)
        error[(size_t) c * n + s] = err_sd * rng.normal(P_ERR, (uint32_t) s, (uint32_t) i, (uint32_t) c);
    // ---- per-spot proposal + acceptance probability (:912-962)
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t j0 = 0; j0 < n0; j0++) {
      vecd err_j((size_t) S * d), emean(d, 0.0), lwl(dd), yr(d), ynew(d);
      for (int r = 0; r < S; r++) {
        const int64_t s = (int64_t) r * n0 + j0;
        for (int c = 0; c < d; c++) err_j[(size_t) r * d + c] = error[(size_t) c * n + s];
        if (adaptive) {   // error_j.row(r) = (A * e^T)^T
          vecd tmp(d);
          matvec(adaptive_mtx[(size_t) s], &err_j[(size_t) r * d], d, tmp.data());
          for (int c = 0; c < d; c++) err_j[(size_t) r * d + c] = tmp[c];
        }
      }
      for (int r = 0; r < S; r++)
        for (int c = 0; c < d; c++) emean[c] += err_j[(size_t) r * d + c];
      for (int c = 0; c < d; c++) emean[c] /= S;
      double p_prev = 0.0, p_new = 0.0;
      for (int r = 0; r < S; r++) {
        const int64_t s = (int64_t) r * n0 + j0;
        double pen_prev = 0.0, pen_new = 0.0;
        for (int c = 0; c < d; c++) {
          yr[c] = yat(io.Y, n, s, c);
          ynew[c] = yr[c] + (err_j[(size_t) r * d + c] - emean[c]);
          Y_new[(size_t) c * n + s] = ynew[c];
          const double y0 = Y0[(size_t) c * n0 + j0];
          pen_prev += (yr[c] - y0) * (yr[c] - y0);
          pen_new += (ynew[c] - y0) * (ynew[c] - y0);
        }
        for (size_t e = 0; e < dd; e++) lwl[e] = lambda_i[e] * w[s];
        const double a = dmvnrm_prec_arma_fast(yr.data(), &mu_long[(size_t) s * d], lwl, d);
        const double b = dmvnrm_prec_arma_fast(ynew.data(), &mu_long[(size_t) s * d], lwl, d);
        if (std::isnan(a) || std::isnan(b)) bad |= 1;
        p_prev += a - io.c * pen_prev;
        p_new += b - io.c * pen_new;
      }
      double pr = std::exp(p_new - p_prev);
      if (pr > 1) pr = 1;
      acceptance_prob[(size_t) j0] = pr;
    }
    if (bad) return 6;
    // ---- accept Y; update w and z (:964-1053)
    int updateCounter = 0;
    vecd ys(d);
    for (int64_t t = 0; t < n0; t++) {
      const int64_t j0 = io.order ? io.order[t] : t;
      double u_y, u_unused;
      rng.uniform2(P_SPOT_YACC, (uint32_t) j0, (uint32_t) i, 0, u_y, u_unused);
      if (std::isnan(acceptance_prob[(size_t) j0])) return 6;
      const bool yes = sample_two(acceptance_prob[(size_t) j0], u_y);
      if (yes) {
        for (int r = 0; r < S; r++) {
          const int64_t s = (int64_t) r * n0 + j0;
          for (int c = 0; c < d; c++) io.Y[(size_t) c * n + s] = Y_new[(size_t) c * n + s];
        }
        num_accepts[(size_t) j0]++;
        updateCounter++;
      } else
        num_rejects[(size_t) j0]++;
      for (int r = 0; r < S; r++) {
        const int64_t s = (int64_t) r * n0 + j0;
        if (adaptive && i > 10 && (io.adapt_before == 0 || i <= io.adapt_before)) {
          vecd raw(d), upd;
          for (int c = 0; c < d; c++) raw[c] = error[(size_t) c * n + s];
          const double acc = (double) num_accepts[(size_t) j0] /
                             (double) (num_accepts[(size_t) j0] + num_rejects[(size_t) j0]);
          if (!adaptive_mcmc(acc, 0.234, (size_t) i, adaptive_mtx[(size_t) s], raw.data(), d, upd))
            return 7;
          adaptive_mtx[(size_t) s] = upd;
        }
        for (int c = 0; c < d; c++) ys[c] = yat(io.Y, n, s, c);
        if (io.tdist) {
          double quad = 0.0;
          for (int a = 0; a < d; a++) {
            double sacc = 0.0;
            for (int b = 0; b < d; b++)
              sacc += at(lambda_i, d, a, b) * (ys[b] - mu_long[(size_t) s * d + b]);
            quad += (ys[a] - mu_long[(size_t) s * d + a]) * sacc;
          }
          w[s] = rng.gamma(P_SPOT_GAMMA, (uint32_t) s, (uint32_t) i, w_alpha, 2.0 / (quad + 4.0));
        }
        double u_prop, u_acc;
        rng.uniform2(P_SPOT_U, (uint32_t) s, (uint32_t) i, 0, u_prop, u_acc);
        const int z_prev = z[s], z_new = propose_label(z_prev, q, u_prop);
        for (size_t e = 0; e < dd; e++) lw[e] = lambda_i[e] * w[s];
        const double l_prev = dmvnrm_prec_arma_fast(ys.data(), &mu[(size_t) (z_prev - 1) * d], lw, d);
        const double l_new = dmvnrm_prec_arma_fast(ys.data(), &mu[(size_t) (z_new - 1) * d], lw, d);
        double pr_prev = 0.0, pr_new = 0.0;
        const std::vector<int32_t> &nb = sub_nb[(size_t) s];
        if (!nb.empty()) {
          int c_prev = 0, c_new = 0;
          for (size_t e = 0; e < nb.size(); e++) {
            c_prev += (z[nb[e]] == z_prev);
            c_new += (z[nb[e]] == z_new);
          }
          pr_prev = io.gamma / nb.size() * 2 * c_prev;
          pr_new = io.gamma / nb.size() * 2 * c_new;
        }
        ll_prev[(size_t) s] = l_prev;
        ll_new[(size_t) s] = l_new;
        const double prob = std::exp((l_new + pr_new) - (l_prev + pr_prev));
        if (std::isnan(prob)) return 6;
        bool upd;
        if (prob >= 1) upd = true;
        else upd = sample_two(prob, u_acc);
        ll_sel[(size_t) s] = upd ? 1 : 0;
        if (upd) z[s] = z_new;
      }
    }
    io.Ychange[i] = updateCounter * 1.0 / n0;
    double pl = 0.0;
    for (int64_t s = 0; s < n; s++) pl += ll_sel[(size_t) s] ? ll_new[(size_t) s] : ll_prev[(size_t) s];
    io.plogLik[i] = pl;
    if ((i + 1) % thin == 0) {
      const size_t r = (size_t) (i + 1) / thin;
      if (r < (size_t) nsnap) {
        std::memcpy(io.mu_out + r * q * d, mu.data(), sizeof(double) * q * d);
        std::memcpy(io.lambda_out + r * dd, lambda_i.data(), dd * sizeof(double));
        if (io.Y_out) std::memcpy(io.Y_out + r * (size_t) n * d, io.Y, sizeof(double) * n * d);
        std::memcpy(io.w_out + r * n, w.data(), sizeof(double) * n);
        for (int64_t s = 0; s < n; s++) io.z_out[r * n + s] = z[s];
      }
    }
  }
  // convert_neighbors2mtx (:147-159): n x 4, 1-based, zero padded (Q8: degree > 4 unsupported)
  if (io.dfj_out) {
    for (int64_t e = 0; e < n * 4; e++) io.dfj_out[e] = 0;
    for (int64_t s = 0; s < n; s++) {
      if (sub_nb[(size_t) s].size() > 4) return 8;
      for (size_t e = 0; e < sub_nb[(size_t) s].size(); e++)
        io.dfj_out[(size_t) e * n + s] = (int64_t) sub_nb[(size_t) s][e] + 1;
    }
  }
  return 0;
}

}   // namespace

// ================================================================================================
// C entry points (ctypes)
// ================================================================================================
extern "C" {

void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  philox4x32_10(ctr, key, out);
}
void orc_uniform2(uint64_t seed, uint32_t purpose, uint32_t index, uint32_t iter, uint32_t slot,
                  double *out) {
  Rng(seed).uniform2(purpose, index, iter, slot, out[0], out[1]);
}
void orc_normal(uint64_t seed, uint32_t purpose, uint32_t index, uint32_t iter, uint32_t m0,
                int count, double *out) {
  Rng r(seed);
  for (int m = 0; m < count; m++) out[m] = r.normal(purpose, index, iter, m0 + (uint32_t) m);
}
// (prop_bits, u_acc, u_first) of one spot and iteration; out[0] holds prop_bits as a double
void orc_spot_draw(uint64_t seed, uint32_t index, uint32_t iter, double *out) {
  const SpotDraw d = spot_draw(Rng(seed), index, iter);
  out[0] = (double) d.prop_bits; out[1] = d.u_acc; out[2] = d.u_first;
}
double orc_gamma_spot(uint64_t seed, uint32_t index, uint32_t iter, double shape, double scale) {
  const Rng rng(seed);
  return gamma_spot(rng, index, iter, shape, scale, spot_draw(rng, index, iter).u_first);
}
int orc_propose_label_bits(int z_prev, int q, uint32_t bits) { return propose_label_bits(z_prev, q, bits); }

void orc_gamma(uint64_t seed, uint32_t purpose, uint32_t index0, int count, uint32_t iter,
               double shape, double scale, double *out) {
  Rng r(seed);
  for (int m = 0; m < count; m++) out[m] = r.gamma(purpose, index0 + (uint32_t) m, iter, shape, scale);
}

// x: n x d col-major; mean: d; sigma: d x d.  form 0: covariance form (dmvnrm_arma_fast),
// form 1: precision form (dmvnrm_prec_arma_fast).  w (optional): sigma/w[j] resp. lambda*w[j].
void orc_dmvnrm(int form, const double *x, int64_t n, int d, const double *mean, const double *mat,
                const double *w, int q1_compat, double *out) {
  vecd M(mat, mat + (size_t) d * d), Mw((size_t) d * d), xj(d);
  for (int64_t j = 0; j < n; j++) {
    for (int c = 0; c < d; c++) xj[c] = x[(size_t) c * n + j];
    const vecd *use = &M;
    if (w) {
      for (size_t e = 0; e < M.size(); e++) Mw[e] = form == 0 ? M[e] / w[j] : M[e] * w[j];
      use = &Mw;
    }
    out[j] = form == 0 ? dmvnrm_arma_fast(xj.data(), mean, *use, d, q1_compat != 0)
                       : dmvnrm_prec_arma_fast(xj.data(), mean, *use, d);
  }
}

int orc_chol_upper(const double *X, int d, double *R) {
  vecd x(X, X + (size_t) d * d), r;
  bool ok = chol_upper(x, d, r);
  std::memcpy(R, r.data(), sizeof(double) * d * d);
  return ok ? 0 : 1;
}
int orc_inv(const double *X, int d, double *Xi) {
  vecd x(X, X + (size_t) d * d), r;
  bool ok = inv_general(x, d, r);
  std::memcpy(Xi, r.data(), sizeof(double) * d * d);
  return ok ? 0 : 1;
}

// Sufficient statistics on a fixed state: n_k (sum of w over cluster), s_k (q x d row-major),
// pooled scatter S (d x d) and per-cluster scatter S_k (q x d x d), all about mu_{z_j}.
void orc_suffstats(const double *Y, int64_t n, int d, int q, const int32_t *z, const double *w,
                   const double *mu, double *nk, double *sk, double *S, double *Sk) {
  std::fill(nk, nk + q, 0.0);
  std::fill(sk, sk + (size_t) q * d, 0.0);
  std::fill(S, S + (size_t) d * d, 0.0);
  std::fill(Sk, Sk + (size_t) q * d * d, 0.0);
  vecd r(d);
  for (int64_t j = 0; j < n; j++) {
    const int k = z[j] - 1;
    const double wj = w ? w[j] : 1.0;
    nk[k] += wj;
    for (int c = 0; c < d; c++) {
      sk[(size_t) k * d + c] += wj * yat(Y, n, j, c);
      r[c] = yat(Y, n, j, c) - mu[(size_t) k * d + c];
    }
    for (int b = 0; b < d; b++)
      for (int a = 0; a < d; a++) {
        const double v = wj * r[a] * r[b];
        S[(size_t) b * d + a] += v;
        Sk[((size_t) k * d + b) * d + a] += v;
      }
  }
}

// h(z_prev), h(z_new) for given proposals on a fixed state, exactly as the z-scan evaluates them.
// sigma: (vvv ? q : 1) x d x d covariance matrices.  variant as in orc_iterate.
int orc_h_values(int variant, uint32_t compat, const double *Y, int64_t n, int d, int q,
                 const int64_t *nbr_ptr, const int32_t *nbr_idx, const int32_t *z, const double *w,
                 const double *mu, const double *sigma, double gamma, const int32_t *z_new,
                 double *h_prev, double *h_new) {
  const bool vvv = variant == 1 || variant == 3, tdist = variant >= 2;
  const bool q1 = compat & COMPAT_Q1, q3 = (compat & COMPAT_Q3) && variant == 1;
  const size_t dd = (size_t) d * d;
  Csr g{nbr_ptr, nbr_idx};
  vecd yj(d), sg(dd);
  for (int64_t j = 0; j < n; j++) {
    for (int c = 0; c < d; c++) yj[c] = yat(Y, n, j, c);
    const int zp = z[j], zn = z_new[j], dg = g.deg(j);
    double hp = 0.0, hn = 0.0;
    if (dg) {
      int cp = 0, cn = 0;
      for (int e = 0; e < dg; e++) {
        const int32_t u = g.nb(j)[e];
        cp += q3 ? (u == zp) : (z[u] == zp);
        cn += (z[u] == zn);
      }
      hp = gamma / dg * 2 * cp;
      hn = gamma / dg * 2 * cn;
    }
    const double wj = tdist ? w[j] : 1.0;
    const double *sp = sigma + (vvv ? (size_t) (zp - 1) * dd : 0), *sn = sigma + (vvv ? (size_t) (zn - 1) * dd : 0);
    for (size_t e = 0; e < dd; e++) sg[e] = tdist ? sp[e] / wj : sp[e];
    hp += dmvnrm_arma_fast(yj.data(), mu + (size_t) (zp - 1) * d, sg, d, q1);
    for (size_t e = 0; e < dd; e++) sg[e] = tdist ? sn[e] / wj : sn[e];
    hn += dmvnrm_arma_fast(yj.data(), mu + (size_t) (zn - 1) * d, sg, d, q1);
    h_prev[j] = hp;
    h_new[j] = hn;
  }
  return 0;
}

// niter z (and w) scans with FIXED mu / lambda (sigma = inverse given by the caller): the Markov kernel of the label
// update alone.  hist[c] counts how often configuration c = sum_j (z_j - 1) q^j was seen after a scan (n small).
int orc_scan_fixed(int variant, uint32_t compat, const double *Y, int64_t n, int d, int q, const int64_t *nbr_ptr,
                   const int32_t *nbr_idx, double gamma, const int32_t *init, const double *mu_in,
                   const double *lambda_in, const double *sigma_in, uint64_t seed, const int32_t *order, int niter,
                   int64_t *hist, double *w_mean) {
  const bool vvv = variant == 1 || variant == 3;
  const int nl = vvv ? q : 1;
  const size_t dd = (size_t) d * d;
  ClusterIO io{variant, Y, n, d, q, Csr{nbr_ptr, nbr_idx}, niter + 1, 1, gamma, init, nullptr, nullptr, 0.0, 0.0, seed, order,
               compat, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<int32_t> z(init, init + n);
  vecd w((size_t) n, 1.0), mu(mu_in, mu_in + (size_t) q * d), mu_long((size_t) n * d), plogLikj((size_t) n);
  std::vector<vecd> lambda(nl), sigma(nl);
  for (int k = 0; k < nl; k++) {
    lambda[k].assign(lambda_in + k * dd, lambda_in + (k + 1) * dd);
    sigma[k].assign(sigma_in + k * dd, sigma_in + (k + 1) * dd);
  }
  const Rng rng(seed);
  try {
    for (int i = 1; i <= niter; i++) {
      for (int64_t j = 0; j < n; j++)
        for (int c = 0; c < d; c++) mu_long[(size_t) j * d + c] = mu[(size_t) (z[j] - 1) * d + c];
      const int rc = scan_zw(io, i, rng, z, w, mu, mu_long, lambda, sigma, plogLikj);
      if (rc) return rc;
      int64_t code = 0, mult = 1;
      for (int64_t j = 0; j < n; j++) { code += (z[j] - 1) * mult; mult *= q; }
      hist[code]++;
      if (w_mean)
        for (int64_t j = 0; j < n; j++) w_mean[j] += w[j] / niter;
    }
  } catch (...) {
    return 99;
  }
  return 0;
}

// map_subspot2ref (src/compare_enhanced.cpp:47-92): per subspot i the reference rows j with dists == min(dists),
// dists = sum over columns of (sub[i, c] - ref[j, c])^2 in column order.  out: rows (i, j, min) appended in i order
// (the reference's single-thread order); returns the number of rows, or -1 if cap is too small.
int64_t orc_map_subspot2ref(const double *sub, int64_t ns, const double *ref, int64_t nr, int k, double *out, int64_t cap) {
  int64_t rows = 0;
  std::vector<double> dists((size_t) nr);
  for (int64_t i = 0; i < ns; i++) {
    double mn = std::numeric_limits<double>::infinity();
    for (int64_t j = 0; j < nr; j++) {
      double s = 0.0;
      for (int c = 0; c < k; c++) {
        const volatile double v = sub[(size_t) c * ns + i] - ref[(size_t) c * nr + j];
        const volatile double v2 = v * v;   // separately rounded, as std::pow(v, 2) then arma::sum
        s += v2;
      }
      dists[(size_t) j] = s;
      if (s < mn) mn = s;
    }
    for (int64_t j = 0; j < nr; j++)
      if (dists[(size_t) j] == mn) {
        if (rows >= cap) return -1;
        out[rows * 3] = (double) i; out[rows * 3 + 1] = (double) j; out[rows * 3 + 2] = mn;
        rows++;
      }
  }
  return rows;
}

// compute_corr (src/compare_enhanced.cpp:94-132): arma::cor of row i of m1 and of m2 (n x p col-major)
void orc_compute_corr(const double *m1, const double *m2, int64_t n, int64_t p, double *cor) {
  for (int64_t i = 0; i < n; i++) {
    double ma = 0.0, mb = 0.0;
    for (int64_t c = 0; c < p; c++) { ma += m1[(size_t) c * n + i]; mb += m2[(size_t) c * n + i]; }
    ma /= (double) p; mb /= (double) p;
    double sab = 0.0, saa = 0.0, sbb = 0.0;
    for (int64_t c = 0; c < p; c++) {
      const double a = m1[(size_t) c * n + i] - ma, b = m2[(size_t) c * n + i] - mb;
      sab += a * b; saa += a * a; sbb += b * b;
    }
    cor[i] = sab / std::sqrt(saa * sbb);
  }
}

int orc_iterate(int variant, const double *Y, int64_t n, int d, int q, const int64_t *nbr_ptr,
                const int32_t *nbr_idx, int nrep, int thin, double gamma, const int32_t *init,
                const double *mu0, const double *lambda0, double alpha, double beta, uint64_t seed,
                const int32_t *order, uint32_t compat, int32_t *z_out, double *mu_out,
                double *lambda_out, double *w_out, double *plogLik) {
  ClusterIO io{variant, Y,     n,    d,       q,     Csr{nbr_ptr, nbr_idx}, nrep,       thin,
               gamma,   init,  mu0,  lambda0, alpha, beta,                  seed,       order,
               compat,  z_out, mu_out, lambda_out, w_out, plogLik};
  try {
    return run_cluster(io);
  } catch (...) {
    return 99;
  }
}

int orc_iterate_deconv(const double *pos, double dist, const int64_t *spot_ptr,
                       const int32_t *spot_idx, double *Y, int tdist, int nrep, int thin, int64_t n,
                       int64_t n0, int d, double gamma, int q, const int32_t *init, int subspots,
                       double jitter_scale, int adapt_before, double c, const double *mu0,
                       const double *lambda0, double alpha, double beta, int thread_num,
                       uint64_t seed, const int32_t *order, int32_t *z_out, double *mu_out,
                       double *lambda_out, double *w_out, double *Y_out, double *Ychange,
                       double *plogLik, int64_t *dfj_out) {
  DeconvIO io{pos,  dist, Csr{spot_ptr, spot_idx}, Y,     tdist,   nrep,  thin,  n,
              n0,   d,    gamma,  q,       init,  subspots, jitter_scale, adapt_before,
              c,    mu0,  lambda0, alpha,  beta,  thread_num, seed, order,
              z_out, mu_out, lambda_out, w_out, Y_out, Ychange, plogLik, dfj_out};
  try {
    return run_deconv(io);
  } catch (...) {
    return 99;
  }
}

// ---- neighbours -------------------------------------------------------------------------------
// Parses n0 "a,b,c" strings (NULL = NA).  Two-pass: idx == NULL returns the count only.
int orc_parse_neighbors(const char *const *strs, int64_t n0, int64_t *ptr, int32_t *idx) {
  std::vector<int32_t> tmp;
  int64_t nnz = 0;
  for (int64_t j = 0; j < n0; j++) {
    if (!parse_neighbor_string(strs[j], tmp)) return 1;
    ptr[j] = nnz;
    if (idx)
      for (size_t e = 0; e < tmp.size(); e++) idx[nnz + (int64_t) e] = tmp[e];
    nnz += (int64_t) tmp.size();
  }
  ptr[n0] = nnz;
  return 0;
}

// find_subspot_neighbors; out_idx capacity must be >= 4*... callers pass cap; returns 0/1/2, -1 if cap short.
int orc_subspot_neighbors(const int64_t *spot_ptr, const int32_t *spot_idx, int64_t n0,
                          const double *pos, int64_t n, double dist, int64_t *ptr, int32_t *idx,
                          int64_t cap) {
  std::vector<std::vector<int32_t>> out;
  int rc = find_subspot_neighbors(Csr{spot_ptr, spot_idx}, n0, pos, n, dist, out);
  if (rc) return rc;
  int64_t nnz = 0;
  for (int64_t s = 0; s < n; s++) {
    ptr[s] = nnz;
    for (size_t e = 0; e < out[(size_t) s].size(); e++) {
      if (nnz >= cap) return -1;
      idx[nnz++] = out[(size_t) s][e];
    }
  }
  ptr[n] = nnz;
  return 0;
}

// .find_neighbors (R/spatialCluster.R:227-302): platform 0 = Visium (hex), 1 = VisiumHD/ST (square).
// Neighbours ascending by spot index.  idx capacity 6n.
int orc_lattice_neighbors(const int32_t *array_row, const int32_t *array_col, int64_t n, int platform,
                          int64_t *ptr, int32_t *idx) {
  static const int hx[6] = {-2, 2, -1, 1, -1, 1}, hy[6] = {0, 0, -1, -1, 1, 1};
  static const int sx[4] = {0, 1, 0, -1}, sy[4] = {-1, 0, 1, 0};
  const int no = platform == 0 ? 6 : 4;
  const int *ox = platform == 0 ? hx : sx, *oy = platform == 0 ? hy : sy;
  // merge(by = (x.pos, y.pos)) -> a sorted coordinate table searched by binary search
  std::vector<std::pair<std::pair<int32_t, int32_t>, int32_t>> tab((size_t) n);
  for (int64_t j = 0; j < n; j++) tab[(size_t) j] = {{array_col[j], array_row[j]}, (int32_t) j};
  std::sort(tab.begin(), tab.end());
  int64_t nnz = 0;
  std::vector<int32_t> nb;
  for (int64_t j = 0; j < n; j++) {
    nb.clear();
    for (int o = 0; o < no; o++) {
      std::pair<int32_t, int32_t> key{array_col[j] + ox[o], array_row[j] + oy[o]};
      auto lo = std::lower_bound(tab.begin(), tab.end(), std::make_pair(key, (int32_t) -1));
      for (; lo != tab.end() && lo->first == key; ++lo) nb.push_back(lo->second);
    }
    std::sort(nb.begin(), nb.end());
    ptr[j] = nnz;
    for (size_t e = 0; e < nb.size(); e++) idx[nnz++] = nb[e];
  }
  ptr[n] = nnz;
  return 0;
}

// find_neighbors (R/utils.R:13-26): 0 < dist <= radius; metric 0 manhattan, 1 euclidean.
// Two-pass (idx == NULL counts).  O(n^2) like the reference.
int64_t orc_radius_neighbors(const double *pos, int64_t n, double radius, int metric, int64_t *ptr,
                             int32_t *idx) {
  int64_t nnz = 0;
  for (int64_t i = 0; i < n; i++) {
    ptr[i] = nnz;
    for (int64_t j = 0; j < n; j++) {
      const double dx = pos[i] - pos[j], dy = pos[n + i] - pos[n + j];
      const double dist = metric == 0 ? std::fabs(dx) + std::fabs(dy) : std::sqrt(dx * dx + dy * dy);
      if (dist <= radius && dist > 0) {
        if (idx) idx[nnz] = (int32_t) j;
        nnz++;
      }
    }
  }
  ptr[n] = nnz;
  return nnz;
}

int orc_adaptive_mcmc(double acc, double target, int64_t iter, const double *A, const double *u,
                      int d, double *out) {
  vecd a(A, A + (size_t) d * d), o;
  if (!adaptive_mcmc(acc, target, (size_t) iter, a, u, d, o)) return 1;
  std::memcpy(out, o.data(), sizeof(double) * d * d);
  return 0;
}

}   // extern "C"
