// rng.cuh -- counter-based randomness of the samplers (device side).
//
// Every variate is a pure function of (seed; purpose, index, iteration, slot): Philox4x32-10 with
// counter = (index, slot, iteration, purpose) and key = seed.  A draw therefore belongs to "spot j
// at iteration i", not to a position in a serial stream, which makes chains independent of the
// colour schedule's launch geometry and of the number of GPUs (BASELINE.json north_star).
// The reference draws from R's global Mersenne-Twister under Rcpp::RNGScope
// (src/RcppExports.cpp:19); stream compatibility with it is not a goal.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsb {

enum Purpose : uint32_t {
  P_MU = 1,          // index = cluster k; normals m = 0..d-1            (rmvnorm, src/cluster.cpp:257)
  P_WISH_NORM = 2,   // index = cluster k (0 if shared); m = i(i-1)/2+j  (rwish off-diagonals, :269)
  P_WISH_CHI = 3,    // index = k*1024 + i; gamma attempts               (rwish diagonal)
  P_SPOT_U = 4,      // index = spot; slot 0 = (u_proposal, u_accept)    (sample(), :278,304)
  P_SPOT_GAMMA = 5,  // index = spot; gamma attempts for w               (R::rgamma, :518)
  P_ERR = 6,         // index = subspot; normal m = column               (rmvnorm, :905)
  P_SPOT_YACC = 7    // index = parent spot; slot 0 = u for the Y move   (sample(), :975)
};

struct Key {
  uint32_t k0, k1;
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               Key key) {
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
  const unsigned long long x = ((unsigned long long) hi << 32) | lo;
  return ((double) (x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void rng_uniform2(Key key, uint32_t purpose, uint32_t index,
                                             uint32_t iter, uint32_t slot, double &u0, double &u1) {
  const uint4 o = philox4x32_10(index, slot, iter, purpose, key);
  u0 = u53(o.x, o.y);
  u1 = u53(o.z, o.w);
}

// Box-Muller pair.
__device__ __forceinline__ void rng_normal2(Key key, uint32_t purpose, uint32_t index, uint32_t iter,
                                            uint32_t slot, double &n0, double &n1) {
  double u0, u1;
  rng_uniform2(key, purpose, index, iter, slot, u0, u1);
  const double r = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ double rng_normal(Key key, uint32_t purpose, uint32_t index, uint32_t iter,
                                             uint32_t m) {
  double a, b;
  rng_normal2(key, purpose, index, iter, m >> 1, a, b);
  return (m & 1u) ? b : a;
}

// First Box-Muller output only (the gamma sampler needs one normal per attempt).
__device__ __forceinline__ double rng_normal_first(Key key, uint32_t purpose, uint32_t index,
                                                   uint32_t iter, uint32_t slot) {
  double u0, u1;
  rng_uniform2(key, purpose, index, iter, slot, u0, u1);
  return sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);
}

// Gamma(shape, scale) by Marsaglia & Tsang (2000); shape < 1 boosted by u^(1/shape).
// Stand-in for R::rgamma (src/cluster.cpp:518,660,1007) and R::rchisq inside RcppDist::rwish.
__device__ __forceinline__ double rng_gamma(Key key, uint32_t purpose, uint32_t index, uint32_t iter,
                                            double shape, double scale) {
  if (!(shape > 0.0) || !(scale > 0.0)) return __longlong_as_double(0x7ff8000000000000ll);
  double boost = 1.0, a = shape;
  if (a < 1.0) {
    double u0, u1;
    rng_uniform2(key, purpose, index, iter, 0x80000000u, u0, u1);
    boost = pow(u0, 1.0 / a);
    a += 1.0;
  }
  const double dd = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
  for (uint32_t t = 0; t < 0x3fffffffu; t++) {
    const double x = rng_normal_first(key, purpose, index, iter, 2u * t);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u, u_unused;
    rng_uniform2(key, purpose, index, iter, 2u * t + 1u, u, u_unused);
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return dd * v * scale * boost;
    if (log(u) < 0.5 * x2 + dd * (1.0 - v + log(v))) return dd * v * scale * boost;
  }
  return __longlong_as_double(0x7ff8000000000000ll);
}

// Rcpp sugar sample(x, 1, TRUE, {1-p, p}): probabilities sorted descending, cumulative inversion
// with one uniform.  True when the second element ("new") is picked.
__device__ __forceinline__ bool sample_two(double p_new, double u) {
  return (p_new >= 0.5) ? (u <= p_new) : (u > 1.0 - p_new);
}

// Rcpp sugar sample(qlessk, 1): x[floor(len*u)] over the labels without the current one, ascending.
// Labels here are 0-based.
__device__ __forceinline__ int propose_label(int z_prev, int q, double u) {
  int t = (int) floor((double) (q - 1) * u);
  if (t > q - 2) t = q - 2;
  return (t < z_prev) ? t : t + 1;
}

// ---- per-spot draws of the cluster samplers' z / w scan --------------------------------------------------------
// ONE Philox block per spot and iteration, counter (spot, 0, iteration, P_SPOT_U), serves the three uniforms of
// the scan: word 0 -> the proposal (sample(qlessk, 1), src/cluster.cpp:278), words 1..2 -> the 53-bit uniform of
// the accept step (sample(.., prob), :304), word 3 -> the uniform of the first Marsaglia-Tsang attempt of the
// t-model weight (R::rgamma, :518).  The normal of attempt t comes from block (spot, t, iteration, P_SPOT_GAMMA);
// attempts t >= 1 (rare) take their uniform from word 0 of block (spot, 0x40000000 + t, ..).
__device__ __forceinline__ double u32c(uint32_t x) { return ((double) x + 0.5) * (1.0 / 4294967296.0); }

__device__ __forceinline__ void rng_spot(Key key, uint32_t index, uint32_t iter, uint32_t &prop_bits, double &u_acc,
                                         double &u_sq) {
  const uint4 o = philox4x32_10(index, 0u, iter, P_SPOT_U, key);
  prop_bits = o.x;
  u_acc = u53(o.y, o.z);
  u_sq = u32c(o.w);
}

// sample(qlessk, 1) from 32 random bits: element floor((q-1) * bits / 2^32) of the labels without the current one.
__device__ __forceinline__ int propose_label_bits(int z_prev, int q, uint32_t bits) {
  const int t = (int) __umulhi(bits, (uint32_t) (q - 1));
  return (t < z_prev) ? t : t + 1;
}

// Accept step of sample({z, z'}, 1, TRUE, {1 - p, p}) with p = min(1, exp(delta)), decided in the log domain:
// with v = u (p >= 1/2) or 1 - u (p < 1/2) the rule of sample_two() reads log v <= delta resp. log v < delta.
// A single-precision logarithm decides whenever its error bound cannot change the answer; the (rare) ambiguous
// cases evaluate exp(delta) and sample_two() exactly as the reference's arithmetic does.  Branch-free up to that
// fallback, so the caller's surrounding code stays one basic block.
__device__ __forceinline__ bool mh_accept(double delta, double u) {
  const double v = (delta >= -0.69314718055994530942) ? u : 1.0 - u;
  const float lf = __logf((float) v);
  const float gap = (float) (delta - (double) lf);   // the margin test itself runs in single precision (FMA pipe)
  const bool decisive = fabsf(gap) > 1e-4f * (1.0f + fabsf(lf));
  bool acc = (delta >= 0.0) | (decisive & (gap > 0.0f));
  if (!decisive && delta < 0.0) acc = sample_two(exp(delta), u);
  return acc;
}

// ---- natural logarithm of a positive normal double from a 128-entry table (in shared memory) ------------------
// m = mantissa scaled to [0.75, 1.5), c = centre of the 2^-8-wide bin of m (exactly representable: the leading bits
// of m), log m = log c + log1p((m - c) / c) with a degree-7 polynomial (|(m-c)/c| < 2^-7.9).  About 25 instructions
// without a branch, against ~60 with special-case branches for log(); accurate to ~1e-16 relative (the oracle uses
// libm's log; the t-model weights are compared at 1e-9).  tab[i] = {1 / c_i, log c_i}.
__device__ __forceinline__ void log_table_entry(int i, double2 &e) {
  const int chi = 0x3fe80000 + (i << 13) + 0x1000;   // high word of the bin centre
  const double c = __hiloint2double(chi, 0);
  e.x = 1.0 / c;
  e.y = log(c);
}
__device__ __forceinline__ double log_tab(double u, const double2 *__restrict__ tab) {
  const int hi = __double2hiint(u), lo = __double2loint(u);
  int e = (hi >> 20) - 1023;
  int mh = (hi & 0x000fffff) | 0x3ff00000;
  const bool up = mh >= 0x3ff80000;   // m >= 1.5: halve it
  mh = up ? mh - 0x00100000 : mh;
  e += up ? 1 : 0;
  const double m = __hiloint2double(mh, lo);
  const double c = __hiloint2double((mh & 0xffffe000) | 0x1000, 0);
  const double2 t = tab[(mh - 0x3fe80000) >> 13];
  const double r = (m - c) * t.x;
  double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  p = fma(p, r, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double l1p = fma(p, r * r, r);
  const double ef = (double) e;
  return fma(ef, 0.693147180369123816490, t.y) + fma(ef, 1.90821492927058770002e-10, l1p);
}

// First Box-Muller output from one Philox block, logarithm from the table.
// Its two other factors, written out so that the sweep's straight-line block has neither the table loads of cospi()
// nor the range checks and slow path of sqrt():
// cos(2 pi u), u in (0, 1]: r = 2u - rint(2u) in [-1/2, 1/2], cos(2 pi u) = (-1)^rint(2u) cos(pi r), and cos(pi r) by
// its Taylor polynomial in r^2 up to r^20 (first dropped term < 2e-17; absolute error ~3e-16, like cospi's).
__device__ __forceinline__ double cos2pi_unit(double u) {
  const double t = u + u, big = 6755399441055744.0;   // 1.5 * 2^52: t + big has rint(t) in its low bits
  const double kb = t + big;
  const double r = t - (kb - big);
  const double s = r * r;
  double p = 0x1.ef6e308d6d1c4p-29;
  p = fma(p, s, -0x1.2a0c591af8314p-23);
  p = fma(p, s, 0x1.20c62c2f2d7f5p-18);
  p = fma(p, s, -0x1.b6e24f44b128fp-14);
  p = fma(p, s, 0x1.f9d38a3763cc3p-10);
  p = fma(p, s, -0x1.a6d1f2a204a8cp-6);
  p = fma(p, s, 0x1.e1f506891babbp-3);
  p = fma(p, s, -0x1.55d3c7e3cbffap+0);
  p = fma(p, s, 0x1.03c1f081b5ac4p+2);
  p = fma(p, s, -0x1.3bd3cc9be45dep+2);
  p = fma(p, s, 1.0);
  return __hiloint2double(__double2hiint(p) ^ (__double2loint(kb) << 31), __double2loint(p));
}
// sin(2 pi u) and cos(2 pi u) together (Box-Muller pairs of the deconvolution jitter): the same reduction, sin(pi r) by
// its Taylor polynomial up to r^21 (first dropped term < 2e-18).
__device__ __forceinline__ void sincos2pi_unit(double u, double &sn, double &cs) {
  const double t = u + u, big = 6755399441055744.0;
  const double kb = t + big;
  const double r = t - (kb - big);
  const double s = r * r;
  double p = 0x1.ef6e308d6d1c4p-29, q = 0x1.2877020d52cf0p-31;
  p = fma(p, s, -0x1.2a0c591af8314p-23); q = fma(q, s, -0x1.8a404211f9547p-26);
  p = fma(p, s, 0x1.20c62c2f2d7f5p-18);  q = fma(q, s, 0x1.aaec32af93359p-21);
  p = fma(p, s, -0x1.b6e24f44b128fp-14); q = fma(q, s, -0x1.6fadb9f155744p-16);
  p = fma(p, s, 0x1.f9d38a3763cc3p-10);  q = fma(q, s, 0x1.e8f434d018d63p-12);
  p = fma(p, s, -0x1.a6d1f2a204a8cp-6);  q = fma(q, s, -0x1.e3074fde8871fp-8);
  p = fma(p, s, 0x1.e1f506891babbp-3);   q = fma(q, s, 0x1.50783487ee782p-4);
  p = fma(p, s, -0x1.55d3c7e3cbffap+0);  q = fma(q, s, -0x1.32d2cce62bd86p-1);
  p = fma(p, s, 0x1.03c1f081b5ac4p+2);   q = fma(q, s, 0x1.466bc6775aae2p+1);
  p = fma(p, s, -0x1.3bd3cc9be45dep+2);  q = fma(q, s, -0x1.4abbce625be53p+2);
  p = fma(p, s, 1.0);                    q = fma(q, s, 0x1.921fb54442d18p+1);
  q *= r;
  const int flip = __double2loint(kb) << 31;
  cs = __hiloint2double(__double2hiint(p) ^ flip, __double2loint(p));
  sn = __hiloint2double(__double2hiint(q) ^ flip, __double2loint(q));
}
// sqrt(|a|) for |a| < 2^1000, |a| below 2^-1022 read as 2^-1022: reciprocal square root seed (2^-22), one coupled
// Goldschmidt step and one correction, within an ulp.
__device__ __forceinline__ double sqrt_small(double a) {
  a = __hiloint2double(max(__double2hiint(a) & 0x7fffffff, 0x00100000), __double2loint(a));
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double g0 = a * r, h0 = 0.5 * r;
  const double e = fma(-h0, g0, 0.5);
  const double g = fma(g0, e, g0), h = fma(h0, e, h0);
  return fma(fma(-g, g, a), h, g);
}
// 1 / x for normal x of moderate size (the t-weight's scale, x >= 2): reciprocal seed and two Newton steps, within an
// ulp of the division without its range checks.
__device__ __forceinline__ double rcp_moderate(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double normal_from_block(uint4 o, const double2 *__restrict__ tab) {
  const double u0 = u53(o.x, o.y), u1 = u53(o.z, o.w);
  return sqrt_small(-2.0 * log_tab(u0, tab)) * cos2pi_unit(u1);
}

// Box-Muller pair with the written-out factors (table logarithm, sqrt_small, sincos2pi_unit): same draws as rng_normal2
// to rounding, about half its instructions.
__device__ __forceinline__ void rng_normal2_tab(Key key, uint32_t purpose, uint32_t index, uint32_t iter, uint32_t slot,
                                                const double2 *__restrict__ tab, double &n0, double &n1) {
  const uint4 o = philox4x32_10(index, slot, iter, purpose, key);
  const double u0 = u53(o.x, o.y), u1 = u53(o.z, o.w);
  const double r = sqrt_small(-2.0 * log_tab(u0, tab));
  double s, c;
  sincos2pi_unit(u1, s, c);
  n0 = r * c;
  n1 = r * s;
}

// One Marsaglia-Tsang attempt for Gamma(shape >= 1): true when (x, u) is accepted, v3 = (1 + c x)^3.
// The second test, log u < x^2/2 + dd (1 - v3 + log v3), is screened in single precision like mh_accept().
__device__ __forceinline__ bool mt_attempt(double x, double u, double dd, double c, double &v3) {
  const double v = 1.0 + c * x;
  v3 = v * v * v;
  const double x2 = x * x;
  const bool squeeze = u < 1.0 - 0.0331 * x2 * x2;
  const float lu = __logf((float) u), lv = __logf((float) v3);
  const float gap = (float) ((0.5 * x2 + dd * ((1.0 - v3) + (double) lv)) - (double) lu);
  const bool decisive = fabsf(gap) > 1e-4f * (1.0f + fabsf(lu) + (float) dd * (1.0f + fabsf(lv)));
  bool ok = (v > 0.0) & (squeeze | (decisive & (gap > 0.0f)));
  if (v > 0.0 && !squeeze && !decisive) ok = log(u) < 0.5 * x2 + dd * (1.0 - v3 + log(v3));
  return ok;
}

// Gamma(shape >= 1, scale) for the t-model weights (Marsaglia & Tsang), draws laid out as described above:
// x0 / u_first belong to attempt 0; later attempts (about 1 in 200 spots at shape 9) draw their own.
__device__ __forceinline__ double rng_gamma_spot(Key key, uint32_t index, uint32_t iter, double shape, double scale,
                                                 double x0, double u_first, const double2 *__restrict__ tab) {
  const double dd = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
  double x = x0, u = u_first, v3;
  uint32_t t = 0;
  while (!mt_attempt(x, u, dd, c, v3)) {
    if (++t >= 0x3fffffffu) return __longlong_as_double(0x7ff8000000000000ll);
    x = normal_from_block(philox4x32_10(index, t, iter, P_SPOT_GAMMA, key), tab);
    u = u32c(philox4x32_10(index, 0x40000000u + t, iter, P_SPOT_GAMMA, key).x);
  }
  return dd * v3 * scale;
}

// The same draw without the table (k_sweep, sweep_kernels.cu): libm-style log() in the Box-Muller transform.
__device__ __forceinline__ double rng_gamma_spot(Key key, uint32_t index, uint32_t iter, double shape, double scale,
                                                 double u_first) {
  if (!(shape >= 1.0) || !(scale > 0.0)) return __longlong_as_double(0x7ff8000000000000ll);
  const double dd = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
  double u = u_first, v3;
  for (uint32_t t = 0; t < 0x3fffffffu; t++) {
    const double x = rng_normal_first(key, P_SPOT_GAMMA, index, iter, t);
    if (t) u = u32c(philox4x32_10(index, 0x40000000u + t, iter, P_SPOT_GAMMA, key).x);
    if (mt_attempt(x, u, dd, c, v3)) return dd * v3 * scale;
  }
  return __longlong_as_double(0x7ff8000000000000ll);
}

}   // namespace bsb
