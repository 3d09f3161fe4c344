// smallmat.cuh -- warp-cooperative dense linear algebra on d x d matrices (d <= 32) held in shared
// memory, column-major with leading dimension d.  One warp per matrix; every routine ends with a
// __syncwarp() so results are visible to all lanes of the calling warp.
//
// These stand in for the Armadillo calls of the reference's conjugate updates
// (chol / inv / products at src/cluster.cpp:254-270, 367-389, 488-505, 619-643, 888-902) and are
// run once per iteration, not once per spot.
#pragma once
#include <cuda_runtime.h>

namespace bsb {

#define SM_AT(M, i, j) (M)[(size_t) (j) * d + (i)]

// R (upper, R^T R = A) from the upper triangle of A.  A and R must not alias.  Lower part zeroed.
// Returns false (on every lane) when A is not positive definite.
__device__ inline bool warp_chol_upper(const double *A, double *R, int d, int lane) {
  bool ok = true;
  for (int j = 0; j < d; j++) {
    double s = SM_AT(A, j, j);
    for (int k = 0; k < j; k++) s -= SM_AT(R, k, j) * SM_AT(R, k, j);
    if (!(s > 0.0)) ok = false;
    const double rjj = sqrt(s);
    for (int i = j + 1 + lane; i < d; i += 32) {
      double t = SM_AT(A, j, i);
      for (int k = 0; k < j; k++) t -= SM_AT(R, k, j) * SM_AT(R, k, i);
      SM_AT(R, j, i) = t / rjj;
    }
    for (int i = lane; i < j; i += 32) SM_AT(R, j, i) = 0.0;
    __syncwarp();
    if (lane == 0) SM_AT(R, j, j) = rjj;
    __syncwarp();
  }
  return ok;
}

// U (upper, U U^T = A) from the upper triangle of A: the Cholesky factorisation taken from the last
// row / column backwards.  If A = X^-1 with X = R^T R (R upper), then U = R^-1: this is how the kernels
// obtain chol(X^-1) and inv(chol(Sigma)) without forming an inverse.  A and U must not alias.
__device__ inline bool warp_revchol_upper(const double *A, double *U, int d, int lane) {
  // right-looking, in place on a copy: after column j is scaled, the leading (j x j) block is
  // down-dated by the outer product of that column -- independent FMAs, no long dependent chain.
  for (int e = lane; e < d * d; e += 32) U[e] = (e % d) <= (e / d) ? A[e] : 0.0;
  __syncwarp();
  bool ok = true;
  for (int j = d - 1; j >= 0; j--) {
    const double s = SM_AT(U, j, j);
    if (!(s > 0.0)) ok = false;
    const double ujj = sqrt(s), inv = 1.0 / ujj;
    __syncwarp();
    for (int i = lane; i < j; i += 32) SM_AT(U, i, j) *= inv;
    if (lane == 0) SM_AT(U, j, j) = ujj;
    __syncwarp();
    for (int k = lane; k < j; k += 32) {
      const double ukj = SM_AT(U, k, j);
#pragma unroll 4
      for (int i = 0; i <= k; i++) SM_AT(U, i, k) -= SM_AT(U, i, j) * ukj;
    }
    __syncwarp();
  }
  return ok;
}

// y = R x for upper-triangular R (y_i = sum_{j>=i} R_ij x_j); x, y in shared memory, not aliased.
__device__ inline void warp_triu_matvec(const double *R, const double *x, double *y, int d, int lane) {
  for (int i = lane; i < d; i += 32) {
    double s = 0.0;
    for (int j = i; j < d; j++) s += SM_AT(R, i, j) * x[j];
    y[i] = s;
  }
  __syncwarp();
}

// y = R^T x for upper-triangular R (y_j = sum_{i<=j} R_ij x_i).
__device__ inline void warp_triu_matvec_t(const double *R, const double *x, double *y, int d, int lane) {
  for (int j = lane; j < d; j += 32) {
    double s = 0.0;
    for (int i = 0; i <= j; i++) s += SM_AT(R, i, j) * x[i];
    y[j] = s;
  }
  __syncwarp();
}

// T = R^-1 for upper-triangular R (T upper).  Must not alias.
__device__ inline void warp_inv_upper(const double *R, double *T, int d, int lane) {
  for (int j = lane; j < d; j += 32) {
    SM_AT(T, j, j) = 1.0 / SM_AT(R, j, j);
    for (int i = j - 1; i >= 0; i--) {
      double s = 0.0;
      for (int k = i + 1; k <= j; k++) s += SM_AT(R, i, k) * SM_AT(T, k, j);
      SM_AT(T, i, j) = -s / SM_AT(R, i, i);
    }
    for (int i = j + 1; i < d; i++) SM_AT(T, i, j) = 0.0;
  }
  __syncwarp();
}

// C = T T^T for upper-triangular T (the inverse of X when T = chol_upper(X)^-1).
__device__ inline void warp_ttt(const double *T, double *C, int d, int lane) {
  for (int j = lane; j < d; j += 32)
    for (int i = 0; i < d; i++) {
      const int k0 = i > j ? i : j;
      double s = 0.0;
      for (int k = k0; k < d; k++) s += SM_AT(T, i, k) * SM_AT(T, j, k);
      SM_AT(C, i, j) = s;
    }
  __syncwarp();
}

// C = T^T T for upper-triangular T.
__device__ inline void warp_ttt_tn(const double *T, double *C, int d, int lane) {
  for (int j = lane; j < d; j += 32)
    for (int i = 0; i < d; i++) {
      const int k1 = i < j ? i : j;
      double s = 0.0;
      for (int k = 0; k <= k1; k++) s += SM_AT(T, k, i) * SM_AT(T, k, j);
      SM_AT(C, i, j) = s;
    }
  __syncwarp();
}

// C = A^T B (general).
__device__ inline void warp_matmul_tn(const double *A, const double *B, double *C, int d, int lane) {
  for (int j = lane; j < d; j += 32)
    for (int i = 0; i < d; i++) {
      double s = 0.0;
      for (int k = 0; k < d; k++) s += SM_AT(A, k, i) * SM_AT(B, k, j);
      SM_AT(C, i, j) = s;
    }
  __syncwarp();
}

// y = A x (general A), x and y length d in shared memory, not aliased.
__device__ inline void warp_matvec(const double *A, const double *x, double *y, int d, int lane) {
  for (int i = lane; i < d; i += 32) {
    double s = 0.0;
    for (int j = 0; j < d; j++) s += SM_AT(A, i, j) * x[j];
    y[i] = s;
  }
  __syncwarp();
}

// Inverse of a symmetric positive definite X through its Cholesky factor.  scratch: 2*d*d doubles.
// Returns false when X is not positive definite.  Xi must not alias X.
__device__ inline bool warp_inv_spd(const double *X, double *Xi, double *scratch, int d, int lane) {
  double *R = scratch, *T = scratch + d * d;
  const bool ok = warp_chol_upper(X, R, d, lane);
  warp_inv_upper(R, T, d, lane);
  warp_ttt(T, Xi, d, lane);
  return ok;
}

}   // namespace bsb
