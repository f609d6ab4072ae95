/* pf_smallmat.h -- deterministic small dense-matrix primitives (host and device),
 * column-major like R / Armadillo.  Used by the host-side Gaussian-approximation
 * pre-pass of the psi-APF for m > 1 (product: bssm_b200/csrc/api.cu) and by the
 * CPU oracle, so that both evaluate products in the SAME summation order
 * (accumulate over the inner index, left to right, no FMA contraction when
 * compiled with -ffp-contract=off).  The Kalman / smoother / conditional-
 * covariance ALGORITHMS are restated separately on both sides; only these
 * primitives are shared, like the scalar functions of pf_math.h.
 *
 * Armadillo evaluates the same expressions through BLAS / LAPACK of the R
 * installation [UNVENDORED]: association order of chained products and the
 * signs of singular vectors are not reproducible outside that build, which is
 * part of why parity is declared "unpinned" at the Armadillo boundary.
 */
#ifndef BSSM_B200_PF_SMALLMAT_H
#define BSSM_B200_PF_SMALLMAT_H

#include <math.h>
#include <stddef.h>

/* the same functions run inside the per-draw approximation kernel (bssm_b200/csrc/approx_device.cuh) */
#if defined(__CUDACC__)
#define PF_SM static __host__ __device__ inline
#else
#define PF_SM static inline
#endif

/* C (r x c) = A (r x k) * B (k x c) */
PF_SM void pf_mm(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i + (size_t)r * l] * B[l + (size_t)k * j];
      C[i + (size_t)r * j] = s;
    }
}
/* C (r x c) = A (r x k) * B' where B is (c x k) */
PF_SM void pf_mmt(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i + (size_t)r * l] * B[j + (size_t)c * l];
      C[i + (size_t)r * j] = s;
    }
}
/* C (r x c) = A' * B where A is (k x r), B is (k x c) */
PF_SM void pf_mtm(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[l + (size_t)k * i] * B[l + (size_t)k * j];
      C[i + (size_t)r * j] = s;
    }
}
/* arma::symmatu: copy the upper triangle onto the lower */
PF_SM void pf_symmatu(double* A, int m) {
  for (int j = 0; j < m; ++j)
    for (int i = j + 1; i < m; ++i) A[i + (size_t)m * j] = A[j + (size_t)m * i];
}

/* Eigen-decomposition of a symmetric m x m matrix by cyclic Jacobi rotations
 * (fixed sweep order, so the result is reproducible).  On return `vals` holds
 * the eigenvalues sorted by DECREASING ABSOLUTE VALUE (the order of singular
 * values) and `vecs` (m x m, column-major) the matching unit eigenvectors, each
 * normalised so that its largest-magnitude component is positive.  This stands
 * in for arma::svd_econ(U, s, V, X, "left") on symmetric X: s = |vals|, U = vecs
 * (src/conditional_dist.cpp:14, :21). */
/* MAXMM: size of the work matrix (m * m <= MAXMM); the arithmetic does not depend on it */
template <int MAXMM>
PF_SM void pf_jacobi_eig_sym_t(const double* X, int m, double* vals, double* vecs) {
  double A[MAXMM];
  for (int i = 0; i < m * m; ++i) A[i] = X[i];
  /* work on the symmetrised matrix */
  for (int j = 0; j < m; ++j)
    for (int i = j + 1; i < m; ++i) {
      double s = 0.5 * (A[i + m * j] + A[j + m * i]);
      A[i + m * j] = s;
      A[j + m * i] = s;
    }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) vecs[i + m * j] = (i == j) ? 1.0 : 0.0;
  /* Up to 60 sweeps, until the off-diagonal part is exactly zero.  Rounding usually leaves an off-diagonal
   * element of ~1e-17 relative size that never becomes zero; the sweeps then reach a FIXED POINT (a sweep that
   * changes no element of A or vecs, typically the 10th-14th) and every later sweep would repeat it: stopping
   * there gives the same bits as running all 60 (checked over random matrices, tests/test_oracle.py). */
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < j; ++i) off += A[i + m * j] * A[i + m * j];
    if (off == 0.0) break;
    int moved = 0;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[p + m * q];
        if (apq == 0.0) continue;
        const double app = A[p + m * p], aqq = A[q + m * q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < m; ++k) { /* A <- A * J */
          const double akp = A[k + m * p], akq = A[k + m * q];
          const double n1 = c * akp - s * akq, n2 = s * akp + c * akq;
          moved |= (n1 != akp) | (n2 != akq);
          A[k + m * p] = n1;
          A[k + m * q] = n2;
        }
        for (int k = 0; k < m; ++k) { /* A <- J' * A */
          const double apk = A[p + m * k], aqk = A[q + m * k];
          const double n1 = c * apk - s * aqk, n2 = s * apk + c * aqk;
          moved |= (n1 != apk) | (n2 != aqk);
          A[p + m * k] = n1;
          A[q + m * k] = n2;
        }
        for (int k = 0; k < m; ++k) {
          const double vkp = vecs[k + m * p], vkq = vecs[k + m * q];
          const double n1 = c * vkp - s * vkq, n2 = s * vkp + c * vkq;
          moved |= (n1 != vkp) | (n2 != vkq);
          vecs[k + m * p] = n1;
          vecs[k + m * q] = n2;
        }
      }
    if (!moved) break; /* fixed point: the remaining sweeps would reproduce this one */
  }
  for (int i = 0; i < m; ++i) vals[i] = A[i + m * i];
  /* selection sort by decreasing |value| (stable for ties), sign normalisation */
  for (int i = 0; i < m - 1; ++i) {
    int best = i;
    for (int j = i + 1; j < m; ++j)
      if (fabs(vals[j]) > fabs(vals[best])) best = j;
    if (best != i) {
      double tv = vals[i]; vals[i] = vals[best]; vals[best] = tv;
      for (int k = 0; k < m; ++k) {
        double t = vecs[k + m * i]; vecs[k + m * i] = vecs[k + m * best]; vecs[k + m * best] = t;
      }
    }
  }
  for (int j = 0; j < m; ++j) {
    int big = 0;
    for (int k = 1; k < m; ++k)
      if (fabs(vecs[k + m * j]) > fabs(vecs[big + m * j])) big = k;
    if (vecs[big + m * j] < 0.0)
      for (int k = 0; k < m; ++k) vecs[k + m * j] = -vecs[k + m * j];
  }
}

PF_SM void pf_jacobi_eig_sym(const double* X, int m, double* vals, double* vecs) {
  pf_jacobi_eig_sym_t<16 * 16>(X, m, vals, vecs);
}

#endif /* BSSM_B200_PF_SMALLMAT_H */
