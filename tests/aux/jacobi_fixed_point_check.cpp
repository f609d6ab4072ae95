// Test helper (tests/test_oracle.py): pf_jacobi_eig_sym stops at the first sweep that changes nothing.  This
// harness keeps the plain rule (always up to 60 sweeps, stop only when the off-diagonal part is exactly zero)
// and compares both over random symmetric matrices, full rank and rank deficient, at several scales: the
// eigenvalues and eigenvectors must have the same bits.  Prints "<compared> <different>".
#include <cmath>
#include <cstdio>
#include <cstring>
#include "bssm_b200/pf_smallmat.h"

static void jacobi_60(const double* X, int m, double* vals, double* vecs) {
  double A[256];
  for (int i = 0; i < m * m; ++i) A[i] = X[i];
  for (int j = 0; j < m; ++j)
    for (int i = j + 1; i < m; ++i) { double s = 0.5 * (A[i + m * j] + A[j + m * i]); A[i + m * j] = s; A[j + m * i] = s; }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) vecs[i + m * j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < j; ++i) off += A[i + m * j] * A[i + m * j];
    if (off == 0.0) break;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[p + m * q];
        if (apq == 0.0) continue;
        const double app = A[p + m * p], aqq = A[q + m * q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < m; ++k) { const double a = A[k + m * p], b = A[k + m * q]; A[k + m * p] = c * a - s * b; A[k + m * q] = s * a + c * b; }
        for (int k = 0; k < m; ++k) { const double a = A[p + m * k], b = A[q + m * k]; A[p + m * k] = c * a - s * b; A[q + m * k] = s * a + c * b; }
        for (int k = 0; k < m; ++k) { const double a = vecs[k + m * p], b = vecs[k + m * q]; vecs[k + m * p] = c * a - s * b; vecs[k + m * q] = s * a + c * b; }
      }
  }
  for (int i = 0; i < m; ++i) vals[i] = A[i + m * i];
  for (int i = 0; i < m - 1; ++i) {
    int best = i;
    for (int j = i + 1; j < m; ++j) if (fabs(vals[j]) > fabs(vals[best])) best = j;
    if (best != i) {
      double tv = vals[i]; vals[i] = vals[best]; vals[best] = tv;
      for (int k = 0; k < m; ++k) { double t = vecs[k + m * i]; vecs[k + m * i] = vecs[k + m * best]; vecs[k + m * best] = t; }
    }
  }
  for (int j = 0; j < m; ++j) {
    int big = 0;
    for (int k = 1; k < m; ++k) if (fabs(vecs[k + m * j]) > fabs(vecs[big + m * j])) big = k;
    if (vecs[big + m * j] < 0.0) for (int k = 0; k < m; ++k) vecs[k + m * j] = -vecs[k + m * j];
  }
}

int main() {
  unsigned long long st = 99;
  auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) / 9007199254740992.0; };
  long bad = 0, tot = 0;
  for (int m = 1; m <= 13; ++m)
    for (int rep = 0; rep < 1500; ++rep) {
      double B[169], X[169], v1[13], U1[169], v2[13], U2[169];
      for (int i = 0; i < m * m; ++i) B[i] = rnd() - 0.5;
      const double sc = rep % 4 == 0 ? 1e-6 : (rep % 4 == 1 ? 1e3 : 1.0);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          double s = 0;
          for (int l = 0; l < (rep % 5 == 0 ? m - 1 : m); ++l) s += B[i + m * l] * B[j + m * l];
          X[i + m * j] = s * sc;
        }
      jacobi_60(X, m, v1, U1);
      pf_jacobi_eig_sym(X, m, v2, U2);
      tot++;
      if (memcmp(v1, v2, sizeof(double) * m) || memcmp(U1, U2, sizeof(double) * m * m)) bad++;
    }
  printf("%ld %ld\n", tot, bad);
  return 0;
}
