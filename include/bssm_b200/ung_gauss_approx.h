/* ung_gauss_approx.h -- host-side Gaussian approximation of an ssm_ung model for
 * any state dimension m (C++17, header only): everything ssm_ung::psi_filter
 * does BEFORE its particle loop (/root/reference/src/model_ssm_ung.cpp:514-533).
 *
 *   approximate()            src/model_ssm_ung.cpp:87-140   Newton / Laplace sweeps
 *   laplace_iter             src/model_ssm_ung.cpp:302-342
 *   update_scales            src/model_ssm_ung.cpp:241-292
 *   compute_const_term       src/model_ssm_ung.cpp:347-370
 *   ssm_ulg::log_likelihood  src/model_ssm_ulg.cpp:106-138
 *   ssm_ulg::fast_smoother   src/model_ssm_ulg.cpp:252-305
 *   ssm_ulg::smoother_ccov   src/model_ssm_ulg.cpp:419-473
 *   conditional_cov (SVD)    src/conditional_dist.cpp:8-32
 *
 * O(n m^3), strictly sequential in t: host work, done once per filter call.
 * Chained matrix products are associated left to right with the primitives of
 * pf_smallmat.h; the SVD of the (symmetric) smoothed covariances is a
 * deterministic Jacobi eigen-decomposition.  For m = 1 every formula reduces
 * bit for bit to the scalar restatements (oracle/pf_oracle.cpp Ulg1,
 * bssm_b200/csrc/prepass_ung.cuh), which tests/test_oracle.py checks.
 *
 * Used by the product for m > 1 (bssm_b200/csrc/api.cu) and by the oracle's
 * general-m psi-APF; the m = 1 paths of both stay independent of this file.
 */
#ifndef BSSM_B200_UNG_GAUSS_APPROX_H
#define BSSM_B200_UNG_GAUSS_APPROX_H

#include <cmath>
#include <limits>
#include <vector>

#include <stdexcept>
#include "bssm_b200.h"
#include "bssm_b200/pf_math.h"
#include "bssm_b200/pf_smallmat.h"

namespace pfga {

/* log g(y | signal) up to constants, src/model_ssm_ung.cpp:447-493 */
PF_HD double ung_log_obs(int dist, double s, double y, double u, double phi) {
  switch (dist) {
    case 0: return -0.5 * (s + (y / phi) * (y / phi) * pf_exp(-s));
    case 1: return y * s - u * pf_exp(s);
    case 2: return y * s - u * pf_log1p(pf_exp(s));
    case 3: return y * s - (y + phi) * pf_log(phi + u * pf_exp(s));
    default: return -phi * (s + (y * pf_exp(-s) / u));
  }
}

/* arma::accu of a vector: two accumulators over the even / odd elements [UNVENDORED], the
 * convention every sum of this header that stands for accu() follows */
inline double accu2(const double* x, size_t n) {
  double a1 = 0.0, a2 = 0.0;
  size_t j;
  for (j = 1; j < n; j += 2) { a1 += x[j - 1]; a2 += x[j]; }
  if ((j - 1) < n) a1 += x[j - 1];
  return a1 + a2;
}

/* conditional_cov, SVD branch (src/conditional_dist.cpp:8-32), in place:
 * Pt (m x m x (n+1) smoothed covariances) becomes the factors L_t = U sqrt(s),
 * ccov (lag-one covariances) becomes C_t = Cov_{t-1}' V_{t-1}^+ */
inline void conditional_cov_svd(int m, int n, std::vector<double>& Pt, std::vector<double>& ccov) {
  const size_t mm = (size_t)m * m;
  const double eps = std::numeric_limits<double>::epsilon();
  std::vector<double> vals(m), U(mm), tmp(mm), W(mm), A(mm);
  auto factor = [&](double* V) { /* V <- U diag(sqrt(s)) */
    pf_jacobi_eig_sym(V, m, vals.data(), U.data());
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) V[i + (size_t)m * j] = U[i + (size_t)m * j] * std::sqrt(std::fabs(vals[j]));
  };
  for (int t = n; t > 0; --t) {
    pf_jacobi_eig_sym(&Pt[mm * (t - 1)], m, vals.data(), U.data());
    const double s0 = std::fabs(vals[0]);
    /* tmp = Ct(t-1)' * U_nz * diag(1/s_nz) * U_nz' */
    pf_mtm(&ccov[mm * (t - 1)], U.data(), W.data(), m, m, m);
    for (int j = 0; j < m; ++j) {
      const double s = std::fabs(vals[j]);
      const double inv = (s > eps * m * s0) ? 1.0 / s : 0.0;
      for (int i = 0; i < m; ++i) W[i + (size_t)m * j] = W[i + (size_t)m * j] * inv;
    }
    pf_mmt(W.data(), U.data(), tmp.data(), m, m, m);
    pf_mm(tmp.data(), &ccov[mm * (t - 1)], A.data(), m, m, m);
    for (size_t e = 0; e < mm; ++e) Pt[mm * t + e] -= A[e];
    for (size_t e = 0; e < mm; ++e) ccov[mm * t + e] = tmp[e];
    factor(&Pt[mm * t]);
  }
  factor(&Pt[0]);
}

/* the stack work arrays are sized for m = 16 and used up to the run-time m * m: gcc cannot see that */
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
#endif
struct UngApprox {
  const bssm_ung_model* M;
  int n, m;
  std::vector<double> xbeta;
  std::vector<double> mode, ay, aH, aHH, scales; /* n each */
  double approx_loglik = 0.0;
  std::vector<double> alphahat, L, Ct; /* m x (n+1), m x m x (n+1), m x m x (n+1) */
  static constexpr double zero_tol = 1e-12; /* src/model_ssm_ulg.h:19 */

  explicit UngApprox(const bssm_ung_model* M_) : M(M_), n(M_->n), m(M_->m) {
    if (m > MAXM) throw std::invalid_argument("UngApprox: state dimension m <= 16");
    xbeta.assign(n, 0.0);
    if (M->xbeta) for (int t = 0; t < n; ++t) xbeta[t] = M->xbeta[t];
    mode.assign(n, 0.0); ay.assign(n, 0.0); aH.assign(n, 0.0); aHH.assign(n, 0.0); scales.assign(n, 0.0);
  }
  const double* Z(int t) const { return M->Z + (size_t)m * (t * M->Ztv); }
  const double* T(int t) const { return M->T + (size_t)m * m * (t * M->Ttv); }
  const double* R(int t) const { return M->R + (size_t)m * M->k * (t * M->Rtv); }
  const double* C(int t) const { return M->C + (size_t)m * (t * M->Ctv); }
  double D(int t) const { return M->D[t * M->Dtv]; }
  double ytmp(int t) const { return M->xbeta ? ay[t] - xbeta[t] : ay[t]; }
  void RR(int t, double* out) const { pf_mmt(R(t), R(t), out, m, M->k, m); } /* R R', compute_RR */

  void laplace_iter(const std::vector<double>& signal) {
    const double phi = M->phi;
    for (int t = 0; t < n; ++t) {
      const double y = M->y[t], u = M->u[t], s = signal[t];
      switch (M->distribution) {
        case 0: {
          double tmp = y;
          if (std::fabs(tmp) < 1e-4) tmp = 1e-4;
          double q = tmp / phi;
          aHH[t] = 2.0 * pf_exp(s) / (q * q);
          ay[t] = s + 1.0 - 0.5 * aHH[t];
        } break;
        case 1: {
          aHH[t] = 1.0 / (pf_exp(s) * u);
          ay[t] = y * aHH[t] + s - 1.0;
        } break;
        case 2: {
          double e = pf_exp(s);
          aHH[t] = (1.0 + e) * (1.0 + e) / (u * e);
          ay[t] = y * aHH[t] + s - 1.0 - e;
        } break;
        case 3: {
          double e = pf_exp(s) * u;
          aHH[t] = (phi + e) * (phi + e) / (phi * e * (y + phi));
          ay[t] = s + (phi + e) * (y - e) / ((y + phi) * e);
        } break;
        default: {
          double e = pf_exp(s) * u;
          aHH[t] = e / (y * phi);
          ay[t] = s - e / y + 1;
        } break;
      }
      aH[t] = std::sqrt(aHH[t]);
    }
  }

  void update_scales() {
    for (int t = 0; t < n; ++t) {
      scales[t] = 0.0;
      if (std::isfinite(M->y[t])) {
        double q = (ay[t] - mode[t]) / aH[t];
        scales[t] = ung_log_obs(M->distribution, mode[t], M->y[t], M->u[t], M->phi) + 0.5 * (q * q);
      }
    }
  }

  /* one Kalman step shared by the three recursions: returns whether y_t was used */
  /* work arrays live on the stack (no allocation per time point): the engine and the oracle stop at m = 16 */
  static constexpr int MAXM = 16;
  struct Step {
    double F, v;
    double K[MAXM];
  };
  bool kalman_gain(int t, const double* at, const double* Pt, Step& st) const {
    double zp[MAXM];
    pf_mtm(Z(t), Pt, zp, 1, m, m);          /* Z' Pt */
    double f = 0.0;
    for (int i = 0; i < m; ++i) f += zp[i] * Z(t)[i];
    st.F = f + aHH[t];
    const double yy = ytmp(t);
    if (!(std::isfinite(yy) && st.F > zero_tol)) return false;
    pf_mm(Pt, Z(t), st.K, m, m, 1);         /* Pt Z / F */
    for (int i = 0; i < m; ++i) st.K[i] = st.K[i] / st.F;
    double za = 0.0;
    for (int i = 0; i < m; ++i) za += Z(t)[i] * at[i];
    st.v = yy - D(t) - za;
    return true;
  }
  void predict_mean(int t, const double* at, const Step* st, double* anext) const {
    double tmp[MAXM], ta[MAXM];
    for (int i = 0; i < m; ++i) tmp[i] = st ? at[i] + st->K[i] * st->v : at[i];
    pf_mm(T(t), tmp, ta, m, m, 1);
    for (int i = 0; i < m; ++i) anext[i] = C(t)[i] + ta[i];
  }
  /* Joseph form used by fast_smoother / smoother_ccov (:276-279, :441-443) */
  void predict_cov_joseph(int t, const double* Pt, const Step* st, double* Pnext) const {
    double X[MAXM * MAXM], A[MAXM * MAXM], B[MAXM * MAXM], rr[MAXM * MAXM];
    if (st) {
      double tmp[MAXM * MAXM];
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) tmp[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - st->K[i] * Z(t)[j];
      pf_mm(tmp, Pt, A, m, m, m);
      pf_mmt(A, tmp, B, m, m, m);  /* tmp Pt tmp' */
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = B[i + (size_t)m * j] + st->K[i] * aHH[t] * st->K[j];
    } else {
      for (int e = 0; e < m * m; ++e) X[e] = Pt[e];
    }
    pf_mm(T(t), X, A, m, m, m);
    pf_mmt(A, T(t), B, m, m, m);
    RR(t, rr);
    for (int e = 0; e < m * m; ++e) Pnext[e] = B[e] + rr[e];
    pf_symmatu(Pnext, m);
  }

  /* ssm_ulg::log_likelihood, src/model_ssm_ulg.cpp:106-138 */
  double log_likelihood() const {
    double sumH = 0.0, sumR = 0.0;
    for (int t = 0; t < n; ++t) sumH += aH[t];
    const int nR = (M->Rtv ? n : 1) * m * M->k;
    for (int e = 0; e < nR; ++e) sumR += M->R[e];
    if (sumH + sumR < zero_tol) return -std::numeric_limits<double>::infinity();
    double logLik = 0.0;
    std::vector<double> at(M->a1, M->a1 + m), Pt(M->P1, M->P1 + (size_t)m * m), an(m), Pn((size_t)m * m);
    const double LOG2PI = std::log(2.0 * M_PI);
    for (int t = 0; t < n; ++t) {
      Step st;
      if (kalman_gain(t, at.data(), Pt.data(), st)) {
        predict_mean(t, at.data(), &st, an.data());
        double X[MAXM * MAXM], A[MAXM * MAXM], B[MAXM * MAXM], rr[MAXM * MAXM];
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = Pt[i + (size_t)m * j] - st.K[i] * st.K[j] * st.F;
        pf_mm(T(t), X, A, m, m, m);
        pf_mmt(A, T(t), B, m, m, m);
        RR(t, rr);
        for (int e = 0; e < m * m; ++e) Pn[e] = B[e] + rr[e];
        pf_symmatu(Pn.data(), m);
        logLik -= 0.5 * (LOG2PI + pf_log(st.F) + st.v * st.v / st.F);
      } else {
        predict_mean(t, at.data(), nullptr, an.data());
        predict_cov_joseph(t, Pt.data(), nullptr, Pn.data());
      }
      at = an; Pt = Pn;
    }
    return logLik;
  }

  /* what ssm_ulg::fast_precomputing_smoother keeps for the simulation smoother (src/model_ssm_ulg.cpp:359-413):
   * F_t, K_t (m x n), L_t = T_t (I - K_t Z_t') (m x m x n; also for t = 0, which the reference rebuilds at :400),
   * the time points that were updated and R_t R_t' */
  struct SimGains {
    std::vector<double> F, K, Lt, RR;
    std::vector<char> used;
  };
  /* ssm_ulg::fast_smoother, src/model_ssm_ulg.cpp:252-305: at is m x (n+1); with `keep` it is
   * fast_precomputing_smoother (:359-413, the same arithmetic) */
  void fast_smoother(std::vector<double>& at, SimGains* keep = nullptr) const {
    at.assign((size_t)m * (n + 1), 0.0);
    std::vector<double> Pt(M->P1, M->P1 + (size_t)m * m), Pn((size_t)m * m);
    std::vector<Step> st(n);
    std::vector<char> used(n, 0);
    for (int i = 0; i < m; ++i) at[i] = M->a1[i];
    for (int t = 0; t < n; ++t) {
      used[t] = kalman_gain(t, &at[(size_t)m * t], Pt.data(), st[t]);
      predict_mean(t, &at[(size_t)m * t], used[t] ? &st[t] : nullptr, &at[(size_t)m * (t + 1)]);
      predict_cov_joseph(t, Pt.data(), used[t] ? &st[t] : nullptr, Pn.data());
      Pt = Pn;
    }
    std::vector<double> rt((size_t)m * n, 0.0);
    auto Lmat = [&](int t, double* Lm) { /* T (I - K Z') */
      double tmp[MAXM * MAXM];
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) tmp[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - st[t].K[i] * Z(t)[j];
      pf_mm(T(t), tmp, Lm, m, m, m);
    };
    std::vector<double> Lm((size_t)m * m), lr(m);
    for (int t = n - 1; t > 0; --t) {
      if (used[t]) {
        Lmat(t, Lm.data());
        pf_mtm(Lm.data(), &rt[(size_t)m * t], lr.data(), m, m, 1);
        for (int i = 0; i < m; ++i) rt[i + (size_t)m * (t - 1)] = Z(t)[i] / st[t].F * st[t].v + lr[i];
      } else {
        pf_mtm(T(t), &rt[(size_t)m * t], &rt[(size_t)m * (t - 1)], m, m, 1);
      }
    }
    std::vector<double> w(m), pw(m);
    if (std::isfinite(ay[0]) && st[0].F > zero_tol) {
      Lmat(0, Lm.data());
      pf_mtm(Lm.data(), &rt[0], lr.data(), m, m, 1);
      for (int i = 0; i < m; ++i) w[i] = Z(0)[i] / st[0].F * st[0].v + lr[i];
      pf_mm(M->P1, w.data(), pw.data(), m, m, 1);
    } else if (m == 1) {
      /* a1 + P1 * T' * rt: Armadillo associates (P1 T') rt for 1 x 1 operands, P1 (T' rt) otherwise */
      pw[0] = M->P1[0] * T(0)[0] * rt[0];
    } else {
      pf_mtm(T(0), &rt[0], w.data(), m, m, 1);
      pf_mm(M->P1, w.data(), pw.data(), m, m, 1);
    }
    for (int i = 0; i < m; ++i) at[i] = M->a1[i] + pw[i];
    std::vector<double> ta(m), rr((size_t)m * m), rrr(m);
    for (int t = 0; t < n - 1; ++t) {
      pf_mm(T(t), &at[(size_t)m * t], ta.data(), m, m, 1);
      RR(t, rr.data());
      pf_mm(rr.data(), &rt[(size_t)m * t], rrr.data(), m, m, 1);
      for (int i = 0; i < m; ++i) at[i + (size_t)m * (t + 1)] = C(t)[i] + ta[i] + rrr[i];
    }
    if (keep) {
      const size_t mm = (size_t)m * m;
      keep->F.assign(n, 0.0); keep->K.assign((size_t)m * n, 0.0); keep->Lt.assign(mm * n, 0.0);
      keep->used.assign(used.begin(), used.end());
      for (int t = 0; t < n; ++t) {
        keep->F[t] = st[t].F;
        if (!used[t]) continue;
        for (int i = 0; i < m; ++i) keep->K[i + (size_t)m * t] = st[t].K[i];
        Lmat(t, &keep->Lt[mm * t]);
      }
      const int nR = M->Rtv ? n : 1;
      keep->RR.assign(mm * nR, 0.0);
      for (int t = 0; t < nR; ++t) RR(t, &keep->RR[mm * t]);
    }
  }

  /* ssm_ung::approximate, src/model_ssm_ung.cpp:87-140 (local approximation branch);
   * mode_lin = the mode the pseudo-observations were last linearised at */
  void approximate() {
    for (int t = 0; t < n; ++t) mode[t] = M->initial_mode[t];
    int i = 0;
    double diff = M->conv_tol + 1;
    std::vector<double> at, mode_new(n);
    while (i < M->max_iter && diff > M->conv_tol) {
      i++;
      laplace_iter(mode);
      fast_smoother(at);
      for (int t = 0; t < n; ++t) {
        if (M->distribution == 0) {
          mode_new[t] = at[(size_t)m * t]; /* head_cols of a 1 x (n+1) matrix: svm has m = 1 */
        } else {
          double za = 0.0;
          for (int d = 0; d < m; ++d) za += Z(t)[d] * at[d + (size_t)m * t];
          mode_new[t] = xbeta[t] + D(t) + za;
        }
      }
      double a1 = 0.0, a2 = 0.0; /* arma::accu, two accumulators [UNVENDORED] */
      int j;
      for (j = 1; j < n; j += 2) {
        a1 += (mode_new[j - 1] - mode[j - 1]) * (mode_new[j - 1] - mode[j - 1]);
        a2 += (mode_new[j] - mode[j]) * (mode_new[j] - mode[j]);
      }
      if ((j - 1) < n) a1 += (mode_new[j - 1] - mode[j - 1]) * (mode_new[j - 1] - mode[j - 1]);
      diff = (a1 + a2) / n;
      mode = mode_new;
    }
  }

  /* ssm_ulg::smoother_ccov (:419-473) followed by conditional_cov (SVD branch) */
  void smoother_and_conditional() {
    const size_t mm = (size_t)m * m;
    alphahat.assign((size_t)m * (n + 1), 0.0);
    std::vector<double> Pt(mm * (n + 1), 0.0), ccov(mm * (n + 1), 0.0);
    std::vector<Step> st(n);
    std::vector<char> used(n, 0);
    for (int i = 0; i < m; ++i) alphahat[i] = M->a1[i];
    for (size_t e = 0; e < mm; ++e) Pt[e] = M->P1[e];
    for (int t = 0; t < n; ++t) {
      used[t] = kalman_gain(t, &alphahat[(size_t)m * t], &Pt[mm * t], st[t]);
      predict_mean(t, &alphahat[(size_t)m * t], used[t] ? &st[t] : nullptr, &alphahat[(size_t)m * (t + 1)]);
      predict_cov_joseph(t, &Pt[mm * t], used[t] ? &st[t] : nullptr, &Pt[mm * (t + 1)]);
      for (size_t e = 0; e < mm; ++e) ccov[mm * t + e] = Pt[mm * (t + 1) + e];
    }
    std::vector<double> rt(m, 0.0), Nt(mm, 0.0), Lm(mm), A(mm), B(mm), E(mm), lr(m), tmpv(m);
    for (int t = n - 1; t >= 0; --t) {
      const double* Tm;
      double Ltmp[MAXM * MAXM];
      if (used[t]) {
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) Ltmp[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - st[t].K[i] * Z(t)[j];
        pf_mm(T(t), Ltmp, Lm.data(), m, m, m);
        Tm = Lm.data();
      } else {
        Tm = T(t);
      }
      /* ccov_t = Pt_t * L' * (I - Nt * ccov_t)   (L = T when y_t is missing) */
      pf_mm(Nt.data(), &ccov[mm * t], A.data(), m, m, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - A[i + (size_t)m * j];
      pf_mmt(&Pt[mm * t], Tm, B.data(), m, m, m);
      pf_mm(B.data(), E.data(), &ccov[mm * t], m, m, m);
      /* rt, Nt */
      pf_mtm(Tm, rt.data(), lr.data(), m, m, 1);
      pf_mtm(Tm, Nt.data(), A.data(), m, m, m);
      pf_mm(A.data(), Tm, B.data(), m, m, m);
      if (used[t]) {
        for (int i = 0; i < m; ++i) rt[i] = Z(t)[i] / st[t].F * st[t].v + lr[i];
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) Nt[i + (size_t)m * j] = Z(t)[i] * Z(t)[j] / st[t].F + B[i + (size_t)m * j];
      } else {
        rt = lr;
        Nt = B;
      }
      pf_symmatu(Nt.data(), m);
      pf_mm(&Pt[mm * t], rt.data(), tmpv.data(), m, m, 1);
      for (int i = 0; i < m; ++i) alphahat[i + (size_t)m * t] += tmpv[i];
      pf_mm(&Pt[mm * t], Nt.data(), A.data(), m, m, m);
      pf_mm(A.data(), &Pt[mm * t], B.data(), m, m, m);
      pf_symmatu(B.data(), m);
      for (size_t e = 0; e < mm; ++e) Pt[mm * t + e] -= B[e];
    }
    for (size_t e = 0; e < mm; ++e) ccov[mm * n + e] = 0.0;

    conditional_cov_svd(m, n, Pt, ccov);
    L = Pt;
    Ct = ccov;
  }

  /* the head of psi_filter: mode_in == nullptr -> approximate() + approx_loglik (:514-528);
   * mode_in != nullptr -> approximate_for_is (:143-161) */
  void run(const double* mode_in) {
    run_head(mode_in);
    smoother_and_conditional();
  }
  /* the same without the conditional covariances (all the simulation smoother needs) */
  void run_head(const double* mode_in) {
    if (mode_in) {
      for (int t = 0; t < n; ++t) mode[t] = mode_in[t];
      laplace_iter(mode);
      update_scales();
      approx_loglik = 0.0;
    } else {
      approximate();
      const double gl = log_likelihood();
      update_scales();
      /* compute_const_term (:347-370) */
      double c = 0.0;
      const double phi = M->phi;
      auto lch = [](double nn, double kk) {
        return -std::log(nn + 1) - std::lgamma(nn - kk + 1) - std::lgamma(kk + 1) + std::lgamma(nn + 2);
      };
      int nobs = 0;
      double sumlog = 0.0;
      for (int t = 0; t < n; ++t) {
        const double y = M->y[t], u = M->u[t];
        if (!std::isfinite(y)) continue;
        nobs++;
        sumlog += pf_log(aH[t]);
        switch (M->distribution) {
          case 0: c += -0.5 * std::log(2.0 * M_PI) - std::log(phi); break;
          case 1: c += y * std::log(u) - std::lgamma(y + 1); break;
          case 2: c += lch(u, y); break;
          case 3: c += lch(y + phi - 1, y) + phi * std::log(phi) + y * std::log(u); break;
          case 4: c += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(y) - phi * std::log(u); break;
        }
      }
      const double norm = -0.5 * nobs * std::log(2 * M_PI) - sumlog;
      double a1 = 0.0, a2 = 0.0;
      int j;
      for (j = 1; j < n; j += 2) { a1 += scales[j - 1]; a2 += scales[j]; }
      if ((j - 1) < n) a1 += scales[j - 1];
      approx_loglik = gl + (c - norm) + (a1 + a2);
    }
  }
};

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic pop
#endif

} /* namespace pfga */
#endif
