/* mng_gauss_approx.h -- host-side Gaussian approximation of an ssm_mng model
 * (p observation series, m states; C++17, header only): everything
 * ssm_mng::psi_filter does BEFORE its particle loop
 * (/root/reference/src/model_ssm_mng.cpp:461-486).
 *
 *   approximate()            src/model_ssm_mng.cpp:72-118    Newton / Laplace sweeps
 *   approximate_for_is       src/model_ssm_mng.cpp:121-137
 *   update_scales            src/model_ssm_mng.cpp:216-256
 *   laplace_iter             src/model_ssm_mng.cpp:268-315
 *   compute_const_term       src/model_ssm_mng.cpp:317-355
 *   ssm_mlg::log_likelihood  src/model_ssm_mlg.cpp:90-137
 *   ssm_mlg::fast_smoother   src/model_ssm_mlg.cpp:220-306
 *   ssm_mlg::smoother_ccov   src/model_ssm_mlg.cpp:310-397
 *   conditional_cov (SVD)    src/conditional_dist.cpp:8-32 (shared with ung_gauss_approx.h)
 *
 * O(n (m^3 + p^3)), strictly sequential in t: host work, once per filter call.
 * Missing series of a time point are masked as the reference's smoothers do
 * (row of Z zeroed, unit variance, zero innovation); log_likelihood() works on
 * the observed sub-block in the reference, which yields the same numbers (the
 * masked rows contribute exact zeros / ones).  Chained products are associated
 * left to right with pf_smallmat.h; Cholesky and triangular inverse are plain
 * fixed-order loops standing in for LAPACK dpotrf / dtrtri [UNVENDORED].
 *
 * Shared by the product (bssm_b200/csrc/api.cu) and the oracle, like the
 * scalar functions of pf_math.h.
 */
#ifndef BSSM_B200_MNG_GAUSS_APPROX_H
#define BSSM_B200_MNG_GAUSS_APPROX_H

#include <cmath>
#include <limits>
#include <vector>

#include "bssm_b200.h"
#include "bssm_b200/pf_math.h"
#include "bssm_b200/pf_smallmat.h"
#include "bssm_b200/ung_gauss_approx.h" /* conditional_cov_svd */

namespace pfga {

/* upper Cholesky factor R (F = R'R) of a p x p matrix; false if not positive definite */
PF_SM bool chol_upper(const double* F, int p, double* R) {
  for (int e = 0; e < p * p; ++e) R[e] = 0.0;
  for (int j = 0; j < p; ++j)
    for (int i = 0; i <= j; ++i) {
      double s = F[i + (size_t)p * j];
      for (int l = 0; l < i; ++l) s -= R[l + (size_t)p * i] * R[l + (size_t)p * j];
      if (i == j) {
        if (!(s > 0.0)) return false;
        R[j + (size_t)p * j] = sqrt(s);
      } else {
        R[i + (size_t)p * j] = s / R[i + (size_t)p * i];
      }
    }
  return true;
}
/* X = R^{-1} for upper-triangular R */
PF_SM void inv_upper(const double* R, int p, double* X) {
  for (int e = 0; e < p * p; ++e) X[e] = 0.0;
  for (int j = 0; j < p; ++j) {
    X[j + (size_t)p * j] = 1.0 / R[j + (size_t)p * j];
    for (int i = j - 1; i >= 0; --i) {
      double s = 0.0;
      for (int l = i + 1; l <= j; ++l) s += R[i + (size_t)p * l] * X[l + (size_t)p * j];
      X[i + (size_t)p * j] = -s / R[i + (size_t)p * i];
    }
  }
}

struct MngApprox {
  const bssm_mng_model* M;
  int n, m, k, p;
  std::vector<double> mode, ay, aH, aHH; /* p x n each, element (j, t) at j + p*t */
  std::vector<double> scales;            /* n */
  double approx_loglik = 0.0;
  bool failed = false;                   /* a Cholesky of F_t failed */
  std::vector<double> alphahat, L, Ct;   /* m x (n+1), m x m x (n+1), m x m x (n+1) */

  explicit MngApprox(const bssm_mng_model* M_) : M(M_), n(M_->n), m(M_->m), k(M_->k), p(M_->p) {
    const size_t pn = (size_t)p * n;
    mode.assign(pn, 0.0); ay.assign(pn, 0.0); aH.assign(pn, 0.0); aHH.assign(pn, 0.0);
    scales.assign(n, 0.0);
  }
  double y(int j, int t) const { return M->y[t + (size_t)n * j]; }
  double u(int j, int t) const { return M->u[t + (size_t)n * j]; }
  const double* Z(int t) const { return M->Z + (size_t)p * m * (t * M->Ztv); }
  const double* T(int t) const { return M->T + (size_t)m * m * (t * M->Ttv); }
  const double* R(int t) const { return M->R + (size_t)m * k * (t * M->Rtv); }
  const double* C(int t) const { return M->C + (size_t)m * (t * M->Ctv); }
  const double* D(int t) const { return M->D + (size_t)p * (t * M->Dtv); }
  void RR(int t, double* out) const { pf_mmt(R(t), R(t), out, m, k, m); }

  /* src/model_ssm_mng.cpp:268-315 */
  void laplace_iter(const std::vector<double>& signal) {
    for (int j = 0; j < p; ++j) {
      const double phi = M->phi[j];
      for (int t = 0; t < n; ++t) {
        const size_t e = j + (size_t)p * t;
        const double yy = y(j, t), uu = u(j, t), s = signal[e];
        switch (M->distribution[j]) {
          case 1: {
            aHH[e] = 1.0 / (pf_exp(s) * uu);
            ay[e] = yy * aHH[e] + s - 1.0;
          } break;
          case 2: {
            const double ex = pf_exp(s);
            aHH[e] = (1.0 + ex) * (1.0 + ex) / (uu * ex);
            ay[e] = yy * aHH[e] + s - 1.0 - ex;
          } break;
          case 3: {
            const double ex = pf_exp(s) * uu;
            aHH[e] = (phi + ex) * (phi + ex) / (phi * ex * (yy + phi));
            ay[e] = s + (phi + ex) * (yy - ex) / ((yy + phi) * ex);
          } break;
          case 4: {
            const double ex = pf_exp(s) * uu;
            aHH[e] = ex / (yy * phi);
            ay[e] = s - ex / yy + 1;
          } break;
          default: {
            aHH[e] = phi * phi;
            ay[e] = yy;
          } break;
        }
        aH[e] = std::sqrt(aHH[e]);
      }
    }
  }

  /* log g(y | s) up to constants for series j, src/model_ssm_mng.cpp:225-250 */
  double log_obs(int j, double s, double yy, double uu) const {
    const double phi = M->phi[j];
    switch (M->distribution[j]) {
      case 1: return yy * s - uu * pf_exp(s);
      case 2: return yy * s - uu * pf_log1p(pf_exp(s));
      case 3: return yy * s - (yy + phi) * pf_log(phi + uu * pf_exp(s));
      case 4: return -phi * s - (yy * phi * pf_exp(-s) / uu);
      default: { const double q = (yy - s) / phi; return -0.5 * (q * q); }
    }
  }

  /* src/model_ssm_mng.cpp:216-256 */
  void update_scales() {
    for (int t = 0; t < n; ++t) {
      double sc = 0.0;
      for (int j = 0; j < p; ++j) {
        const size_t e = j + (size_t)p * t;
        if (std::isfinite(y(j, t))) {
          sc += log_obs(j, mode[e], y(j, t), u(j, t));
          const double q = (ay[e] - mode[e]) / aH[e];
          sc += 0.5 * (q * q);
        }
      }
      scales[t] = sc;
    }
  }

  /* one Kalman step with masking of the missing series (src/model_ssm_mlg.cpp:232-270) */
  struct Step {
    int nobs = 0;
    std::vector<double> Zt, hh, v, ZFinv, K, cholF, inv;
  };
  /* returns false when the Cholesky of F_t fails */
  bool gain(int t, const double* at, const double* Pt, Step& st) const {
    st.nobs = 0;
    st.Zt.assign(Z(t), Z(t) + (size_t)p * m);
    st.hh.assign(p, 1.0);
    for (int j = 0; j < p; ++j) {
      /* the approximating model's y is the pseudo-observation; it is missing where y is */
      if (std::isfinite(y(j, t))) { st.nobs++; st.hh[j] = aHH[j + (size_t)p * t]; }
      else for (int d = 0; d < m; ++d) st.Zt[j + (size_t)p * d] = 0.0;
    }
    if (st.nobs == 0) return true;
    std::vector<double> A((size_t)p * m), F((size_t)p * p);
    pf_mm(st.Zt.data(), Pt, A.data(), p, m, m);
    pf_mmt(A.data(), st.Zt.data(), F.data(), p, m, p);
    for (int j = 0; j < p; ++j) F[j + (size_t)p * j] += st.hh[j];
    for (int e = 0; e < p * p; ++e) if (!std::isfinite(F[e])) return false;
    for (int j = 0; j < p; ++j) if (!(F[j + (size_t)p * j] > 0)) return false;
    st.cholF.resize((size_t)p * p); st.inv.resize((size_t)p * p);
    if (!chol_upper(F.data(), p, st.cholF.data())) return false;
    inv_upper(st.cholF.data(), p, st.inv.data());
    st.v.assign(p, 0.0);
    std::vector<double> za(p);
    pf_mm(st.Zt.data(), at, za.data(), p, m, 1);
    for (int j = 0; j < p; ++j)
      st.v[j] = std::isfinite(y(j, t)) ? (ay[j + (size_t)p * t] - D(t)[j]) - za[j] : 0.0;
    std::vector<double> W((size_t)m * p);
    st.ZFinv.resize((size_t)m * p); st.K.resize((size_t)m * p);
    pf_mtm(st.Zt.data(), st.inv.data(), W.data(), m, p, p);      /* Zt' inv */
    pf_mmt(W.data(), st.inv.data(), st.ZFinv.data(), m, p, p);   /* (Zt' inv) inv' */
    pf_mm(Pt, st.ZFinv.data(), st.K.data(), m, m, p);
    return true;
  }
  void predict_mean(int t, const double* at, const Step& st, double* anext) const {
    std::vector<double> tmp(at, at + m), kv(m), ta(m);
    if (st.nobs > 0) {
      pf_mm(st.K.data(), st.v.data(), kv.data(), m, p, 1);
      for (int i = 0; i < m; ++i) tmp[i] = at[i] + kv[i];
    }
    pf_mm(T(t), tmp.data(), ta.data(), m, m, 1);
    for (int i = 0; i < m; ++i) anext[i] = C(t)[i] + ta[i];
  }
  void ikz(const Step& st, double* out) const { /* I - K Zt */
    std::vector<double> kz((size_t)m * m);
    pf_mm(st.K.data(), st.Zt.data(), kz.data(), m, p, m);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) out[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - kz[i + (size_t)m * j];
  }
  void predict_cov(int t, const double* Pt, const Step& st, double* Pnext) const {
    const size_t mm = (size_t)m * m;
    std::vector<double> X(mm), A(mm), B(mm), rr(mm);
    if (st.nobs > 0) {
      std::vector<double> tmp(mm), KH((size_t)m * p), khk(mm);
      ikz(st, tmp.data());
      pf_mm(tmp.data(), Pt, A.data(), m, m, m);
      pf_mmt(A.data(), tmp.data(), B.data(), m, m, m);
      for (int j = 0; j < p; ++j)
        for (int i = 0; i < m; ++i) KH[i + (size_t)m * j] = st.K[i + (size_t)m * j] * st.hh[j];
      pf_mmt(KH.data(), st.K.data(), khk.data(), m, p, m);
      for (size_t e = 0; e < mm; ++e) X[e] = B[e] + khk[e];
    } else {
      for (size_t e = 0; e < mm; ++e) X[e] = Pt[e];
    }
    pf_mm(T(t), X.data(), A.data(), m, m, m);
    pf_mmt(A.data(), T(t), B.data(), m, m, m);
    RR(t, rr.data());
    for (size_t e = 0; e < mm; ++e) Pnext[e] = B[e] + rr[e];
    pf_symmatu(Pnext, m);
  }

  /* ssm_mlg::log_likelihood, src/model_ssm_mlg.cpp:90-137 */
  double log_likelihood() const {
    double sumH = 0.0, sumR = 0.0;
    for (size_t e = 0; e < aH.size(); ++e) sumH += aH[e];
    const int nR = (M->Rtv ? n : 1) * m * k;
    for (int e = 0; e < nR; ++e) sumR += M->R[e];
    if (sumH + sumR < 1e-12) return -std::numeric_limits<double>::infinity();
    const size_t mm = (size_t)m * m;
    double logLik = 0.0;
    std::vector<double> at(M->a1, M->a1 + m), Pt(M->P1, M->P1 + mm), an(m), Pn(mm);
    const double LOG2PI = std::log(2.0 * M_PI);
    for (int t = 0; t < n; ++t) {
      Step st;
      if (!gain(t, at.data(), Pt.data(), st)) return -std::numeric_limits<double>::infinity();
      predict_mean(t, at.data(), st, an.data());
      predict_cov(t, Pt.data(), st, Pn.data());
      if (st.nobs > 0) {
        std::vector<double> Fv(p);
        pf_mtm(st.inv.data(), st.v.data(), Fv.data(), p, p, 1);
        double ld = 0.0, q = 0.0;
        for (int j = 0; j < p; ++j) ld += pf_log(st.cholF[j + (size_t)p * j]);
        for (int j = 0; j < p; ++j) q += Fv[j] * Fv[j];
        logLik -= 0.5 * (st.nobs * LOG2PI + 2.0 * ld + q);
      }
      at = an; Pt = Pn;
    }
    return logLik;
  }

  /* ssm_mlg::fast_smoother, src/model_ssm_mlg.cpp:220-306: at is m x (n+1) */
  bool fast_smoother(std::vector<double>& at) const {
    const size_t mm = (size_t)m * m;
    at.assign((size_t)m * (n + 1), 0.0);
    std::vector<double> Pt(M->P1, M->P1 + mm), Pn(mm);
    std::vector<Step> st(n);
    for (int i = 0; i < m; ++i) at[i] = M->a1[i];
    for (int t = 0; t < n; ++t) {
      if (!gain(t, &at[(size_t)m * t], Pt.data(), st[t])) return false;
      predict_mean(t, &at[(size_t)m * t], st[t], &at[(size_t)m * (t + 1)]);
      predict_cov(t, Pt.data(), st[t], Pn.data());
      Pt = Pn;
    }
    std::vector<double> rt((size_t)m * n, 0.0), Lm(mm), tmp(mm), lr(m), zv(m);
    auto back = [&](int t, const double* r_in, double* r_out) { /* r_{t-1} from r_t */
      if (st[t].nobs > 0) {
        ikz(st[t], tmp.data());
        pf_mm(T(t), tmp.data(), Lm.data(), m, m, m);
        pf_mtm(Lm.data(), r_in, lr.data(), m, m, 1);
        pf_mm(st[t].ZFinv.data(), st[t].v.data(), zv.data(), m, p, 1);
        for (int i = 0; i < m; ++i) r_out[i] = zv[i] + lr[i];
      } else {
        pf_mtm(T(t), r_in, r_out, m, m, 1);
      }
    };
    for (int t = n - 1; t > 0; --t) back(t, &rt[(size_t)m * t], &rt[(size_t)m * (t - 1)]);
    std::vector<double> w(m), pw(m);
    if (st[0].nobs > 0) {
      back(0, &rt[0], w.data());
      pf_mm(M->P1, w.data(), pw.data(), m, m, 1);
    } else if (m == 1) {
      pw[0] = M->P1[0] * T(0)[0] * rt[0]; /* (P1 T') rt for 1 x 1 operands, see ung_gauss_approx.h */
    } else {
      pf_mtm(T(0), &rt[0], w.data(), m, m, 1);
      pf_mm(M->P1, w.data(), pw.data(), m, m, 1);
    }
    for (int i = 0; i < m; ++i) at[i] = M->a1[i] + pw[i];
    std::vector<double> ta(m), rr(mm), rrr(m);
    for (int t = 0; t < n - 1; ++t) {
      pf_mm(T(t), &at[(size_t)m * t], ta.data(), m, m, 1);
      RR(t, rr.data());
      pf_mm(rr.data(), &rt[(size_t)m * t], rrr.data(), m, m, 1);
      for (int i = 0; i < m; ++i) at[i + (size_t)m * (t + 1)] = C(t)[i] + ta[i] + rrr[i];
    }
    return true;
  }

  /* ssm_mng::approximate, src/model_ssm_mng.cpp:72-118 (iteration branch) */
  void approximate() {
    const size_t pn = (size_t)p * n;
    for (int t = 0; t < n; ++t)
      for (int j = 0; j < p; ++j) mode[j + (size_t)p * t] = M->initial_mode[t + (size_t)n * j];
    int i = 0;
    double diff = M->conv_tol + 1;
    std::vector<double> at, mode_new(pn), za(p);
    while (i < M->max_iter && diff > M->conv_tol) {
      i++;
      laplace_iter(mode);
      if (!fast_smoother(at)) { failed = true; return; }
      for (int t = 0; t < n; ++t) {
        pf_mm(Z(t), &at[(size_t)m * t], za.data(), p, m, 1);
        for (int j = 0; j < p; ++j) mode_new[j + (size_t)p * t] = D(t)[j] + za[j];
      }
      double a1 = 0.0, a2 = 0.0; /* arma::accu, two accumulators [UNVENDORED] */
      size_t j;
      for (j = 1; j < pn; j += 2) {
        a1 += (mode_new[j - 1] - mode[j - 1]) * (mode_new[j - 1] - mode[j - 1]);
        a2 += (mode_new[j] - mode[j]) * (mode_new[j] - mode[j]);
      }
      if ((j - 1) < pn) a1 += (mode_new[j - 1] - mode[j - 1]) * (mode_new[j - 1] - mode[j - 1]);
      diff = (a1 + a2) / (double)(n * p);
      mode = mode_new;
    }
  }

  /* ssm_mlg::smoother_ccov (:310-397) followed by conditional_cov (SVD branch) */
  void smoother_and_conditional() {
    const size_t mm = (size_t)m * m;
    alphahat.assign((size_t)m * (n + 1), 0.0);
    std::vector<double> Pt(mm * (n + 1), 0.0), ccov(mm * (n + 1), 0.0);
    std::vector<Step> st(n);
    for (int i = 0; i < m; ++i) alphahat[i] = M->a1[i];
    for (size_t e = 0; e < mm; ++e) Pt[e] = M->P1[e];
    for (int t = 0; t < n; ++t) {
      if (!gain(t, &alphahat[(size_t)m * t], &Pt[mm * t], st[t])) {
        failed = true;
        const double inf = std::numeric_limits<double>::infinity();
        alphahat.assign(alphahat.size(), inf); Pt.assign(Pt.size(), inf); ccov.assign(ccov.size(), inf);
        L = Pt; Ct = ccov;
        return;
      }
      predict_mean(t, &alphahat[(size_t)m * t], st[t], &alphahat[(size_t)m * (t + 1)]);
      predict_cov(t, &Pt[mm * t], st[t], &Pt[mm * (t + 1)]);
      for (size_t e = 0; e < mm; ++e) ccov[mm * t + e] = Pt[mm * (t + 1) + e];
    }
    std::vector<double> rt(m, 0.0), Nt(mm, 0.0), Lm(mm), A(mm), B(mm), E(mm), lr(m), tmpv(m), tmp(mm), zv(m), zz(mm);
    for (int t = n - 1; t >= 0; --t) {
      const double* Tm;
      const bool used = st[t].nobs > 0;
      if (used) {
        ikz(st[t], tmp.data());
        pf_mm(T(t), tmp.data(), Lm.data(), m, m, m);
        Tm = Lm.data();
      } else {
        Tm = T(t);
      }
      /* ccov_t = Pt_t * L' * (I - Nt * ccov_t)   (L = T when nothing is observed at t) */
      pf_mm(Nt.data(), &ccov[mm * t], A.data(), m, m, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - A[i + (size_t)m * j];
      pf_mmt(&Pt[mm * t], Tm, B.data(), m, m, m);
      pf_mm(B.data(), E.data(), &ccov[mm * t], m, m, m);
      pf_mtm(Tm, rt.data(), lr.data(), m, m, 1);
      pf_mtm(Tm, Nt.data(), A.data(), m, m, m);
      pf_mm(A.data(), Tm, B.data(), m, m, m);
      if (used) {
        pf_mm(st[t].ZFinv.data(), st[t].v.data(), zv.data(), m, p, 1);
        pf_mm(st[t].ZFinv.data(), st[t].Zt.data(), zz.data(), m, p, m);
        for (int i = 0; i < m; ++i) rt[i] = zv[i] + lr[i];
        for (size_t e = 0; e < mm; ++e) Nt[e] = zz[e] + B[e];
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

  /* compute_const_term, src/model_ssm_mng.cpp:317-355 */
  double const_term() const {
    auto lch = [](double nn, double kk) {
      return -std::log(nn + 1) - std::lgamma(nn - kk + 1) - std::lgamma(kk + 1) + std::lgamma(nn + 2);
    };
    double c = 0.0;
    for (int j = 0; j < p; ++j) {
      const double phi = M->phi[j];
      int nobs = 0;
      for (int t = 0; t < n; ++t) {
        const double yy = y(j, t), uu = u(j, t);
        if (!std::isfinite(yy)) continue;
        nobs++;
        switch (M->distribution[j]) {
          case 1: c += -std::lgamma(yy + 1) + yy * std::log(uu); break;
          case 2: c += lch(uu, yy); break;
          case 3: c += lch(yy + phi - 1, yy) + phi * std::log(phi) + yy * std::log(uu); break;
          case 4: c += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(yy) - phi * std::log(uu); break;
          default: break;
        }
      }
      if (M->distribution[j] == 5) c += nobs * (-0.5 * std::log(2.0 * M_PI) - std::log(phi));
      for (int t = 0; t < n; ++t)
        if (std::isfinite(y(j, t))) c -= -0.5 * std::log(2.0 * M_PI) - pf_log(aH[j + (size_t)p * t]);
    }
    return c;
  }

  /* the head of psi_filter: mode_in == nullptr -> approximate() + approx_loglik (:464-477);
   * mode_in (p x n) != nullptr -> approximate_for_is (:121-137) */
  void run(const double* mode_in) {
    if (mode_in) {
      for (size_t e = 0; e < mode.size(); ++e) mode[e] = mode_in[e];
      laplace_iter(mode);
      update_scales();
      approx_loglik = 0.0;
    } else {
      approximate();
      if (failed) { approx_loglik = -std::numeric_limits<double>::infinity(); }
      const double gl = log_likelihood();
      update_scales();
      const double c = const_term();
      double a1 = 0.0, a2 = 0.0;
      int j;
      for (j = 1; j < n; j += 2) { a1 += scales[j - 1]; a2 += scales[j]; }
      if ((j - 1) < n) a1 += scales[j - 1];
      approx_loglik = gl + c + (a1 + a2);
    }
    smoother_and_conditional();
  }
};

} /* namespace pfga */
#endif
