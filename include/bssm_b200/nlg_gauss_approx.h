/* nlg_gauss_approx.h -- host-side Gaussian approximation of an ssm_nlg model with
 * one observation series (p = 1), for the psi-APF: what ssm_nlg::psi_filter needs
 * before its particle loop (/root/reference/src/model_ssm_nlg.cpp:971-993).
 *
 *   ekf / iterated ekf      src/model_ssm_nlg.cpp:284-378
 *   approximate_by_ekf      src/model_ssm_nlg.cpp:188-222
 *   approximate             src/model_ssm_nlg.cpp:62-150  (relinearisation sweeps with
 *                                                          log_signal_pdf convergence and
 *                                                          step halving)
 *   approximate_for_is      src/model_ssm_nlg.cpp:153-186
 *   log_signal_pdf          src/model_ssm_nlg.cpp:1295-1319
 *   update_scales           src/model_ssm_nlg.cpp:854-879
 *   dmvnorm                 src/dmvnorm.cpp:4-60
 * The linear-Gaussian approximating model (class ssm_mlg in the reference) has
 * time-varying Z_t, T_t, D_t, C_t, H_t, R_t; with p = 1 its Kalman recursions are
 * those of ung_gauss_approx.h (UngApprox with the pseudo-observations replaced by
 * y and H_t^2), which this header reuses: log_likelihood, fast_smoother,
 * smoother_ccov, conditional_cov.
 *
 * The user functions (function pointers compiled by R in the reference,
 * src/model_ssm_nlg.h:10-21) come from the registry: NlgFns evaluates the model
 * selected by model_id with the same formulas as the device functors of
 * bssm_b200/csrc/models_nlg_sde.cuh.
 */
#ifndef BSSM_B200_NLG_GAUSS_APPROX_H
#define BSSM_B200_NLG_GAUSS_APPROX_H

#include <cmath>
#include <limits>
#include <vector>

#include "bssm_b200.h"
#include "bssm_b200/pf_math.h"
#include "bssm_b200/pf_smallmat.h"
#include "bssm_b200/ung_gauss_approx.h"

/* member functions that also run in device code */
#if defined(__CUDACC__)
#define PF_HDM __host__ __device__
#else
#define PF_HDM
#endif

namespace pfga {

/* registry functors on the host; growth model: R/cpp_example_models.R:350-460 */
struct NlgFns {
  int model_id;
  const double* theta;
  const double* known;
  /* registry (include/bssm_b200.h): 0 = logistic growth (m = k = 2), R/cpp_example_models.R:336-511;
   * 1 = AR(1) state observed through exp(alpha) (m = k = 1), R/cpp_example_models.R:194-335,
   *     theta = (mu, rho, log sigma, log H) */
  /* 2 = sine transition observed through exp(alpha) (m = k = 1), R/cpp_example_models.R:635-757,
   *     theta = (log H, log R);
   * 3 = linear-Gaussian random walk (m = k = 1), R/cpp_example_models.R:512-634, theta = (log H) */
  PF_HDM int m() const { return model_id == 0 ? 2 : 1; }
  PF_HDM int k() const { return model_id == 0 ? 2 : 1; }
  PF_HDM void a1_fn(double* a1) const {
    if (model_id == 2 || model_id == 3) { a1[0] = 0; return; }
    if (model_id == 1) { a1[0] = theta[0]; return; }
    a1[0] = known[2]; a1[1] = known[3];
  }
  PF_HDM void P1_fn(double* P1) const {
    if (model_id == 2 || model_id == 3) { P1[0] = 1; return; }
    if (model_id == 1) { const double e = pf_exp(theta[2]); P1[0] = e * e / (1 - theta[1] * theta[1]); return; }
    P1[0] = known[4]; P1[1] = 0; P1[2] = 0; P1[3] = known[5];
  }
  PF_HDM double H_fn(int, const double*) const { return model_id == 1 ? pf_exp(theta[3]) : pf_exp(theta[0]); }
  PF_HDM void R_fn(int, const double*, double* R) const {
    if (model_id == 3) { R[0] = 1; return; }
    if (model_id == 2) { R[0] = pf_exp(theta[1]); return; }
    if (model_id == 1) { R[0] = pf_exp(theta[2]); return; }
    R[0] = pf_exp(theta[1]); R[1] = 0; R[2] = 0; R[3] = pf_exp(theta[2]);
  }
  PF_HDM double Z_fn(int, const double* alpha) const { return model_id == 3 ? alpha[0] : (model_id != 0 ? pf_exp(alpha[0]) : alpha[1]); }
  PF_HDM void Z_gn(int, const double* alpha, double* Zg) const {
    if (model_id == 3) { Zg[0] = 1.0; return; }
    if (model_id != 0) { Zg[0] = pf_exp(alpha[0]); return; }
    Zg[0] = 0.0; Zg[1] = 1.0;
  }
  PF_HDM void T_fn(int, const double* alpha, double* an) const {
    if (model_id == 3) { an[0] = alpha[0]; return; }
    if (model_id == 2) { an[0] = pf_sin(alpha[0]); return; }
    if (model_id == 1) { an[0] = theta[0] * (1 - theta[1]) + theta[1] * alpha[0]; return; }
    const double dT = known[0], K = known[1];
    an[0] = alpha[0];
    const double r = pf_exp(alpha[0]) / (1.0 + pf_exp(alpha[0]));
    an[1] = K * alpha[1] * pf_exp(r * dT) / (K + alpha[1] * (pf_exp(r * dT) - 1));
  }
  PF_HDM void T_gn(int, const double* alpha, double* Tg) const { /* column-major m x m */
    if (model_id == 3) { Tg[0] = 1.0; return; }
    if (model_id == 2) { Tg[0] = pf_cos(alpha[0]); return; }
    if (model_id == 1) { Tg[0] = theta[1]; return; }
    const double dT = known[0], K = known[1];
    const double r = pf_exp(alpha[0]) / (1 + pf_exp(alpha[0]));
    const double den = K + alpha[1] * (pf_exp(r * dT) - 1);
    const double tmp = pf_exp(r * dT) / (den * den);
    Tg[0] = 1.0;                                                                       /* (0,0) */
    Tg[2] = 0;                                                                         /* (0,1) */
    Tg[1] = dT * K * alpha[1] * (K - alpha[1]) * tmp * r / (1 + pf_exp(alpha[0]));     /* (1,0) */
    Tg[3] = K * K * tmp;                                                               /* (1,1) */
  }
};

/* log N(x; mean, .) pieces of dmvnorm (src/dmvnorm.cpp:4-60) for a full-rank m x m case:
 * rooti = inv(chol(S, lower)), cst = sum(log(diag(rooti))) - 0.5 * m * log(2 pi) */
struct MvnFactor {
  int m;
  std::vector<double> rooti;
  double cst;
  double var1 = 0.0; /* m == 1: dmvnorm takes its scalar branch (src/dmvnorm.cpp:9-16, lwr = false) */
  void set(const double* S, int m_) { /* S symmetric positive definite (diag > eps assumed) */
    m = m_;
    if (m == 1) { var1 = S[0]; rooti.assign(1, 0.0); cst = std::log(2.0 * M_PI) + pf_log(var1); return; }
    std::vector<double> Lc((size_t)m * m, 0.0);
    for (int j = 0; j < m; ++j)
      for (int i = j; i < m; ++i) {
        double s = S[i + (size_t)m * j];
        for (int l = 0; l < j; ++l) s -= Lc[i + (size_t)m * l] * Lc[j + (size_t)m * l];
        Lc[i + (size_t)m * j] = (i == j) ? std::sqrt(s) : s / Lc[j + (size_t)m * j];
      }
    rooti.assign((size_t)m * m, 0.0); /* inverse of a lower-triangular matrix by forward substitution */
    for (int j = 0; j < m; ++j) {
      rooti[j + (size_t)m * j] = 1.0 / Lc[j + (size_t)m * j];
      for (int i = j + 1; i < m; ++i) {
        double s = 0.0;
        for (int l = j; l < i; ++l) s += Lc[i + (size_t)m * l] * rooti[l + (size_t)m * j];
        rooti[i + (size_t)m * j] = -s / Lc[i + (size_t)m * i];
      }
    }
    double sl = 0.0;
    for (int i = 0; i < m; ++i) sl += pf_log(rooti[i + (size_t)m * i]);
    cst = sl - 0.5 * m * std::log(2.0 * M_PI);
  }
  double logpdf(const double* x, const double* mean) const {
    if (m == 1) return -0.5 * (cst + (x[0] - mean[0]) * (x[0] - mean[0]) / var1);
    double dot = 0.0;
    for (int i = 0; i < m; ++i) {
      double z2 = 0.0;
      for (int l = 0; l < m; ++l) z2 += rooti[i + (size_t)m * l] * (x[l] - mean[l]);
      dot += z2 * z2;
    }
    return cst - 0.5 * dot;
  }
};

/* dmvnorm for one series: lwr = true takes a standard deviation, lwr = false a variance
 * (src/dmvnorm.cpp:9-13) */
inline double dnorm1_sd(double x, double mean, double sd) {
  const double q = (x - mean) / sd;
  return -0.5 * (std::log(2.0 * M_PI) + 2.0 * pf_log(sd) + q * q);
}
inline double dnorm1_var(double x, double mean, double var) {
  return -0.5 * (std::log(2.0 * M_PI) + pf_log(var) + (x - mean) * (x - mean) / var);
}

struct NlgApprox {
  const bssm_nlg_model* M;
  NlgFns fn;
  int n, m, k;
  int max_iter = 100;
  int iekf_iter = 0;
  double conv_tol = 1e-8;
  /* approximating linear-Gaussian model, everything time varying */
  std::vector<double> Z, T, R, C, D, H, a1, P1;
  std::vector<double> mode;   /* m x n */
  std::vector<double> scales; /* n */
  double approx_loglik = 0.0;
  bool finite_mode = true;
  std::vector<double> alphahat, L, Ct; /* m x (n+1), m x m x (n+1) twice */
  /* per-time factors of the transition densities used in the particle weights (index t = 1..n-1
   * uses slot t-1): approximate model N(.; C + T a, RR) and true model N(.; T_fn(a), R R') */
  std::vector<MvnFactor> f_approx, f_true;

  explicit NlgApprox(const bssm_nlg_model* M_)
      : M(M_), fn{M_->model_id, M_->theta, M_->known_params}, n(M_->n), m(fn.m()), k(fn.k()) {
    const size_t mm = (size_t)m * m;
    Z.assign((size_t)m * n, 0.0); T.assign(mm * n, 0.0); R.assign((size_t)m * k * n, 0.0);
    C.assign((size_t)m * n, 0.0); D.assign(n, 0.0); H.assign(n, 0.0);
    a1.assign(m, 0.0); P1.assign(mm, 0.0);
    mode.assign((size_t)m * n, 0.0); scales.assign(n, 0.0);
  }

  /* linearise at the columns of `zs` (for Z, D, H) and `ts` (for T, C, R):
   * approximate_by_ekf uses (at, att), the sweeps and approximate_for_is use (mode, mode) */
  void linearise(const double* zs, const double* ts) {
    std::vector<double> tv(m), tmpv(m);
    for (int t = 0; t < n; ++t) {
      const double* az = zs + (size_t)m * t;
      const double* at = ts + (size_t)m * t;
      fn.Z_gn(t, az, &Z[(size_t)m * t]);
      H[t] = fn.H_fn(t, az);
      fn.T_gn(t, at, &T[(size_t)m * m * t]);
      fn.R_fn(t, at, &R[(size_t)m * k * t]);
      double za = 0.0;
      for (int d = 0; d < m; ++d) za += Z[d + (size_t)m * t] * az[d];
      D[t] = fn.Z_fn(t, az) - za;
      fn.T_fn(t, at, tv.data());
      pf_mm(&T[(size_t)m * m * t], at, tmpv.data(), m, m, 1);
      for (int d = 0; d < m; ++d) C[d + (size_t)m * t] = tv[d] - tmpv[d];
    }
  }

  /* the approximating model seen through the univariate-observation Kalman code */
  struct View {
    bssm_ung_model U;
    UngApprox ga;
    explicit View(NlgApprox& A) : U(A.view()), ga(&U) {
      for (int t = 0; t < A.n; ++t) { ga.ay[t] = A.M->y[t]; ga.aH[t] = A.H[t]; ga.aHH[t] = A.H[t] * A.H[t]; }
    }
  };
  bssm_ung_model view() {
    bssm_ung_model U{};
    U.n = n; U.m = m; U.k = k; U.distribution = 1;
    U.Ztv = U.Ttv = U.Rtv = U.Dtv = U.Ctv = 1;
    U.y = M->y; U.Z = Z.data(); U.T = T.data(); U.R = R.data(); U.a1 = a1.data(); U.P1 = P1.data();
    U.D = D.data(); U.C = C.data(); U.xbeta = nullptr; U.u = nullptr; U.phi = 1.0;
    U.max_iter = max_iter; U.conv_tol = conv_tol; U.local_approx = 1; U.initial_mode = nullptr;
    return U;
  }

  /* ekf, iterated when iekf_iter > 0 (:284-378): at m x (n+1), att m x n; *loglik (optional) receives the
   * EKF log-likelihood (:369-371), -inf when an innovation variance is not positive */
  bool ekf(std::vector<double>& at, std::vector<double>& att, double* loglik = nullptr) const {
    double ll = 0.0;
    if (loglik) *loglik = -std::numeric_limits<double>::infinity();
    const size_t mm = (size_t)m * m;
    at.assign((size_t)m * (n + 1), 0.0); att.assign((size_t)m * n, 0.0);
    std::vector<double> Pt(mm), Ptt(mm), Zg(m), K(m), Tg(mm), Rt((size_t)m * k), A(mm), B(mm), rr(mm), tmp(mm), zp(m);
    fn.a1_fn(at.data());
    fn.P1_fn(Pt.data());
    for (int t = 0; t < n; ++t) {
      const double* a = &at[(size_t)m * t];
      double* af = &att[(size_t)m * t];
      if (std::isfinite(M->y[t])) {
        fn.Z_gn(t, a, Zg.data());
        const double Hs = fn.H_fn(t, a);
        const double HHt = Hs * Hs;
        pf_mtm(Zg.data(), Pt.data(), zp.data(), 1, m, m);
        double F = 0.0;
        for (int i = 0; i < m; ++i) F += zp[i] * Zg[i];
        F = F + HHt;
        if (!(std::isfinite(F) && F > 0)) return false;
        const double cholF = std::sqrt(F), inv = 1.0 / cholF;
        const double v = M->y[t] - fn.Z_fn(t, a);
        pf_mm(Pt.data(), Zg.data(), K.data(), m, m, 1); /* Pt Zg' inv_cholF inv_cholF' */
        for (int i = 0; i < m; ++i) K[i] = K[i] * inv * inv;
        for (int i = 0; i < m; ++i) af[i] = a[i] + K[i] * v;
        /* iterated EKF (:325-356): relinearise the observation equation at the updated state */
        double HHi = HHt;
        double v_last = v, cholF_last = cholF; /* what the log-likelihood term sees after the iterations */
        {
          double diff = 1.0;
          int it = 0;
          std::vector<double> anew(m);
          while (diff > 1e-4 && it < iekf_iter) {
            it++;
            fn.Z_gn(t, af, Zg.data());
            const double Hi = fn.H_fn(t, af);
            HHi = Hi * Hi;
            pf_mtm(Zg.data(), Pt.data(), zp.data(), 1, m, m);
            double Fi = 0.0;
            for (int i = 0; i < m; ++i) Fi += zp[i] * Zg[i];
            Fi = Fi + HHi;
            if (!(std::isfinite(Fi) && Fi > 0)) return false;
            const double invi = 1.0 / std::sqrt(Fi);
            cholF_last = std::sqrt(Fi);
            double zd = 0.0;
            for (int i = 0; i < m; ++i) zd += Zg[i] * (a[i] - af[i]);
            const double vi = M->y[t] - fn.Z_fn(t, af) - zd;
            v_last = vi;
            pf_mm(Pt.data(), Zg.data(), K.data(), m, m, 1);
            for (int i = 0; i < m; ++i) K[i] = K[i] * invi * invi;
            for (int i = 0; i < m; ++i) anew[i] = a[i] + K[i] * vi;
            double a1s = 0.0, a2s = 0.0; /* arma::mean(arma::square(.)): two accumulators, then / m */
            int j;
            for (j = 1; j < m; j += 2) {
              a1s += (af[j - 1] - anew[j - 1]) * (af[j - 1] - anew[j - 1]);
              a2s += (af[j] - anew[j]) * (af[j] - anew[j]);
            }
            if ((j - 1) < m) a1s += (af[j - 1] - anew[j - 1]) * (af[j - 1] - anew[j - 1]);
            diff = (a1s + a2s) / (double)m;
            for (int i = 0; i < m; ++i) af[i] = anew[i];
          }
        }
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) tmp[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - K[i] * Zg[j];
        pf_mm(tmp.data(), Pt.data(), A.data(), m, m, m);
        pf_mmt(A.data(), tmp.data(), B.data(), m, m, m);
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) Ptt[i + (size_t)m * j] = B[i + (size_t)m * j] + K[i] * HHi * K[j];
        const double Fv = v_last / cholF_last;
        ll -= 0.5 * (std::log(2.0 * M_PI) + 2.0 * std::log(cholF_last) + Fv * Fv);
      } else {
        for (int i = 0; i < m; ++i) af[i] = a[i];
        Ptt = Pt;
      }
      fn.T_fn(t, af, &at[(size_t)m * (t + 1)]);
      fn.T_gn(t, af, Tg.data());
      fn.R_fn(t, af, Rt.data());
      pf_mm(Tg.data(), Ptt.data(), A.data(), m, m, m);
      pf_mmt(A.data(), Tg.data(), B.data(), m, m, m);
      pf_mmt(Rt.data(), Rt.data(), rr.data(), m, k, m);
      for (size_t e = 0; e < mm; ++e) Pt[e] = B[e] + rr[e];
    }
    if (loglik) *loglik = ll;
    return true;
  }

  /* log_signal_pdf (:1295-1319): joint log density of states and observations */
  double log_signal_pdf(const double* alpha) const {
    std::vector<double> a1v(m), P1v((size_t)m * m), mean(m), Rt((size_t)m * k), cov((size_t)m * m);
    fn.a1_fn(a1v.data());
    fn.P1_fn(P1v.data());
    MvnFactor f;
    f.set(P1v.data(), m);
    double ll = f.logpdf(alpha, a1v.data());
    if (std::isfinite(M->y[0])) ll += dnorm1_sd(M->y[0], fn.Z_fn(0, alpha), fn.H_fn(0, alpha));
    for (int t = 0; t < n - 1; ++t) {
      const double* a = alpha + (size_t)m * t;
      fn.T_fn(t, a, mean.data());
      fn.R_fn(t, a, Rt.data());
      pf_mmt(Rt.data(), Rt.data(), cov.data(), m, k, m);
      f.set(cov.data(), m);
      ll += f.logpdf(a + m, mean.data());
      if (std::isfinite(M->y[t + 1])) ll += dnorm1_sd(M->y[t + 1], fn.Z_fn(t + 1, a + m), fn.H_fn(t + 1, a + m));
    }
    return ll;
  }

  bool all_finite(const std::vector<double>& v) const {
    for (double x : v) if (!std::isfinite(x)) return false;
    return true;
  }

  /* approximate (:62-150) */
  void approximate() {
    std::vector<double> at, att;
    fn.a1_fn(a1.data());
    fn.P1_fn(P1.data());
    if (!ekf(at, att)) { finite_mode = false; return; }
    linearise(at.data(), att.data()); /* approximate_by_ekf */
    std::vector<double> sm;
    { View v(*this); v.ga.fast_smoother(sm); }
    mode.assign(sm.begin(), sm.begin() + (size_t)m * n);
    if (!all_finite(mode)) { finite_mode = false; return; }
    double ll = 0.0;
    if (max_iter > 0) ll = log_signal_pdf(mode.data());
    int i = 0;
    double rel_diff = 1.0e300, abs_diff = 1;
    while (i < max_iter && rel_diff > conv_tol && abs_diff > 1e-4) {
      i++;
      linearise(mode.data(), mode.data());
      { View v(*this); v.ga.fast_smoother(sm); }
      std::vector<double> mode_new(sm.begin(), sm.begin() + (size_t)m * n);
      double ll_new = log_signal_pdf(mode_new.data());
      abs_diff = ll_new - ll;
      rel_diff = abs_diff / std::fabs(ll);
      if (!all_finite(mode_new) || !std::isfinite(ll_new)) { finite_mode = false; return; }
      if (rel_diff < -conv_tol && i > 1 && abs_diff > 1e-4) {
        /* unreachable as written in the reference (abs_diff > 1e-4 contradicts rel_diff < 0);
         * kept for fidelity */
        int ii = 0;
        double step_size = 1.0;
        std::vector<double> mode_old = mode;
        while (rel_diff < -conv_tol && ii < 15 && abs_diff > 1e-4) {
          step_size = step_size / 2.0;
          for (size_t e = 0; e < mode.size(); ++e) mode[e] = (1.0 - step_size) * mode_old[e] + step_size * mode_new[e];
          ll_new = log_signal_pdf(mode.data());
          abs_diff = ll_new - ll;
          rel_diff = abs_diff / std::fabs(ll);
          ii++;
          if (!all_finite(mode) || !std::isfinite(ll_new)) { finite_mode = false; return; }
        }
        if (ii == 15) { finite_mode = false; return; }
        mode_new = mode;
      }
      mode = mode_new;
      ll = ll_new;
    }
  }

  /* update_scales (:854-879) */
  void update_scales() {
    std::vector<double> mean(m), Rt((size_t)m * k), cov((size_t)m * m), amean(m), tmpv(m), rr((size_t)m * m);
    for (int t = 0; t < n; ++t) {
      scales[t] = 0.0;
      const double* mt = &mode[(size_t)m * t];
      if (std::isfinite(M->y[t])) {
        double za = 0.0;
        for (int d = 0; d < m; ++d) za += Z[d + (size_t)m * t] * mt[d];
        scales[t] = dnorm1_sd(M->y[t], fn.Z_fn(t, mt), fn.H_fn(t, mt)) - dnorm1_var(M->y[t], D[t] + za, H[t] * H[t]);
      }
    }
    for (int t = 1; t < n; ++t) {
      const double* mp = &mode[(size_t)m * (t - 1)];
      fn.T_fn(t - 1, mp, mean.data());
      fn.R_fn(t - 1, mp, Rt.data());
      pf_mmt(Rt.data(), Rt.data(), cov.data(), m, k, m);
      pf_mm(&T[(size_t)m * m * (t - 1)], mp, tmpv.data(), m, m, 1);
      for (int d = 0; d < m; ++d) amean[d] = C[d + (size_t)m * (t - 1)] + tmpv[d];
      pf_mmt(&R[(size_t)m * k * (t - 1)], &R[(size_t)m * k * (t - 1)], rr.data(), m, k, m);
      MvnFactor ft, fa;
      ft.set(cov.data(), m);
      fa.set(rr.data(), m);
      scales[t] += ft.logpdf(&mode[(size_t)m * t], mean.data()) - fa.logpdf(&mode[(size_t)m * t], amean.data());
    }
  }

  /* mode_in == nullptr: approximate() + approx_loglik (:973-981).
   * mode_in != nullptr: approximate_for_is (:153-186) -- relinearise at the stored mode and
   * nothing else: unlike its ssm_ung namesake it neither recomputes the scales nor touches
   * approx_loglik, so inside the IS-correction (fresh model object, src/R_postcorrection.cpp:
   * 290-300) both stay at their initial value zero.  Mirrored here. */
  void run(const double* mode_in) {
    fn.a1_fn(a1.data());
    fn.P1_fn(P1.data());
    if (mode_in) {
      mode.assign(mode_in, mode_in + (size_t)m * n);
      linearise(mode.data(), mode.data());
    } else {
      approximate();
      if (!finite_mode) return;
    }
    View v(*this);
    if (mode_in) {
      scales.assign(n, 0.0);
      approx_loglik = 0.0;
    } else {
      const double gl = v.ga.log_likelihood();
      update_scales();
      double acc1 = 0.0, acc2 = 0.0; /* arma::accu(scales) */
      int j;
      for (j = 1; j < n; j += 2) { acc1 += scales[j - 1]; acc2 += scales[j]; }
      if ((j - 1) < n) acc1 += scales[j - 1];
      approx_loglik = gl + (acc1 + acc2);
    }
    v.ga.smoother_and_conditional();
    alphahat = v.ga.alphahat; L = v.ga.L; Ct = v.ga.Ct;
    /* transition-density factors for the particle weights (log_weights :915-931) */
    f_approx.resize(n > 1 ? n - 1 : 0);
    f_true.resize(n > 1 ? n - 1 : 0);
    std::vector<double> rr((size_t)m * m), Rt((size_t)m * k), cov((size_t)m * m);
    for (int t = 0; t + 1 < n; ++t) {
      pf_mmt(&R[(size_t)m * k * t], &R[(size_t)m * k * t], rr.data(), m, k, m);
      f_approx[t].set(rr.data(), m);
      /* the registry models have state-independent R_fn; evaluated at the mode */
      fn.R_fn(t, &mode[(size_t)m * t], Rt.data());
      pf_mmt(Rt.data(), Rt.data(), cov.data(), m, k, m);
      f_true[t].set(cov.data(), m);
    }
  }
};

} /* namespace pfga */
#endif
