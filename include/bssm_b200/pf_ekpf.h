/* pf_ekpf.h -- per-particle pieces of the EKF-based particle filter of ssm_nlg
 * (ssm_nlg::ekf_filter, /root/reference/src/model_ssm_nlg.cpp:1150-1260) for one
 * observation series and M in {1, 2} states: the EKF update step
 * (ekf_update_step, :1262-1291), psd_chol (src/psd_chol.cpp:5-18) and the two
 * dmvnorm calls of the importance weights (src/dmvnorm.cpp:4-36).
 *
 * C++ header shared by the device policy (bssm_b200/csrc/models_nlg_sde.cuh) and
 * the CPU oracle, like pf_math.h: every function is straight-line IEEE
 * arithmetic in a fixed order, so host and device agree bit for bit.
 */
#ifndef BSSM_B200_PF_EKPF_H
#define BSSM_B200_PF_EKPF_H

#include "bssm_b200/pf_math.h"

#define PF_LOG2PI 1.8378770664093453 /* log(2 pi) */
#define PF_EPS 2.220446049250313e-16

/* arma::chol(S, "lower") for the rows/columns of S (M x M, column-major) whose diagonal exceeds
 * `thr`, zeros elsewhere: psd_chol when thr = max(eps, eps * M * max diag) */
template <int M>
PF_HD void pf_chol_lower_nz(const double* S, double thr, double* Lc) {
  if (M == 1) {
    Lc[0] = S[0] > thr ? sqrt(S[0]) : 0.0;
    return;
  }
  const bool n0 = S[0] > thr, n1 = S[3] > thr;
  Lc[0] = 0.0; Lc[1] = 0.0; Lc[2] = 0.0; Lc[3] = 0.0;
  if (n0 && n1) {
    Lc[0] = sqrt(S[0]);
    Lc[1] = S[1] / Lc[0];
    Lc[3] = sqrt(S[3] - Lc[1] * Lc[1]);
  } else if (n0) {
    Lc[0] = sqrt(S[0]);
  } else if (n1) {
    Lc[3] = sqrt(S[3]);
  }
}
template <int M>
PF_HD void pf_psd_chol(const double* S, double* Lc) {
  double mx = S[0];
  if (M == 2 && S[3] > mx) mx = S[3];
  double thr = PF_EPS * (double)M * mx;
  if (PF_EPS > thr) thr = PF_EPS;
  pf_chol_lower_nz<M>(S, thr, Lc);
}

/* dmvnorm(x, mean, S, lwr = false, logd = true): S a covariance matrix */
template <int M>
PF_HD double pf_dmvnorm_var(const double* x, const double* mean, const double* S) {
  if (M == 1) return -0.5 * (PF_LOG2PI + pf_log(S[0]) + (x[0] - mean[0]) * (x[0] - mean[0]) / S[0]);
  const double z0 = x[0] - mean[0], z1 = x[1] - mean[1];
  const bool n0 = S[0] > PF_EPS, n1 = S[3] > PF_EPS;
  if (n0 && n1) {
    const double l00 = sqrt(S[0]), l10 = S[1] / l00, l11 = sqrt(S[3] - l10 * l10);
    const double r00 = 1.0 / l00, r11 = 1.0 / l11, r10 = -(l10 * r00) / l11;
    const double a = r00 * z0, b = r10 * z0 + r11 * z1;
    return ((pf_log(r00) + pf_log(r11)) - 0.5 * 2.0 * PF_LOG2PI) - 0.5 * (a * a + b * b);
  }
  if (n0 || n1) {
    const double s = n0 ? S[0] : S[3], z = n0 ? z0 : z1;
    const double r = 1.0 / sqrt(s), a = r * z;
    return (pf_log(r) - 0.5 * 1.0 * PF_LOG2PI) - 0.5 * (a * a);
  }
  return 0.0;
}
/* dmvnorm(x, mean, Lc, lwr = true, logd = true): the covariance is Lc Lc' */
template <int M>
PF_HD double pf_dmvnorm_lwr(const double* x, const double* mean, const double* Lc) {
  if (M == 1) {
    const double q = (x[0] - mean[0]) / Lc[0];
    return -0.5 * (PF_LOG2PI + 2.0 * pf_log(Lc[0]) + q * q);
  }
  double S[4];
  S[0] = Lc[0] * Lc[0] + Lc[2] * Lc[2];
  S[1] = Lc[1] * Lc[0] + Lc[3] * Lc[2];
  S[2] = Lc[0] * Lc[1] + Lc[2] * Lc[3];
  S[3] = Lc[1] * Lc[1] + Lc[3] * Lc[3];
  return pf_dmvnorm_var<2>(x, mean, S);
}

/* ekf_update_step for one finite observation y: att (M), Ptt (M x M) from at, Pt with the
 * linearised observation row Zg (M), its value zf = Z_fn(at) and HH = H^2 */
template <int M>
PF_HD void pf_ekf_update(double y, const double* at, const double* Pt, const double* Zg, double zf, double HH,
                         double* att, double* Ptt) {
  double pz[M], K[M];
  double F = 0.0;
  {
    double zp[M];
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
      for (int l = 0; l < M; ++l) s += Zg[l] * Pt[l + M * j];
      zp[j] = s;
    }
    for (int j = 0; j < M; ++j) F += zp[j] * Zg[j];
  }
  F = F + HH;
  const double inv = 1.0 / sqrt(F);
  const double v = y - zf;
  for (int i = 0; i < M; ++i) {
    double s = 0.0;
    for (int l = 0; l < M; ++l) s += Pt[i + M * l] * Zg[l];
    pz[i] = s;
    K[i] = pz[i] * inv * inv;
    att[i] = at[i] + K[i] * v;
  }
  double tmp[M * M], A[M * M];
  for (int j = 0; j < M; ++j)
    for (int i = 0; i < M; ++i) tmp[i + M * j] = (i == j ? 1.0 : 0.0) - K[i] * Zg[j];
  for (int j = 0; j < M; ++j)
    for (int i = 0; i < M; ++i) {
      double s = 0.0;
      for (int l = 0; l < M; ++l) s += tmp[i + M * l] * Pt[l + M * j];
      A[i + M * j] = s;
    }
  for (int j = 0; j < M; ++j)
    for (int i = 0; i < M; ++i) {
      double s = 0.0;
      for (int l = 0; l < M; ++l) s += A[i + M * l] * tmp[j + M * l];
      Ptt[i + M * j] = s + K[i] * HH * K[j];
    }
}

#endif /* BSSM_B200_PF_EKPF_H */
