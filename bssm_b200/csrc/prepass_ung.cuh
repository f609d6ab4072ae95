// prepass_ung.cuh -- psi-APF pre-pass for m = 1 ssm_ung models, on the device,
// one warp per filter (= per MCMC draw in post_correct).
//
// Per filter it does what the head of ssm_ung::psi_filter does
// (/root/reference/src/model_ssm_ung.cpp:514-533) once the mode is known:
//   svm::update_model            src/model_svm.cpp:12-29   (theta -> T,R,P1,phi|a1,C)
//   laplace_iter                 src/model_ssm_ung.cpp:302-342
//   update_scales                src/model_ssm_ung.cpp:241-292
//   approx_model.log_likelihood  src/model_ssm_ulg.cpp:106-138 (only when asked)
//   compute_const_term           src/model_ssm_ung.cpp:347-370
//   smoother_ccov                src/model_ssm_ulg.cpp:419-473
//   conditional_cov (SVD branch) src/conditional_dist.cpp:8-32, scalar case
// The element-wise parts are spread over the lanes; the Kalman recursions are
// inherently sequential in t and run on lane 0.
#pragma once
#include "pf_device.cuh"
#include "models_ung.cuh"

namespace pf {

struct PrepassParams {
  int n, distribution;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double *y, *Z, *T, *R, *a1, *P1, *D, *C, *xbeta, *u;
  int has_xbeta;
  double phi;
  // per-draw scalar model (post_correct): 6 per filter = T, R, a1, P1, C, phi,
  // i.e. what update_model(theta_i) sets (src/model_svm.cpp:12-29,
  // src/model_ar1_ng.cpp:12-36), evaluated for all draws up front like the
  // reference's parset_ung does (src/parset_ng.cpp:5-72).  null = use the model.
  const double* draw_par;
  // mode of the Gaussian approximation: n per filter.  `modes_lin` is the point
  // the pseudo-observations are linearised at: after ssm_ung::approximate() that
  // is the mode of the LAST BUT ONE Newton sweep (the reference does not
  // re-linearise after its final smoother pass, src/model_ssm_ung.cpp:115-137),
  // while update_scales (:241-292) is evaluated at the final mode; for
  // approximate_for_is (:143-161) both are the stored mode.
  const double* modes;
  const double* modes_lin;
  size_t mode_stride;
  int compute_loglik;  // 0: approximate_for_is (approx_loglik = 0)
  double dist_const;   // state-independent constants of the observation density
  double log2pi;
  // outputs, per filter stride pstride (>= n+1)
  double *alphahat, *Lt, *Ct, *ytilde, *Htilde, *scales, *approx_loglik, *phi_f;
  double* scratch;  // 3 * pstride per filter: vt, Ft, Kt
  size_t pstride;
};

__global__ void __launch_bounds__(128)
ung_prepass_m1_kernel(const PrepassParams pp, uint32_t n_filters) {
  const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= n_filters) return;
  const int n = pp.n;
  const size_t o = (size_t)f * pp.pstride;
  double* ah = pp.alphahat + o;
  double* Vt = pp.Lt + o;
  double* cc = pp.Ct + o;
  double* ay = pp.ytilde + o;
  double* aH = pp.Htilde + o;
  double* sc = pp.scales + o;
  double* vt = pp.scratch + 3 * o;
  double* Ft = vt + pp.pstride;
  double* Kt = Ft + pp.pstride;
  const double* mode = pp.modes + (size_t)f * pp.mode_stride;
  const double* mode_lin = pp.modes_lin + (size_t)f * pp.mode_stride;

  // ---- model of this filter
  const bool from_theta = pp.draw_par != nullptr;
  double phi = pp.phi, Tc = 0, Rc = 0, a1 = pp.a1[0], P1 = pp.P1[0], Cc = 0;
  if (from_theta) {
    const double* dp = pp.draw_par + (size_t)f * 6;
    Tc = dp[0]; Rc = dp[1]; a1 = dp[2]; P1 = dp[3]; Cc = dp[4]; phi = dp[5];
  }
  if (lane == 0) pp.phi_f[f] = phi;
  auto Zt = [&](int t) { return pp.Z[t * pp.Ztv]; };
  auto Tt = [&](int t) { return from_theta ? Tc : pp.T[t * pp.Ttv]; };
  auto RRt = [&](int t) { double r = from_theta ? Rc : pp.R[t * pp.Rtv]; return r * r; };
  auto Ct = [&](int t) { return from_theta ? Cc : pp.C[t * pp.Ctv]; };
  auto Dt = [&](int t) { return pp.D[t * pp.Dtv]; };

  // ---- laplace_iter (:302-342) and update_scales (:241-292), lanes over t
  const int dist = pp.distribution;
  for (int t = lane; t < n; t += 32) {
    const double y = pp.y[t], u = pp.u[t];
    double s = mode_lin[t];
    double HH, yt;
    switch (dist) {
      case 0: {
        double tmp = y;
        if (fabs(tmp) < 1e-4) tmp = 1e-4;
        double q = tmp / phi;
        HH = 2.0 * pf_exp(s) / (q * q);
        yt = s + 1.0 - 0.5 * HH;
      } break;
      case 1: {
        HH = 1.0 / (pf_exp(s) * u);
        yt = y * HH + s - 1.0;
      } break;
      case 2: {
        double e = pf_exp(s);
        HH = (1.0 + e) * (1.0 + e) / (u * e);
        yt = y * HH + s - 1.0 - e;
      } break;
      case 3: {
        double e = pf_exp(s) * u;
        HH = (phi + e) * (phi + e) / (phi * e * (y + phi));
        yt = s + (phi + e) * (y - e) / ((y + phi) * e);
      } break;
      default: {
        double e = pf_exp(s) * u;
        HH = e / (y * phi);
        yt = s - e / y + 1;
      } break;
    }
    const double H = sqrt(HH);
    ay[t] = yt;
    aH[t] = H;
    Ft[t] = HH; // parked here until the forward pass overwrites it
    double scale = 0.0;
    s = mode[t];
    if (pf_isfinite(y)) {
      double qy = y / phi;
      double q = (yt - s) / H;
      scale = ung_log_obs(dist, s, y, u, phi, qy * qy) + 0.5 * (q * q);
    }
    sc[t] = scale;
  }
  __syncwarp();
  if (lane != 0) return;

  const double zero_tol = 1e-12; // src/model_ssm_ulg.h:19
  auto ytmp = [&](int t) { return pp.has_xbeta ? ay[t] - pp.xbeta[t] : ay[t]; };

  // ---- approx_loglik (:519-528)
  double approx_loglik = 0.0;
  if (pp.compute_loglik) {
    double sumH = 0.0, sumR = 0.0;
    for (int t = 0; t < n; ++t) sumH += aH[t];
    if (from_theta) sumR = Rc;
    else { int nR = pp.Rtv ? n : 1; for (int t = 0; t < nR; ++t) sumR += pp.R[t]; }
    double gl = 0.0;
    if (sumH + sumR < zero_tol) {
      gl = -pf_inf();
    } else {
      double at = a1, Pt = P1;
      for (int t = 0; t < n; ++t) {
        const double HH = Ft[t];
        double F = Zt(t) * Pt * Zt(t) + HH;
        double yt = ytmp(t);
        if (pf_isfinite(yt) && F > zero_tol) {
          double v = yt - Dt(t) - Zt(t) * at;
          double K = Pt * Zt(t) / F;
          at = Ct(t) + Tt(t) * (at + K * v);
          Pt = Tt(t) * (Pt - K * K * F) * Tt(t) + RRt(t);
          gl -= 0.5 * (pp.log2pi + pf_log(F) + v * v / F);
        } else {
          at = Ct(t) + Tt(t) * at;
          Pt = Tt(t) * Pt * Tt(t) + RRt(t);
        }
      }
    }
    int nobs = 0;
    double sumlog = 0.0;
    for (int t = 0; t < n; ++t)
      if (pf_isfinite(pp.y[t])) { nobs++; sumlog += pf_log(aH[t]); }
    double norm = -0.5 * nobs * pp.log2pi - sumlog;
    double ct = pp.dist_const - norm;
    // arma::accu(scales): two accumulators [UNVENDORED]
    double acc1 = 0.0, acc2 = 0.0;
    int j;
    for (j = 1; j < n; j += 2) { acc1 += sc[j - 1]; acc2 += sc[j]; }
    if ((j - 1) < n) acc1 += sc[j - 1];
    approx_loglik = gl + ct + (acc1 + acc2);
  }
  pp.approx_loglik[f] = approx_loglik;

  // ---- smoother_ccov (src/model_ssm_ulg.cpp:419-473); Vt holds Pt
  ah[0] = a1;
  Vt[0] = P1;
  for (int t = 0; t < n; ++t) {
    const double HH = Ft[t];
    const double Pt = Vt[t];
    const double F = Zt(t) * Pt * Zt(t) + HH;
    Ft[t] = F;
    const double yt = ytmp(t);
    if (pf_isfinite(yt) && F > zero_tol) {
      const double K = Pt * Zt(t) / F;
      Kt[t] = K;
      const double v = yt - Dt(t) - Zt(t) * ah[t];
      vt[t] = v;
      ah[t + 1] = Ct(t) + Tt(t) * (ah[t] + K * v);
      const double tmp = 1.0 - K * Zt(t);
      Vt[t + 1] = Tt(t) * (tmp * Pt * tmp + K * HH * K) * Tt(t) + RRt(t);
    } else {
      ah[t + 1] = Ct(t) + Tt(t) * ah[t];
      Vt[t + 1] = Tt(t) * Pt * Tt(t) + RRt(t);
    }
    cc[t] = Vt[t + 1];
  }
  double rt = 0.0, Nt = 0.0;
  for (int t = n - 1; t >= 0; --t) {
    if (pf_isfinite(ytmp(t)) && Ft[t] > zero_tol) {
      const double L = Tt(t) * (1.0 - Kt[t] * Zt(t));
      cc[t] = Vt[t] * L * (1.0 - Nt * cc[t]);
      rt = Zt(t) / Ft[t] * vt[t] + L * rt;
      Nt = Zt(t) * Zt(t) / Ft[t] + L * Nt * L;
    } else {
      cc[t] = Vt[t] * Tt(t) * (1.0 - Nt * cc[t]);
      rt = Tt(t) * rt;
      Nt = Tt(t) * Nt * Tt(t);
    }
    ah[t] += Vt[t] * rt;
    Vt[t] -= Vt[t] * Nt * Vt[t];
  }
  cc[n] = 0.0;

  // ---- conditional_cov, scalar SVD branch (src/conditional_dist.cpp:8-32)
  for (int t = n; t > 0; --t) {
    const double s = fabs(Vt[t - 1]);
    const double tmp = (s > 2.220446049250313e-16 * s) ? cc[t - 1] * (1.0 / s) : 0.0;
    Vt[t] -= tmp * cc[t - 1];
    cc[t] = tmp;
    Vt[t] = sqrt(fabs(Vt[t]));
  }
  Vt[0] = sqrt(fabs(Vt[0]));
}

} // namespace pf
