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
// inherently sequential in t: all lanes advance the same scalar state, with the
// per-step inputs loaded coalesced per tile of 32 and broadcast by shuffles.
#pragma once
#include "pf_device.cuh"
#include "models_ung.cuh"

namespace pf {

struct PrepassParams {
  int n, distribution;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double *y, *Z, *T, *R, *a1, *P1, *D, *C, *xbeta, *u;
  int has_xbeta;
  size_t xbeta_stride; // 0: one regression part for all filters; n: one per filter (beta is part of theta)
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
  // per-step scalars every thread of the filter kernel would otherwise divide for: 1 / Htilde_t and (y_t / phi)^2
  double *rHt, *yphi2;
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
    pp.rHt[o + t] = 1.0 / H;
    { const double qy = y / phi; pp.yphi2[o + t] = qy * qy; }
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

  // The Kalman recursions below are sequential in t.  They run REDUNDANTLY on all 32 lanes:
  // a tile of 32 time points is loaded coalesced (one element per lane), each step broadcasts
  // its inputs with a shuffle and every lane advances the same scalar state, lane k keeping the
  // outputs of step k for a coalesced store after the tile.  The dependent chain of a step is
  // then arithmetic + shuffles only (no L2 round trip per step): 1.9 ms -> 0.6 ms for n = 1000.
  // Every expression is evaluated exactly as in the one-lane formulation (same operands, same order).
  auto bc = [](double v, int k) { return __shfl_sync(0xffffffffu, v, k); };
  const double zero_tol = 1e-12; // src/model_ssm_ulg.h:19
  const double* xbeta = pp.xbeta + (size_t)f * pp.xbeta_stride;
  auto ytmp = [&](int t) { return pp.has_xbeta ? ay[t] - xbeta[t] : ay[t]; };

  // coefficients that do not vary in time are read once instead of being broadcast every step
  const bool tvZ = pp.Ztv != 0, tvT = !from_theta && pp.Ttv != 0, tvR = !from_theta && pp.Rtv != 0;
  const bool tvC = !from_theta && pp.Ctv != 0, tvD = pp.Dtv != 0;
  const double z0 = Zt(0), T0 = Tt(0), RR0 = RRt(0), C0 = Ct(0), D0 = Dt(0);

  // ---- approx_loglik (:519-528) and the forward pass of smoother_ccov
  // (src/model_ssm_ulg.cpp:419-473; Vt holds Pt) in ONE sweep over t: the two Kalman recursions
  // (different covariance updates, hence different roundings) are independent dependency chains
  // and overlap in the pipeline.
  double approx_loglik = 0.0;
  {
    const bool ll = pp.compute_loglik != 0;
    double sumH = 0.0, sumR = 0.0;
    if (ll) {
      for (int base = 0; base < n; base += 32) {
        const int tl = base + lane;
        const double h_l = tl < n ? aH[tl] : 0.0;
        const int cnt = n - base < 32 ? n - base : 32;
        for (int k = 0; k < cnt; ++k) sumH += bc(h_l, k);
      }
      if (from_theta) sumR = Rc;
      else { int nR = pp.Rtv ? n : 1; for (int t = 0; t < nR; ++t) sumR += pp.R[t]; }
    }
    const bool run_ll = ll && !(sumH + sumR < zero_tol);
    double gl = (ll && !run_ll) ? -pf_inf() : 0.0;
    double at = a1, Pl = P1;   // log-likelihood recursion
    double aht = a1, Pt = P1;  // smoother forward recursion
    int nobs = 0;
    double sumlog = 0.0, acc1 = 0.0, acc2 = 0.0; // arma::accu(scales): two accumulators [UNVENDORED]
    if (lane == 0) { ah[0] = a1; Vt[0] = P1; }
    for (int base = 0; base < n; base += 32) {
      const int tl = base + lane;
      const bool in = tl < n;
      const double z_l = (in && tvZ) ? Zt(tl) : 0.0, T_l = (in && tvT) ? Tt(tl) : 0.0, RR_l = (in && tvR) ? RRt(tl) : 0.0;
      const double C_l = (in && tvC) ? Ct(tl) : 0.0, D_l = (in && tvD) ? Dt(tl) : 0.0, HH_l = in ? Ft[tl] : 0.0;
      const double yt_l = in ? ytmp(tl) : 0.0, sc_l = (in && ll) ? sc[tl] : 0.0;
      const bool obs_l = in && ll && pf_isfinite(pp.y[tl]);
      const double lg_l = obs_l ? pf_log(aH[tl]) : 0.0;
      const unsigned obs_mask = __ballot_sync(0xffffffffu, obs_l);
      double F_o = 0.0, K_o = 0.0, v_o = 0.0, ahn_o = 0.0, Vn_o = 0.0;
      const int cnt = n - base < 32 ? n - base : 32;
      for (int k = 0; k < cnt; ++k) {
        const double z = tvZ ? bc(z_l, k) : z0, Tk = tvT ? bc(T_l, k) : T0, RRk = tvR ? bc(RR_l, k) : RR0;
        const double Ck = tvC ? bc(C_l, k) : C0, Dk = tvD ? bc(D_l, k) : D0;
        const double HH = bc(HH_l, k), yt = bc(yt_l, k);
        const bool yfin = pf_isfinite(yt);
        if (run_ll) {
          double F = z * Pl * z + HH;
          if (yfin && F > zero_tol) {
            double v = yt - Dk - z * at;
            double K = Pl * z / F;
            at = Ck + Tk * (at + K * v);
            Pl = Tk * (Pl - K * K * F) * Tk + RRk;
            gl -= 0.5 * (pp.log2pi + pf_log(F) + v * v / F);
          } else {
            at = Ck + Tk * at;
            Pl = Tk * Pl * Tk + RRk;
          }
        }
        if (ll) {
          const double lgk = bc(lg_l, k), sck = bc(sc_l, k);
          if ((obs_mask >> k) & 1u) { nobs++; sumlog += lgk; }
          if (k & 1) acc2 += sck; else acc1 += sck; // base is a multiple of 32: parity of t = parity of k
        }
        const double F = z * Pt * z + HH;
        double Kk = 0.0, vk = 0.0, ahn, Vn;
        if (yfin && F > zero_tol) {
          Kk = Pt * z / F;
          vk = yt - Dk - z * aht;
          ahn = Ck + Tk * (aht + Kk * vk);
          const double tmp = 1.0 - Kk * z;
          Vn = Tk * (tmp * Pt * tmp + Kk * HH * Kk) * Tk + RRk;
        } else {
          ahn = Ck + Tk * aht;
          Vn = Tk * Pt * Tk + RRk;
        }
        if (lane == k) { F_o = F; K_o = Kk; v_o = vk; ahn_o = ahn; Vn_o = Vn; }
        aht = ahn; Pt = Vn;
      }
      if (in) { Ft[tl] = F_o; Kt[tl] = K_o; vt[tl] = v_o; ah[tl + 1] = ahn_o; Vt[tl + 1] = Vn_o; cc[tl] = Vn_o; }
    }
    if (ll) {
      double norm = -0.5 * nobs * pp.log2pi - sumlog;
      double ct = pp.dist_const - norm;
      approx_loglik = gl + ct + (acc1 + acc2);
    }
  }
  if (lane == 0) pp.approx_loglik[f] = approx_loglik;
  __syncwarp();
  // backward pass: everything that does not depend on (r_t, N_t) is computed per lane, outside the chain
  {
    double rt = 0.0, Nt = 0.0;
    for (int base = ((n - 1) / 32) * 32; base >= 0; base -= 32) {
      const int tl = base + lane;
      const bool in = tl < n;
      const double z = in ? Zt(tl) : 0.0, Tk_l = in ? Tt(tl) : 0.0;
      const double F = in ? Ft[tl] : 0.0, Kk = in ? Kt[tl] : 0.0, vk = in ? vt[tl] : 0.0;
      const bool used_l = in && pf_isfinite(ytmp(tl)) && F > zero_tol;
      const double L_l = used_l ? Tk_l * (1.0 - Kk * z) : Tk_l; // the multiplier of r_t and N_t
      const double zFv_l = used_l ? z / F * vk : 0.0;
      const double zzF_l = used_l ? z * z / F : 0.0;
      const double V_l = in ? Vt[tl] : 0.0, cc_l = in ? cc[tl] : 0.0, ah_l = in ? ah[tl] : 0.0;
      const unsigned used_mask = __ballot_sync(0xffffffffu, used_l);
      double cc_o = 0.0, ah_o = 0.0, V_o = 0.0;
      const int cnt = n - base < 32 ? n - base : 32;
      for (int k = cnt - 1; k >= 0; --k) {
        const double L = bc(L_l, k), V = bc(V_l, k), ck = bc(cc_l, k), ahk = bc(ah_l, k);
        const double ccn = V * L * (1.0 - Nt * ck);
        if ((used_mask >> k) & 1u) {
          rt = bc(zFv_l, k) + L * rt;
          Nt = bc(zzF_l, k) + L * Nt * L;
        } else {
          rt = L * rt;
          Nt = L * Nt * L;
        }
        const double ahn = ahk + V * rt;
        const double Vn = V - V * Nt * V;
        if (lane == k) { cc_o = ccn; ah_o = ahn; V_o = Vn; }
      }
      if (in) { cc[tl] = cc_o; ah[tl] = ah_o; Vt[tl] = V_o; }
    }
    if (lane == 0) cc[n] = 0.0;
  }
  __syncwarp();

  // ---- conditional_cov, scalar SVD branch (src/conditional_dist.cpp:8-32).  Iteration t reads
  // Vt[t-1], cc[t-1] (not yet rewritten: t runs downwards) and rewrites Vt[t], cc[t]: no carried
  // state, so the lanes take one t each; tiles run downwards, reads of a tile before its writes.
  for (int base = ((n - 1) / 32) * 32; base >= 0; base -= 32) {
    const int t = base + lane + 1; // 1..n
    const bool in = t <= n;
    double Vnew = 0.0, cnew = 0.0;
    if (in) {
      const double s = fabs(Vt[t - 1]);
      const double tmp = (s > 2.220446049250313e-16 * s) ? cc[t - 1] * (1.0 / s) : 0.0;
      Vnew = Vt[t] - tmp * cc[t - 1];
      cnew = tmp;
      Vnew = sqrt(fabs(Vnew));
    }
    __syncwarp();
    if (in) { Vt[t] = Vnew; cc[t] = cnew; }
    __syncwarp();
  }
  if (lane == 0) Vt[0] = sqrt(fabs(Vt[0]));
}

} // namespace pf
