// approx_device.cuh -- approximate_for_is for ssm_ung models of any state dimension ON THE DEVICE,
// one thread per stored draw: what pfga::UngApprox::run(mode) (include/bssm_b200/ung_gauss_approx.h) does on
// the host -- laplace_iter at the stored mode (src/model_ssm_ung.cpp:302-342), update_scales (:241-292),
// smoother_ccov (src/model_ssm_ulg.cpp:419-473), conditional_cov by SVD (src/conditional_dist.cpp:8-32).
//
// Why a second placement: post_correct of a bsm_ng model (m = 2, 3, 5, 13) spent 95 % of its wall time in the
// host's OpenMP loop over the draws (4 000 draws, n = 500, 16 cores: 297 ms of approximations around 16 ms of
// particle filters).  The recursion is sequential in t, but the draws are independent: one thread per draw
// runs the whole chain with its small matrices in registers / local memory.  Every expression is the host
// header's, evaluated with the same primitives (pf_smallmat.h: products accumulate over the inner index, left
// to right; the SVD is the same cyclic Jacobi sweep), so the two placements give the same bits
// (tests: BSSM_B200_HOST_APPROX=1 forces the host loop).
#pragma once
#include "pf_device.cuh"
#include "bssm_b200/pf_smallmat.h"
#include "bssm_b200/ung_gauss_approx.h"

namespace pf {

struct ApproxParams {
  int n, m, k, distribution;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double *y, *u;
  // model arrays, per draw with stride fs_* (0 = shared)
  const double *Z, *T, *R, *a1, *P1, *D, *C, *xbeta;
  size_t fs_Z, fs_T, fs_R, fs_a1, fs_P1, fs_D, fs_C, fs_x;
  int has_xbeta;
  double phi;
  const double* phi_f; // per draw, or null
  const double* modes; // n per draw
  // outputs, per draw
  double *alphahat, *Lt, *Ct, *ytilde, *Htilde, *scales;
  size_t fs_a, fs_m, fs_n;
  double* scratch; // (m + 3) * n per draw: aHH, F, v, K
};

// FIXED: the state dimension IS MAXM (bsm_ng: 2 or 3) -- every loop over it unrolls and the small matrices live
// in registers; otherwise m <= MAXM at run time (seasonal models) and they live in local memory.
template <int MAXM, bool FIXED>
__global__ void __launch_bounds__(32) ung_approx_for_is_kernel(const ApproxParams p, uint32_t nf) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  constexpr int MM = MAXM * MAXM;
  const int n = p.n, m = FIXED ? MAXM : p.m, k = p.k, dist = p.distribution;
  const size_t mm = (size_t)m * m;
  const double zero_tol = 1e-12;
  const double* Zb = p.Z + (size_t)f * p.fs_Z;
  const double* Tb = p.T + (size_t)f * p.fs_T;
  const double* Rb = p.R + (size_t)f * p.fs_R;
  const double* Cb = p.C + (size_t)f * p.fs_C;
  const double* Db = p.D + (size_t)f * p.fs_D;
  const double* a1 = p.a1 + (size_t)f * p.fs_a1;
  const double* P1 = p.P1 + (size_t)f * p.fs_P1;
  const double* xb = p.has_xbeta ? p.xbeta + (size_t)f * p.fs_x : nullptr;
  auto Z = [&](int t) { return Zb + (size_t)m * (t * p.Ztv); };
  auto T = [&](int t) { return Tb + (size_t)m * m * (t * p.Ttv); };
  auto R = [&](int t) { return Rb + (size_t)m * k * (t * p.Rtv); };
  auto C = [&](int t) { return Cb + (size_t)m * (t * p.Ctv); };
  auto D = [&](int t) { return Db[t * p.Dtv]; };
  const double phi = p.phi_f ? p.phi_f[f] : p.phi;
  const double* mode = p.modes + (size_t)f * n;
  double* ah = p.alphahat + (size_t)f * p.fs_a;
  double* Pt = p.Lt + (size_t)f * p.fs_m;
  double* ccov = p.Ct + (size_t)f * p.fs_m;
  double* ay = p.ytilde + (size_t)f * p.fs_n;
  double* aH = p.Htilde + (size_t)f * p.fs_n;
  double* sc = p.scales + (size_t)f * p.fs_n;
  double* aHH = p.scratch + (size_t)f * (size_t)(m + 3) * n;
  double* sF = aHH + n;
  double* sv = sF + n;
  double* sK = sv + n;
  auto ytmp = [&](int t) { return xb ? ay[t] - xb[t] : ay[t]; };

  // ---- laplace_iter at the stored mode, update_scales
  for (int t = 0; t < n; ++t) {
    const double y = p.y[t], u = p.u[t], s = mode[t];
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
    aHH[t] = HH;
    ay[t] = yt;
    const double H = sqrt(HH);
    aH[t] = H;
    double scale = 0.0;
    if (pf_isfinite(y)) {
      double q = (yt - s) / H;
      scale = pfga::ung_log_obs(dist, s, y, u, phi) + 0.5 * (q * q);
    }
    sc[t] = scale;
  }

  // ---- forward pass of smoother_ccov (Kalman filter, Joseph form)
  double A[MM], B[MM], X[MM], E[MM], rr[MM];
  // the carried state (a_t, P_t) stays in registers: reading back what the previous iteration stored would put
  // an L2 round trip on every link of the chain
  double at[MAXM], Pc[MM];
  for (int i = 0; i < m; ++i) { at[i] = a1[i]; ah[i] = at[i]; }
  for (size_t e = 0; e < mm; ++e) { Pc[e] = P1[e]; Pt[e] = Pc[e]; }
  for (int t = 0; t < n; ++t) {
    double zp[MAXM], K[MAXM], tmpv[MAXM], ta[MAXM];
    pf_mtm(Z(t), Pc, zp, 1, m, m); /* Z' Pt */
    double fs = 0.0;
    for (int i = 0; i < m; ++i) fs += zp[i] * Z(t)[i];
    const double F = fs + aHH[t];
    const double yy = ytmp(t);
    const bool used = pf_isfinite(yy) && F > zero_tol;
    double v = 0.0;
    if (used) {
      pf_mm(Pc, Z(t), K, m, m, 1); /* Pt Z / F */
      for (int i = 0; i < m; ++i) K[i] = K[i] / F;
      double za = 0.0;
      for (int i = 0; i < m; ++i) za += Z(t)[i] * at[i];
      v = yy - D(t) - za;
    } else {
      for (int i = 0; i < m; ++i) K[i] = 0.0;
    }
    sF[t] = F; sv[t] = v;
    for (int i = 0; i < m; ++i) sK[(size_t)m * t + i] = K[i];
    // predict_mean
    for (int i = 0; i < m; ++i) tmpv[i] = used ? at[i] + K[i] * v : at[i];
    pf_mm(T(t), tmpv, ta, m, m, 1);
    double an[MAXM];
    for (int i = 0; i < m; ++i) { an[i] = C(t)[i] + ta[i]; ah[(size_t)m * (t + 1) + i] = an[i]; }
    // predict_cov_joseph
    if (used) {
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - K[i] * Z(t)[j];
      pf_mm(E, Pc, A, m, m, m);
      pf_mmt(A, E, B, m, m, m); /* tmp Pt tmp' */
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = B[i + (size_t)m * j] + K[i] * aHH[t] * K[j];
    } else {
      for (size_t e = 0; e < mm; ++e) X[e] = Pc[e];
    }
    pf_mm(T(t), X, A, m, m, m);
    pf_mmt(A, T(t), B, m, m, m);
    pf_mmt(R(t), R(t), rr, m, k, m); /* R R' */
    double Pn[MM];
    for (size_t e = 0; e < mm; ++e) Pn[e] = B[e] + rr[e];
    pf_symmatu(Pn, m);
    for (size_t e = 0; e < mm; ++e) { Pt[mm * (t + 1) + e] = Pn[e]; ccov[mm * t + e] = Pn[e]; Pc[e] = Pn[e]; }
    for (int i = 0; i < m; ++i) at[i] = an[i];
  }

  // ---- backward pass
  {
    double rt[MAXM], Nt[MM], Lm[MM], lr[MAXM], tmpv[MAXM];
    for (int i = 0; i < m; ++i) rt[i] = 0.0;
    for (size_t e = 0; e < mm; ++e) Nt[e] = 0.0;
    for (int t = n - 1; t >= 0; --t) {
      const double F = sF[t], v = sv[t];
      const double* K = sK + (size_t)m * t;
      const bool used = pf_isfinite(ytmp(t)) && F > zero_tol;
      const double* Tm;
      if (used) {
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - K[i] * Z(t)[j];
        pf_mm(T(t), E, Lm, m, m, m);
        Tm = Lm;
      } else {
        Tm = T(t);
      }
      double* Pc = Pt + mm * t;
      double* cct = ccov + mm * t;
      /* ccov_t = Pt_t * L' * (I - Nt * ccov_t) */
      pf_mm(Nt, cct, A, m, m, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - A[i + (size_t)m * j];
      pf_mmt(Pc, Tm, B, m, m, m);
      pf_mm(B, X, cct, m, m, m);
      /* rt, Nt */
      pf_mtm(Tm, rt, lr, m, m, 1);
      pf_mtm(Tm, Nt, A, m, m, m);
      pf_mm(A, Tm, B, m, m, m);
      if (used) {
        for (int i = 0; i < m; ++i) rt[i] = Z(t)[i] / F * v + lr[i];
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) Nt[i + (size_t)m * j] = Z(t)[i] * Z(t)[j] / F + B[i + (size_t)m * j];
      } else {
        for (int i = 0; i < m; ++i) rt[i] = lr[i];
        for (size_t e = 0; e < mm; ++e) Nt[e] = B[e];
      }
      pf_symmatu(Nt, m);
      pf_mm(Pc, rt, tmpv, m, m, 1);
      for (int i = 0; i < m; ++i) ah[(size_t)m * t + i] += tmpv[i];
      pf_mm(Pc, Nt, A, m, m, m);
      pf_mm(A, Pc, B, m, m, m);
      pf_symmatu(B, m);
      for (size_t e = 0; e < mm; ++e) Pc[e] -= B[e];
    }
    for (size_t e = 0; e < mm; ++e) ccov[mm * n + e] = 0.0;
  }

  // ---- conditional_cov, SVD branch (pfga::conditional_cov_svd)
  {
    const double eps = 2.220446049250313e-16;
    double vals[MAXM], U[MM], W[MM];
    auto factor = [&](double* V) { /* V <- U diag(sqrt(s)) */
      pf_jacobi_eig_sym_t<MM>(V, m, vals, U);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) V[i + (size_t)m * j] = U[i + (size_t)m * j] * sqrt(fabs(vals[j]));
    };
    for (int t = n; t > 0; --t) {
      pf_jacobi_eig_sym_t<MM>(Pt + mm * (t - 1), m, vals, U);
      const double s0 = fabs(vals[0]);
      pf_mtm(ccov + mm * (t - 1), U, W, m, m, m);
      for (int j = 0; j < m; ++j) {
        const double s = fabs(vals[j]);
        const double inv = (s > eps * m * s0) ? 1.0 / s : 0.0;
        for (int i = 0; i < m; ++i) W[i + (size_t)m * j] = W[i + (size_t)m * j] * inv;
      }
      pf_mmt(W, U, X, m, m, m);
      pf_mm(X, ccov + mm * (t - 1), A, m, m, m);
      for (size_t e = 0; e < mm; ++e) Pt[mm * t + e] -= A[e];
      for (size_t e = 0; e < mm; ++e) ccov[mm * t + e] = X[e];
      factor(Pt + mm * t);
    }
    factor(Pt);
  }
}

} // namespace pf
