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
#include "sim_smoother.cuh"

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

// smoother_ccov + conditional_cov of one approximating linear-Gaussian model with scalar observations
// (pfga::UngApprox::smoother_and_conditional): pseudo-observations ay with variances aHH, model arrays as pointers
// with time-variation flags.  Outputs ah (m x (n+1)), Pt -> L (m x m x (n+1)), ccov -> C (m x m x (n+1)).
struct SmootherIO {
  int n, m, k;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double *Zb, *Tb, *Rb, *Cb, *Db, *a1, *P1, *xb; // xb: regression part to subtract from ay, or null
  const double *ay, *aHH;
  double *ah, *Pt, *ccov, *sF, *sv, *sK;
};
template <int MAXM, bool FIXED>
__device__ void smoother_ccov_conditional(const SmootherIO& io) {
  constexpr int MM = MAXM * MAXM;
  const int n = io.n, m = FIXED ? MAXM : io.m, k = io.k;
  const size_t mm = (size_t)m * m;
  const double zero_tol = 1e-12;
  auto Z = [&](int t) { return io.Zb + (size_t)m * (t * io.Ztv); };
  auto T = [&](int t) { return io.Tb + (size_t)m * m * (t * io.Ttv); };
  auto R = [&](int t) { return io.Rb + (size_t)m * k * (t * io.Rtv); };
  auto C = [&](int t) { return io.Cb + (size_t)m * (t * io.Ctv); };
  auto D = [&](int t) { return io.Db[t * io.Dtv]; };
  const double *a1 = io.a1, *P1 = io.P1, *xb = io.xb, *ay = io.ay, *aHH = io.aHH;
  double *ah = io.ah, *Pt = io.Pt, *ccov = io.ccov, *sF = io.sF, *sv = io.sv, *sK = io.sK;
  auto ytmp = [&](int t) { return xb ? ay[t] - xb[t] : ay[t]; };
#ifdef PF_APPROX_TIMING
  const bool tim = blockIdx.x == 0 && threadIdx.x == 0;
  long long tk0 = clock64();
#endif

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

#ifdef PF_APPROX_TIMING
  long long tk1 = clock64();
#endif
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

#ifdef PF_APPROX_TIMING
  long long tk2 = clock64();
#endif
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
#ifdef PF_APPROX_TIMING
  if (tim) printf("[approx cycles per t] forward %lld backward %lld conditional %lld\n", (tk1 - tk0) / n, (tk2 - tk1) / n, (clock64() - tk2) / n);
#endif
}

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

  SmootherIO io;
  io.n = n; io.m = m; io.k = k;
  io.Ztv = p.Ztv; io.Ttv = p.Ttv; io.Rtv = p.Rtv; io.Dtv = p.Dtv; io.Ctv = p.Ctv;
  io.Zb = Zb; io.Tb = Tb; io.Rb = Rb; io.Cb = Cb; io.Db = Db; io.a1 = a1; io.P1 = P1; io.xb = xb;
  io.ay = ay; io.aHH = aHH;
  io.ah = ah; io.Pt = Pt; io.ccov = ccov; io.sF = sF; io.sv = sv; io.sK = sK;
  smoother_ccov_conditional<MAXM, FIXED>(io);
}

// ------------------------------------------------------------------ SPDK
// approximate_for_is (laplace_iter at the stored mode + update_scales) and fast_precomputing_smoother
// (src/model_ssm_ulg.cpp:359-413) of one stored draw per thread, for is_correction_spdk: what
// pfga::UngApprox::run_head(mode) + fast_smoother(at, &gains) do on the host, written into the draw's block of
// the simulation smoother (sim_smoother.cuh) -- pseudo-observations, 1 / sd, (y / phi)^2, F_t, K_t,
// L_t = T_t (I - K_t Z_t'), R_t R_t', the smoothed mean -- plus accu(scales).  Same primitives, same order:
// same bits as the host header (the oracle's independent restatement is the reference of the parity tests).
struct SpdkGainParams {
  int n, m, k, distribution;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  SimBlock L;
  double* blocks;      // nf blocks of L.total doubles; the model part is filled in
  const double* modes; // n per draw
  double* ssum;        // nf: accu(scales)
  double* scratch;     // (1 + 1 + m) * n per draw: v, scales, r
};

template <int MAXM, bool FIXED>
__global__ void __launch_bounds__(32) spdk_gains_kernel(const SpdkGainParams p, uint32_t nf) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  constexpr int MM = MAXM * MAXM;
  const int n = p.n, m = FIXED ? MAXM : p.m, k = p.k, dist = p.distribution;
  const size_t mm = (size_t)m * m;
  const double zero_tol = 1e-12;
  double* b = p.blocks + (size_t)f * p.L.total;
  const double *y = b + p.L.y, *u = b + p.L.u, *xb = p.L.has_xbeta ? b + p.L.xbeta : nullptr;
  const double *Zb = b + p.L.Z, *Tb = b + p.L.T, *Rb = b + p.L.R, *Cb = b + p.L.C, *Db = b + p.L.D;
  const double *a1 = b + p.L.a1, *P1 = b + p.L.P1;
  const double phi = b[p.L.phi];
  double *ay = b + p.L.ay, *aH = b + p.L.aH, *rH = b + p.L.rH, *yp2 = b + p.L.yphi2;
  double *sF = b + p.L.F, *sK = b + p.L.K, *sL = b + p.L.Lt, *sRR = b + p.L.RR, *ah = b + p.L.ahat;
  const double* mode = p.modes + (size_t)f * n;
  double* sv = p.scratch + (size_t)f * (size_t)(m + 2) * n;
  double* sc = sv + n;
  double* rt = sc + n;
  auto Z = [&](int t) { return Zb + (size_t)m * (t * p.Ztv); };
  auto T = [&](int t) { return Tb + (size_t)m * m * (t * p.Ttv); };
  auto R = [&](int t) { return Rb + (size_t)m * k * (t * p.Rtv); };
  auto C = [&](int t) { return Cb + (size_t)m * (t * p.Ctv); };
  auto D = [&](int t) { return Db[t * p.Dtv]; };
  auto ytmp = [&](int t) { return xb ? ay[t] - xb[t] : ay[t]; };

  // ---- laplace_iter at the stored mode, update_scales (as in ung_approx_for_is_kernel)
  for (int t = 0; t < n; ++t) {
    const double yt_ = y[t], ut = u[t], s = mode[t];
    double HH, yt;
    switch (dist) {
      case 0: {
        double tmp = yt_;
        if (fabs(tmp) < 1e-4) tmp = 1e-4;
        double q = tmp / phi;
        HH = 2.0 * pf_exp(s) / (q * q);
        yt = s + 1.0 - 0.5 * HH;
      } break;
      case 1: {
        HH = 1.0 / (pf_exp(s) * ut);
        yt = yt_ * HH + s - 1.0;
      } break;
      case 2: {
        double e = pf_exp(s);
        HH = (1.0 + e) * (1.0 + e) / (ut * e);
        yt = yt_ * HH + s - 1.0 - e;
      } break;
      case 3: {
        double e = pf_exp(s) * ut;
        HH = (phi + e) * (phi + e) / (phi * e * (yt_ + phi));
        yt = s + (phi + e) * (yt_ - e) / ((yt_ + phi) * e);
      } break;
      default: {
        double e = pf_exp(s) * ut;
        HH = e / (yt_ * phi);
        yt = s - e / yt_ + 1;
      } break;
    }
    const double H = sqrt(HH);
    ay[t] = yt; aH[t] = H; rH[t] = 1.0 / H;
    yp2[t] = HH; // the variance is parked here until the forward pass has used it
    double scale = 0.0;
    if (pf_isfinite(yt_)) {
      double q = (yt - s) / H;
      scale = pfga::ung_log_obs(dist, s, yt_, ut, phi) + 0.5 * (q * q);
    }
    sc[t] = scale;
  }
  { // arma::accu, two accumulators (pfga::accu2)
    double s1 = 0.0, s2 = 0.0;
    int j;
    for (j = 1; j < n; j += 2) { s1 += sc[j - 1]; s2 += sc[j]; }
    if ((j - 1) < n) s1 += sc[j - 1];
    p.ssum[f] = s1 + s2;
  }

  // ---- forward pass: Kalman filter with the Joseph form (kalman_gain, predict_mean, predict_cov_joseph)
  double A[MM], B[MM], X[MM], E[MM], rr[MM], Pc[MM], at[MAXM];
  for (int i = 0; i < m; ++i) { at[i] = a1[i]; ah[i] = at[i]; }
  for (size_t e = 0; e < mm; ++e) Pc[e] = P1[e];
  for (int t = 0; t < n; ++t) {
    double zp[MAXM], K[MAXM], tmpv[MAXM], ta[MAXM];
    const double HH = yp2[t];
    pf_mtm(Z(t), Pc, zp, 1, m, m);
    double fs = 0.0;
    for (int i = 0; i < m; ++i) fs += zp[i] * Z(t)[i];
    const double F = fs + HH;
    const double yy = ytmp(t);
    const bool used = pf_isfinite(yy) && F > zero_tol;
    double v = 0.0;
    if (used) {
      pf_mm(Pc, Z(t), K, m, m, 1);
      for (int i = 0; i < m; ++i) K[i] = K[i] / F;
      double za = 0.0;
      for (int i = 0; i < m; ++i) za += Z(t)[i] * at[i];
      v = yy - D(t) - za;
    } else {
      for (int i = 0; i < m; ++i) K[i] = 0.0;
    }
    sF[t] = F; sv[t] = v;
    for (int i = 0; i < m; ++i) sK[(size_t)m * t + i] = K[i];
    for (int i = 0; i < m; ++i) tmpv[i] = used ? at[i] + K[i] * v : at[i];
    pf_mm(T(t), tmpv, ta, m, m, 1);
    for (int i = 0; i < m; ++i) { at[i] = C(t)[i] + ta[i]; ah[(size_t)m * (t + 1) + i] = at[i]; }
    if (used) {
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - K[i] * Z(t)[j];
      pf_mm(T(t), E, &sL[mm * t], m, m, m); // L_t = T_t (I - K_t Z_t'), kept for the simulated series
      pf_mm(E, Pc, A, m, m, m);
      pf_mmt(A, E, B, m, m, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = B[i + (size_t)m * j] + K[i] * HH * K[j];
    } else {
      for (size_t e = 0; e < mm; ++e) { X[e] = Pc[e]; sL[mm * t + e] = 0.0; }
    }
    pf_mm(T(t), X, A, m, m, m);
    pf_mmt(A, T(t), B, m, m, m);
    pf_mmt(R(t), R(t), rr, m, k, m);
    if (p.Rtv || t == 0)
      for (size_t e = 0; e < mm; ++e) sRR[mm * (p.Rtv ? t : 0) + e] = rr[e];
    for (size_t e = 0; e < mm; ++e) Pc[e] = B[e] + rr[e];
    pf_symmatu(Pc, m);
    const double q = y[t] / phi; // (y_t / phi)^2 of the weights takes the parking place of the variance
    yp2[t] = q * q;
  }

  // ---- backward pass for r_t and the smoothed mean (fast_smoother, src/model_ssm_ulg.cpp:335-354)
  double lr[MAXM], w[MAXM], pw[MAXM];
  for (int i = 0; i < m; ++i) rt[(size_t)m * (n - 1) + i] = 0.0;
  for (int t = n - 1; t > 0; --t) {
    const bool used = pf_isfinite(ytmp(t)) && sF[t] > zero_tol;
    if (used) {
      pf_mtm(&sL[mm * t], &rt[(size_t)m * t], lr, m, m, 1);
      for (int i = 0; i < m; ++i) rt[(size_t)m * (t - 1) + i] = Z(t)[i] / sF[t] * sv[t] + lr[i];
    } else {
      pf_mtm(T(t), &rt[(size_t)m * t], &rt[(size_t)m * (t - 1)], m, m, 1);
    }
  }
  if (pf_isfinite(ay[0]) && sF[0] > zero_tol) {
    pf_mtm(sL, rt, lr, m, m, 1);
    for (int i = 0; i < m; ++i) w[i] = Z(0)[i] / sF[0] * sv[0] + lr[i];
    pf_mm(P1, w, pw, m, m, 1);
  } else if (m == 1) {
    pw[0] = P1[0] * T(0)[0] * rt[0];
  } else {
    pf_mtm(T(0), rt, w, m, m, 1);
    pf_mm(P1, w, pw, m, m, 1);
  }
  for (int i = 0; i < m; ++i) { at[i] = a1[i] + pw[i]; ah[i] = at[i]; }
  for (int t = 0; t < n - 1; ++t) {
    double ta[MAXM], rrr[MAXM];
    pf_mm(T(t), at, ta, m, m, 1);
    pf_mm(&sRR[mm * (p.Rtv ? t : 0)], &rt[(size_t)m * t], rrr, m, m, 1);
    for (int i = 0; i < m; ++i) { at[i] = C(t)[i] + ta[i] + rrr[i]; ah[(size_t)m * (t + 1) + i] = at[i]; }
  }
}

} // namespace pf

// ------------------------------------------------------------------ ssm_nlg
// approximate_for_is of an ssm_nlg registry model (src/model_ssm_nlg.cpp:153-186) per stored draw on the device:
// relinearise Z_fn / T_fn at the stored mode (pfga::NlgApprox::linearise), smooth the linearised model (the
// function above, everything time varying), factor the transition covariances of the true and the approximating
// model (pfga::MvnFactor) and pack the 32 doubles per time point the psi-APF weights read (NlgPsiParams::stepc).
#include "bssm_b200/nlg_gauss_approx.h"
namespace pf {

struct NlgApproxParams {
  int n, model_id, n_par;
  const double* y;      // n
  const double* theta;  // n_par per draw
  const double* known;  // known_params
  const double* modes;  // m x n per draw
  double *stepc, *alphahat, *Lt, *Ct; // per draw: 32 n, m (n+1), m m (n+1) twice
  double* scratch;      // per draw: (2 m + m m + m k + 3 + m + 2) n
};

// log N factors of dmvnorm (pfga::MvnFactor::set) with fixed storage
template <int M_>
struct MvnFactorDev {
  double rooti[M_ * M_], cst, var1;
  __device__ void set(const double* S) {
    constexpr int m = M_;
    var1 = 0.0;
    if (m == 1) { var1 = S[0]; rooti[0] = 0.0; cst = 1.8378770664093453 + pf_log(var1); return; }
    double Lc[M_ * M_];
    for (int e = 0; e < m * m; ++e) { Lc[e] = 0.0; rooti[e] = 0.0; }
    for (int j = 0; j < m; ++j)
      for (int i = j; i < m; ++i) {
        double s = S[i + m * j];
        for (int l = 0; l < j; ++l) s -= Lc[i + m * l] * Lc[j + m * l];
        Lc[i + m * j] = (i == j) ? sqrt(s) : s / Lc[j + m * j];
      }
    for (int j = 0; j < m; ++j) {
      rooti[j + m * j] = 1.0 / Lc[j + m * j];
      for (int i = j + 1; i < m; ++i) {
        double s = 0.0;
        for (int l = j; l < i; ++l) s += Lc[i + m * l] * rooti[l + m * j];
        rooti[i + m * j] = -s / Lc[i + m * i];
      }
    }
    double sl = 0.0;
    for (int i = 0; i < m; ++i) sl += pf_log(rooti[i + m * i]);
    cst = sl - 0.5 * m * 1.8378770664093453;
  }
};

template <int M_, int K_>
__global__ void __launch_bounds__(32) nlg_approx_for_is_kernel(const NlgApproxParams p, uint32_t nf) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  constexpr int m = M_, k = K_, mm = M_ * M_, mk = M_ * K_;
  const int n = p.n;
  const size_t n1 = (size_t)n + 1;
  pfga::NlgFns fn{p.model_id, p.theta + (size_t)f * p.n_par, p.known};
  const double* mode = p.modes + (size_t)f * m * n;
  double* sc = p.scratch + (size_t)f * (size_t)(2 * m + mm + mk + 3 + m + 2) * n;
  double* Zv = sc;                 // m n
  double* Tv = Zv + (size_t)m * n; // mm n
  double* Rv = Tv + (size_t)mm * n; // mk n
  double* Cv = Rv + (size_t)mk * n; // m n
  double* Dv = Cv + (size_t)m * n;  // n
  double* Hv = Dv + n;              // n
  double* HHv = Hv + n;             // n
  double* sF = HHv + n;             // n
  double* sv = sF + n;              // n
  double* sK = sv + n;              // m n
  double a1[M_], P1[M_ * M_];
  fn.a1_fn(a1);
  fn.P1_fn(P1);
  // ---- linearise at the stored mode (zs = ts = mode)
  for (int t = 0; t < n; ++t) {
    const double* az = mode + (size_t)m * t;
    double tv[M_], tmpv[M_];
    fn.Z_gn(t, az, Zv + (size_t)m * t);
    Hv[t] = fn.H_fn(t, az);
    fn.T_gn(t, az, Tv + (size_t)mm * t);
    fn.R_fn(t, az, Rv + (size_t)mk * t);
    double za = 0.0;
    for (int d = 0; d < m; ++d) za += Zv[d + (size_t)m * t] * az[d];
    Dv[t] = fn.Z_fn(t, az) - za;
    fn.T_fn(t, az, tv);
    pf_mm(Tv + (size_t)mm * t, az, tmpv, m, m, 1);
    for (int d = 0; d < m; ++d) Cv[d + (size_t)m * t] = tv[d] - tmpv[d];
    HHv[t] = Hv[t] * Hv[t];
  }
  // ---- smoother of the linearised model: ay = y, aHH = H^2 (NlgApprox::View)
  SmootherIO io;
  io.n = n; io.m = m; io.k = k;
  io.Ztv = io.Ttv = io.Rtv = io.Dtv = io.Ctv = 1;
  io.Zb = Zv; io.Tb = Tv; io.Rb = Rv; io.Cb = Cv; io.Db = Dv; io.a1 = a1; io.P1 = P1; io.xb = nullptr;
  io.ay = p.y; io.aHH = HHv;
  io.ah = p.alphahat + (size_t)f * m * n1; io.Pt = p.Lt + (size_t)f * mm * n1; io.ccov = p.Ct + (size_t)f * mm * n1;
  io.sF = sF; io.sv = sv; io.sK = sK;
  smoother_ccov_conditional<M_, true>(io);
  // ---- the 32 doubles per time point (pack_nlg_step, api.cu; scales = 0 for approximate_for_is)
  const double LOG2PI = 1.8378770664093453;
  double* cb = p.stepc + (size_t)f * 32 * n;
  for (int t = 0; t < n; ++t) {
    double* c = cb + (size_t)32 * t;
    for (int e = 0; e < 32; ++e) c[e] = 0.0;
    const double Ht = Hv[t], HHt = Ht * Ht;
    c[0] = p.y[t]; c[1] = LOG2PI + 2.0 * pf_log(Ht); c[2] = Ht; c[3] = LOG2PI + pf_log(HHt); c[4] = HHt;
    c[5] = Dv[t]; c[8] = 0.0;
    for (int d = 0; d < m; ++d) c[6 + d] = Zv[(size_t)m * t + d];
    if (t >= 1) {
      for (int d = 0; d < m; ++d) c[10 + d] = Cv[(size_t)m * (t - 1) + d];
      for (int e = 0; e < mm; ++e) c[12 + e] = Tv[(size_t)mm * (t - 1) + e];
      double rr[M_ * M_], Rt[M_ * K_], cov[M_ * M_];
      MvnFactorDev<M_> fa, ft;
      pf_mmt(Rv + (size_t)mk * (t - 1), Rv + (size_t)mk * (t - 1), rr, m, k, m);
      fa.set(rr);
      fn.R_fn(t - 1, mode + (size_t)m * (t - 1), Rt);
      pf_mmt(Rt, Rt, cov, m, k, m);
      ft.set(cov);
      if (m == 1) {
        c[16] = fa.var1;
        c[21] = ft.var1;
        c[26] = fn.theta[0]; c[27] = fn.theta[1];
        if (p.model_id == BSSM_NLG_LINEAR_GAUSSIAN) { c[26] = 0.0; c[27] = 1.0; }
      } else {
        for (int e = 0; e < mm; ++e) { c[16 + e] = fa.rooti[e]; c[21 + e] = ft.rooti[e]; }
      }
      c[20] = fa.cst;
      c[25] = ft.cst;
    }
  }
}

} // namespace pf

// ------------------------------------------------------------------ ssm_mng
// approximate_for_is of an ssm_mng model per stored draw (pfga::MngApprox::run(mode), include/bssm_b200/
// mng_gauss_approx.h; src/model_ssm_mng.cpp:119-135, 216-315, src/model_ssm_mlg.cpp:232-397): laplace_iter per
// series, scales, the multivariate Kalman filter with masked missing series (Cholesky of F_t), the backward
// recursion and conditional_cov.  One thread per draw, run-time (m, p); the same expressions and primitives as
// the host header.
#include "bssm_b200/mng_gauss_approx.h"
namespace pf {

struct MngApproxParams {
  int n, m, k, p;
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double *y, *u; // n x p
  const int* distribution;
  const double *Z, *T, *R, *a1, *P1, *D, *C, *phi; // per draw with stride fs_* (0 = shared)
  size_t fs_Z, fs_T, fs_R, fs_a1, fs_P1, fs_D, fs_C, fs_phi;
  const double* modes; // p x n per draw
  double *alphahat, *Lt, *Ct, *ytilde, *Htilde, *scales; // per draw: m (n+1), m m (n+1) twice, n p twice (layout of y), n
  double* scratch; // per draw: (3 p + 1 + 2 p + 3 p m) n
};

template <int MAXM, int MAXP>
__global__ void __launch_bounds__(32) mng_approx_for_is_kernel(const MngApproxParams q, uint32_t nf) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  constexpr int MM = MAXM * MAXM, PP = MAXP * MAXP, PM = MAXP * MAXM;
  const int n = q.n, m = q.m, k = q.k, p = q.p;
  const size_t mm = (size_t)m * m, pm = (size_t)p * m, n1 = (size_t)n + 1;
  const double* Zb = q.Z + (size_t)f * q.fs_Z;
  const double* Tb = q.T + (size_t)f * q.fs_T;
  const double* Rb = q.R + (size_t)f * q.fs_R;
  const double* Cb = q.C + (size_t)f * q.fs_C;
  const double* Db = q.D + (size_t)f * q.fs_D;
  const double* a1 = q.a1 + (size_t)f * q.fs_a1;
  const double* P1 = q.P1 + (size_t)f * q.fs_P1;
  const double* phi = q.phi + (size_t)f * q.fs_phi;
  auto Z = [&](int t) { return Zb + pm * (t * q.Ztv); };
  auto T = [&](int t) { return Tb + mm * (t * q.Ttv); };
  auto R = [&](int t) { return Rb + (size_t)m * k * (t * q.Rtv); };
  auto C = [&](int t) { return Cb + (size_t)m * (t * q.Ctv); };
  auto D = [&](int t) { return Db + (size_t)p * (t * q.Dtv); };
  auto y = [&](int j, int t) { return q.y[t + (size_t)n * j]; };
  auto u = [&](int j, int t) { return q.u[t + (size_t)n * j]; };
  const double* mode = q.modes + (size_t)f * p * n;
  double* ah = q.alphahat + (size_t)f * m * n1;
  double* Pt = q.Lt + (size_t)f * mm * n1;
  double* ccov = q.Ct + (size_t)f * mm * n1;
  double* sc = q.scales + (size_t)f * n;
  const size_t per_t = (size_t)(3 * p + 1 + 2 * p) + 3 * pm;
  double* base = q.scratch + (size_t)f * per_t * n;
  double* ay = base;                      // p n, element (j, t) at j + p t
  double* aHH = ay + (size_t)p * n;
  double* aH = aHH + (size_t)p * n;
  double* s_nobs = aH + (size_t)p * n;    // n
  double* s_hh = s_nobs + n;              // p n
  double* s_v = s_hh + (size_t)p * n;     // p n
  double* s_Zt = s_v + (size_t)p * n;     // p m n
  double* s_ZF = s_Zt + pm * n;           // m p n
  double* s_K = s_ZF + pm * n;            // m p n

  // ---- laplace_iter (src/model_ssm_mng.cpp:268-315) and update_scales (:216-256)
  for (int j = 0; j < p; ++j) {
    const double ph = phi[j];
    for (int t = 0; t < n; ++t) {
      const size_t e = j + (size_t)p * t;
      const double yy = y(j, t), uu = u(j, t), s = mode[e];
      switch (q.distribution[j]) {
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
          aHH[e] = (ph + ex) * (ph + ex) / (ph * ex * (yy + ph));
          ay[e] = s + (ph + ex) * (yy - ex) / ((yy + ph) * ex);
        } break;
        case 4: {
          const double ex = pf_exp(s) * uu;
          aHH[e] = ex / (yy * ph);
          ay[e] = s - ex / yy + 1;
        } break;
        default: {
          aHH[e] = ph * ph;
          ay[e] = yy;
        } break;
      }
      aH[e] = sqrt(aHH[e]);
    }
  }
  for (int t = 0; t < n; ++t) {
    double sct = 0.0;
    for (int j = 0; j < p; ++j) {
      const size_t e = j + (size_t)p * t;
      if (pf_isfinite(y(j, t))) {
        const double yy = y(j, t), uu = u(j, t), s = mode[e], ph = phi[j];
        double lo;
        switch (q.distribution[j]) {
          case 1: lo = yy * s - uu * pf_exp(s); break;
          case 2: lo = yy * s - uu * pf_log1p(pf_exp(s)); break;
          case 3: lo = yy * s - (yy + ph) * pf_log(ph + uu * pf_exp(s)); break;
          case 4: lo = -ph * s - (yy * ph * pf_exp(-s) / uu); break;
          default: { const double qq = (yy - s) / ph; lo = -0.5 * (qq * qq); } break;
        }
        sct += lo;
        const double qd = (ay[e] - mode[e]) / aH[e];
        sct += 0.5 * (qd * qd);
      }
    }
    sc[t] = sct;
  }
  // pseudo-observations in the layout of y (n x p)
  {
    double* yo = q.ytilde + (size_t)f * n * p;
    double* Ho = q.Htilde + (size_t)f * n * p;
    for (int t = 0; t < n; ++t)
      for (int j = 0; j < p; ++j) {
        yo[t + (size_t)n * j] = ay[j + (size_t)p * t];
        Ho[t + (size_t)n * j] = aH[j + (size_t)p * t];
      }
  }

  // ---- forward pass: Kalman filter with the missing series masked (gain / predict_mean / predict_cov)
  double A[MM], B[MM], X[MM], E[MM], rr[MM];
  double at[MAXM], Pc[MM];
  for (int i = 0; i < m; ++i) { at[i] = a1[i]; ah[i] = at[i]; }
  for (size_t e = 0; e < mm; ++e) { Pc[e] = P1[e]; Pt[e] = Pc[e]; }
  bool failed = false;
  for (int t = 0; t < n && !failed; ++t) {
    double* Zt = s_Zt + pm * t;
    double* hh = s_hh + (size_t)p * t;
    double* v = s_v + (size_t)p * t;
    double* ZFinv = s_ZF + pm * t;
    double* K = s_K + pm * t;
    int nobs = 0;
    for (size_t e = 0; e < pm; ++e) Zt[e] = Z(t)[e];
    for (int j = 0; j < p; ++j) hh[j] = 1.0;
    for (int j = 0; j < p; ++j) {
      if (pf_isfinite(y(j, t))) { nobs++; hh[j] = aHH[j + (size_t)p * t]; }
      else for (int d = 0; d < m; ++d) Zt[j + (size_t)p * d] = 0.0;
    }
    s_nobs[t] = (double)nobs;
    if (nobs > 0) {
      double Apm[PM], F[PP], cholF[PP], inv[PP], za[MAXP], W[PM];
      pf_mm(Zt, Pc, Apm, p, m, m);
      pf_mmt(Apm, Zt, F, p, m, p);
      for (int j = 0; j < p; ++j) F[j + (size_t)p * j] += hh[j];
      for (int e = 0; e < p * p; ++e) if (!pf_isfinite(F[e])) failed = true;
      for (int j = 0; j < p; ++j) if (!(F[j + (size_t)p * j] > 0)) failed = true;
      if (!failed && !pfga::chol_upper(F, p, cholF)) failed = true;
      if (failed) break;
      pfga::inv_upper(cholF, p, inv);
      pf_mm(Zt, at, za, p, m, 1);
      for (int j = 0; j < p; ++j) v[j] = pf_isfinite(y(j, t)) ? (ay[j + (size_t)p * t] - D(t)[j]) - za[j] : 0.0;
      pf_mtm(Zt, inv, W, m, p, p);      /* Zt' inv */
      pf_mmt(W, inv, ZFinv, m, p, p);   /* (Zt' inv) inv' */
      pf_mm(Pc, ZFinv, K, m, m, p);
    }
    // predict_mean
    double tmpv[MAXM], kv[MAXM], ta[MAXM], an[MAXM];
    for (int i = 0; i < m; ++i) tmpv[i] = at[i];
    if (nobs > 0) {
      pf_mm(K, v, kv, m, p, 1);
      for (int i = 0; i < m; ++i) tmpv[i] = at[i] + kv[i];
    }
    pf_mm(T(t), tmpv, ta, m, m, 1);
    for (int i = 0; i < m; ++i) { an[i] = C(t)[i] + ta[i]; ah[(size_t)m * (t + 1) + i] = an[i]; }
    // predict_cov
    if (nobs > 0) {
      double kz[MM], KH[PM], khk[MM];
      pf_mm(K, Zt, kz, m, p, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - kz[i + (size_t)m * j];
      pf_mm(E, Pc, A, m, m, m);
      pf_mmt(A, E, B, m, m, m);
      for (int j = 0; j < p; ++j)
        for (int i = 0; i < m; ++i) KH[i + (size_t)m * j] = K[i + (size_t)m * j] * hh[j];
      pf_mmt(KH, K, khk, m, p, m);
      for (size_t e = 0; e < mm; ++e) X[e] = B[e] + khk[e];
    } else {
      for (size_t e = 0; e < mm; ++e) X[e] = Pc[e];
    }
    pf_mm(T(t), X, A, m, m, m);
    pf_mmt(A, T(t), B, m, m, m);
    pf_mmt(R(t), R(t), rr, m, k, m);
    double Pn[MM];
    for (size_t e = 0; e < mm; ++e) Pn[e] = B[e] + rr[e];
    pf_symmatu(Pn, m);
    for (size_t e = 0; e < mm; ++e) { Pt[mm * (t + 1) + e] = Pn[e]; ccov[mm * t + e] = Pn[e]; Pc[e] = Pn[e]; }
    for (int i = 0; i < m; ++i) at[i] = an[i];
  }
  if (failed) { // a Cholesky of F_t failed: everything infinite, as on the host
    const double inf = pf_inf();
    for (size_t e = 0; e < (size_t)m * n1; ++e) ah[e] = inf;
    for (size_t e = 0; e < mm * n1; ++e) { Pt[e] = inf; ccov[e] = inf; }
    return;
  }

  // ---- backward pass
  {
    double rt[MAXM], Nt[MM], Lm[MM], lr[MAXM], tmpv[MAXM], zv[MAXM], zz[MM];
    for (int i = 0; i < m; ++i) rt[i] = 0.0;
    for (size_t e = 0; e < mm; ++e) Nt[e] = 0.0;
    for (int t = n - 1; t >= 0; --t) {
      const double* Zt = s_Zt + pm * t;
      const double* v = s_v + (size_t)p * t;
      const double* ZFinv = s_ZF + pm * t;
      const double* K = s_K + pm * t;
      const bool used = s_nobs[t] > 0.0;
      const double* Tm;
      if (used) {
        double kz[MM];
        pf_mm(K, Zt, kz, m, p, m);
        for (int j = 0; j < m; ++j)
          for (int i = 0; i < m; ++i) E[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - kz[i + (size_t)m * j];
        pf_mm(T(t), E, Lm, m, m, m);
        Tm = Lm;
      } else {
        Tm = T(t);
      }
      double* Pcur = Pt + mm * t;
      double* cct = ccov + mm * t;
      pf_mm(Nt, cct, A, m, m, m);
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) X[i + (size_t)m * j] = (i == j ? 1.0 : 0.0) - A[i + (size_t)m * j];
      pf_mmt(Pcur, Tm, B, m, m, m);
      pf_mm(B, X, cct, m, m, m);
      pf_mtm(Tm, rt, lr, m, m, 1);
      pf_mtm(Tm, Nt, A, m, m, m);
      pf_mm(A, Tm, B, m, m, m);
      if (used) {
        pf_mm(ZFinv, v, zv, m, p, 1);
        pf_mm(ZFinv, Zt, zz, m, p, m);
        for (int i = 0; i < m; ++i) rt[i] = zv[i] + lr[i];
        for (size_t e = 0; e < mm; ++e) Nt[e] = zz[e] + B[e];
      } else {
        for (int i = 0; i < m; ++i) rt[i] = lr[i];
        for (size_t e = 0; e < mm; ++e) Nt[e] = B[e];
      }
      pf_symmatu(Nt, m);
      pf_mm(Pcur, rt, tmpv, m, m, 1);
      for (int i = 0; i < m; ++i) ah[(size_t)m * t + i] += tmpv[i];
      pf_mm(Pcur, Nt, A, m, m, m);
      pf_mm(A, Pcur, B, m, m, m);
      pf_symmatu(B, m);
      for (size_t e = 0; e < mm; ++e) Pcur[e] -= B[e];
    }
    for (size_t e = 0; e < mm; ++e) ccov[mm * n + e] = 0.0;
  }

  // ---- conditional_cov, SVD branch (pfga::conditional_cov_svd)
  {
    const double eps = 2.220446049250313e-16;
    double vals[MAXM], U[MM], W[MM];
    auto factor = [&](double* V) {
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
