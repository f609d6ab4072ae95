// models_ung_dyn.cuh -- ssm_ung family with a state dimension chosen at run time (m <= MAXM,
// k <= m): the bootstrap filter and the psi-APF of models whose (m, k) has no specialised
// instantiation in models_ung.cuh -- bsm_ng with a seasonal component has m = 2 + period - 1
// (src/model_bsm_ng.cpp:17-51: 5 for quarterly, 13 for monthly data).  Same formulas, same random
// stream layout (flat normal index i * K + d, K = m for the initial draw and the psi-APF, k for
// bootstrap steps) and therefore the same bits as the static policies and the oracle; loops run to
// the run-time m, so the per-particle arrays live in local memory -- this is the general path, not
// the fast one (the batched (N x m)(m x m) products are measured in tools/bench_mma.cu).
#pragma once
#include "models_ung.cuh"

namespace pf {

template <int MAXM>
struct UngDynBase {
  static constexpr int M = MAXM;
  static constexpr bool DYN = true;
  typedef UngParams Params;
  const UngParams& mp;
  int m, k;
  const double *pZ, *pD; // this filter's Z and D (parameter sets: one model per filter, else the shared arrays)
  template <int P> struct Noise { double z[P * MAXM]; };
  __device__ UngDynBase(const UngParams& p, uint32_t f) : mp(p), m(p.m_dyn), k(p.k_dyn) {
    pZ = p.Z + (size_t)f * p.fs_Z;
    pD = p.D + (size_t)f * p.fs_D;
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ __forceinline__ double finish_logw(double w, double) const { return w; }
  // `cnt` standard normals with flat indices f0, f0 + 1, ... (f0 even) of column t
  __device__ void normals(const StepCtx& c, uint32_t f0, int cnt, double* z) const {
    for (int q = 0; q < (cnt + 1) / 2; ++q) {
      double z0, z1;
      pf_normal_pair(c.seed, c.filter, c.t, 0u, (f0 >> 1) + q, &z0, &z1);
      z[2 * q] = z0;
      if (2 * q + 1 < cnt) z[2 * q + 1] = z1;
    }
  }
  __device__ double signal(const double* x, double Dt, double xb, int t) const {
    if (mp.distribution == 0) return x[0];
    const double* Z = pZ + (size_t)m * (t * mp.Ztv);
    double zs = 0.0;
    for (int d = 0; d < m; ++d) zs += Z[d] * x[d];
    return Dt + zs + xb;
  }
};

// ---------------------------------------------------------------- bootstrap filter
template <int MAXM>
struct UngBsfDyn : UngDynBase<MAXM> {
  typedef UngDynBase<MAXM> B;
  using B::mp; using B::m; using B::k; using B::pZ; using B::pD;
  double y, u, yphi2, Dt, xb, HHt, hc, phi, lconst;
  const double *pT, *pR, *pC, *pa1, *pLP1, *p_xb;
  int tw, ts;
  __device__ UngBsfDyn(const UngParams& p, uint32_t f) : B(p, f), hc(0.0), tw(0), ts(0) {
    phi = p.phi_f ? p.phi_f[f] : p.phi;
    lconst = p.lconst_f ? p.lconst_f[f] : p.loglik_const;
    pT = p.T + (size_t)f * p.fs_T; pR = p.R + (size_t)f * p.fs_R; pC = p.C + (size_t)f * p.fs_C;
    pa1 = p.a1 + (size_t)f * p.fs_a1; pLP1 = p.LP1 + (size_t)f * p.fs_LP1;
    p_xb = p.xbeta + (size_t)f * p.fs_x;
  }
  __device__ double step_const() const { return hc; }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return lconst; }
  template <int P>
  __device__ void init_all(const StepCtx& c, double (&x)[P][MAXM]) const {
    double z[P * MAXM];
    B::normals(c, c.i0 * (uint32_t)m, P * m, z); // i0 * m is even: P is
    for (int p = 0; p < P; ++p)
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += pLP1[d + m * j] * z[p * m + j];
        x[p][d] = pa1[d] + s;
      }
  }
  __device__ void begin_weight(int t) {
    tw = t;
    y = mp.y[t];
    u = mp.u[t];
    double q = y / phi;
    yphi2 = q * q;
    Dt = pD[t * mp.Dtv];
    xb = p_xb[t];
    if (mp.distribution == 5) { HHt = mp.HH[t]; hc = mp.hconst[t]; }
  }
  __device__ double logw(const double* x) const {
    if (mp.distribution == 5) { // ssm_ulg::bsf_filter, src/model_ssm_ulg.cpp:577-580: mu = D + Z'a
      const double* Z = pZ + (size_t)m * (tw * mp.Ztv);
      double zs = 0.0;
      for (int d = 0; d < m; ++d) zs += Z[d] * x[d];
      const double mu = Dt + zs;
      return -0.5 * ((y - mu) * (y - mu)) / HHt;
    }
    return ung_log_obs(mp.distribution, B::signal(x, Dt, xb, tw), y, u, phi, yphi2);
  }
  __device__ void begin_step(int t) { ts = t; }
  template <int P>
  __device__ void draw(const StepCtx& c, typename B::template Noise<P>& nz) const {
    B::normals(c, c.i0 * (uint32_t)k, P * k, nz.z);
  }
  template <int P>
  __device__ void propagate_all(const StepCtx&, const typename B::template Noise<P>& nz, const double (&xa)[P][MAXM],
                                double (&x)[P][MAXM], double (&carry)[P]) const {
    const double* T = pT + (size_t)m * m * (ts * mp.Ttv);
    const double* R = pR + (size_t)m * k * (ts * mp.Rtv);
    const double* C = pC + (size_t)m * (ts * mp.Ctv);
    for (int p = 0; p < P; ++p) {
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += T[d + m * j] * xa[p][j];
        for (int j = 0; j < k; ++j) ru += R[d + m * j] * nz.z[p * k + j];
        x[p][d] = C[d] + ta + ru;
      }
      carry[p] = 0.0;
    }
  }
};

// ------------------------------------------------------------------------ psi-APF
// alpha_{t+1} = ahat_{t+1} + Ct_{t+1} (alpha_anc - ahat_t) + L_{t+1} u, u ~ N(0, I_m)
// (src/model_ssm_ung.cpp:580-587), Gaussian approximation from the host (ung_gauss_approx.h)
template <int MAXM>
struct UngPsiDyn : UngDynBase<MAXM> {
  typedef UngDynBase<MAXM> B;
  using B::mp; using B::m; using B::pD;
  double al0, phi;
  const double *p_ah, *p_Lt, *p_Ct, *p_yt, *p_Ht, *p_sc, *p_xb;
  double y, u, yphi2, Dt, xb, ytil, rH, scale;
  int tw, ts;
  __device__ UngPsiDyn(const UngParams& p, uint32_t f) : B(p, f), tw(0), ts(0) {
    al0 = p.approx_loglik[f];
    phi = p.phi_f ? p.phi_f[f] : p.phi;
    p_ah = p.alphahat + (size_t)f * p.fs_a;
    p_Lt = p.Lt + (size_t)f * p.fs_m; p_Ct = p.Ct + (size_t)f * p.fs_m;
    p_yt = p.ytilde + (size_t)f * p.fs_n; p_Ht = p.Htilde + (size_t)f * p.fs_n; p_sc = p.scales + (size_t)f * p.fs_n;
    p_xb = p.xbeta + (size_t)f * p.fs_x;
  }
  __device__ double step_const() const { return 0.0; }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  template <int P>
  __device__ void init_all(const StepCtx& c, double (&x)[P][MAXM]) const {
    double z[P * MAXM];
    B::normals(c, c.i0 * (uint32_t)m, P * m, z);
    for (int p = 0; p < P; ++p)
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += p_Lt[d + m * j] * z[p * m + j];
        x[p][d] = p_ah[d] + s;
      }
  }
  __device__ void begin_weight(int t) {
    tw = t;
    y = mp.y[t];
    u = mp.u[t];
    double q = y / phi;
    yphi2 = q * q;
    Dt = pD[t * mp.Dtv];
    xb = p_xb[t];
    ytil = p_yt[t]; rH = 1.0 / p_Ht[t]; scale = p_sc[t];
  }
  __device__ double logw(const double* x) const {
    const double s = B::signal(x, Dt, xb, tw);
    const double q = (ytil - s) * rH;
    return (ung_log_obs(mp.distribution, s, y, u, phi, yphi2) + 0.5 * (q * q)) - scale;
  }
  __device__ void begin_step(int t) { ts = t; }
  template <int P>
  __device__ void draw(const StepCtx& c, typename B::template Noise<P>& nz) const {
    B::normals(c, c.i0 * (uint32_t)m, P * m, nz.z);
  }
  template <int P>
  __device__ void propagate_all(const StepCtx&, const typename B::template Noise<P>& nz, const double (&xa)[P][MAXM],
                                double (&x)[P][MAXM], double (&carry)[P]) const {
    const double* ah0 = p_ah + (size_t)m * ts;
    const double* ah1 = p_ah + (size_t)m * (ts + 1);
    const double* Cm = p_Ct + (size_t)m * m * (ts + 1);
    const double* Lm = p_Lt + (size_t)m * m * (ts + 1);
    for (int p = 0; p < P; ++p) {
      double diff[MAXM];
      for (int j = 0; j < m; ++j) diff[j] = xa[p][j] - ah0[j];
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        for (int j = 0; j < m; ++j) cd += Cm[d + m * j] * diff[j];
        for (int j = 0; j < m; ++j) lu += Lm[d + m * j] * nz.z[p * m + j];
        x[p][d] = ah1[d] + cd + lu;
      }
      carry[p] = 0.0;
    }
  }
};

// ------------------------------------------------------------- ssm_mng, run-time m
// MngBsf / MngPsi (models_ung.cuh) with the state dimension chosen at run time (m <= MAXM, k <= m, any p):
// multivariate structural models (a trend and a seasonal per series) reach m = 8-16.  The per-filter strides
// of MngParams are honoured (0 = shared), so the same two policies serve the parameter-set jobs.
template <int MAXM>
struct MngDynBase {
  static constexpr int M = MAXM;
  static constexpr bool DYN = true;
  typedef MngParams Params;
  const MngParams& mp;
  int m, k, tw, ts;
  const double *pZ, *pD, *pphi;
  template <int P> struct Noise { double z[P * MAXM]; };
  __device__ MngDynBase(const MngParams& p, uint32_t f) : mp(p), m(p.m_dyn), k(p.k_dyn), tw(0), ts(0) {
    pZ = p.Z + (size_t)f * p.fs_Z; pD = p.D + (size_t)f * p.fs_D; pphi = p.phi + (size_t)f * p.fs_phi;
  }
  __device__ bool observed(int t) const {
    for (int j = 0; j < mp.p; ++j)
      if (pf_isfinite(mp.y[t + (size_t)mp.n * j])) return true;
    return false;
  }
  __device__ __forceinline__ double finish_logw(double w, double) const { return w; }
  __device__ double step_const() const { return 0.0; }
  __device__ void normals(const StepCtx& c, uint32_t f0, int cnt, double* z) const {
    for (int q = 0; q < (cnt + 1) / 2; ++q) {
      double z0, z1;
      pf_normal_pair(c.seed, c.filter, c.t, 0u, (f0 >> 1) + q, &z0, &z1);
      z[2 * q] = z0;
      if (2 * q + 1 < cnt) z[2 * q + 1] = z1;
    }
  }
  // sum over the observed series of log g(y_j | s_j) [+ 0.5 ((ytilde_j - s_j) / Htilde_j)^2 when yt != null]
  __device__ double obs_sum(const double* x, const double* yt, const double* Ht) const {
    double w = 0.0;
    const double* Zt = pZ + (size_t)mp.p * m * (tw * mp.Ztv);
    const double* Dt = pD + (size_t)mp.p * (tw * mp.Dtv);
    for (int j = 0; j < mp.p; ++j) {
      const size_t e = tw + (size_t)mp.n * j;
      const double y = mp.y[e];
      if (!pf_isfinite(y)) continue;
      double zs = 0.0;
      for (int d = 0; d < m; ++d) zs += Zt[j + mp.p * d] * x[d];
      const double s = Dt[j] + zs;
      w += mng_log_obs(mp.distribution[j], s, y, mp.u[e], pphi[j]);
      if (yt) {
        const double q = (yt[e] - s) / Ht[e];
        w += 0.5 * (q * q);
      }
    }
    return w;
  }
};

template <int MAXM>
struct MngBsfDyn : MngDynBase<MAXM> {
  typedef MngDynBase<MAXM> B;
  using B::mp; using B::m; using B::k; using B::tw; using B::ts;
  const double *pT, *pR, *pC, *pa1, *pLP1;
  double lconst;
  __device__ MngBsfDyn(const MngParams& p, uint32_t f) : B(p, f) {
    pT = p.T + (size_t)f * p.fs_T; pR = p.R + (size_t)f * p.fs_R; pC = p.C + (size_t)f * p.fs_C;
    pa1 = p.a1 + (size_t)f * p.fs_a1; pLP1 = p.LP1 + (size_t)f * p.fs_LP1;
    lconst = p.lconst_f ? p.lconst_f[f] : p.loglik_const;
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return lconst; }
  template <int P>
  __device__ void init_all(const StepCtx& c, double (&x)[P][MAXM]) const {
    double z[P * MAXM];
    B::normals(c, c.i0 * (uint32_t)m, P * m, z);
    for (int p = 0; p < P; ++p)
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += pLP1[d + m * j] * z[p * m + j];
        x[p][d] = pa1[d] + s;
      }
  }
  __device__ void begin_weight(int t) { tw = t; }
  __device__ double logw(const double* x) const { return B::obs_sum(x, nullptr, nullptr); }
  __device__ void begin_step(int t) { ts = t; }
  template <int P>
  __device__ void draw(const StepCtx& c, typename B::template Noise<P>& nz) const {
    B::normals(c, c.i0 * (uint32_t)k, P * k, nz.z);
  }
  template <int P>
  __device__ void propagate_all(const StepCtx&, const typename B::template Noise<P>& nz, const double (&xa)[P][MAXM],
                                double (&x)[P][MAXM], double (&carry)[P]) const {
    const double* T = pT + (size_t)m * m * (ts * mp.Ttv);
    const double* R = pR + (size_t)m * k * (ts * mp.Rtv);
    const double* C = pC + (size_t)m * (ts * mp.Ctv);
    for (int p = 0; p < P; ++p) {
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += T[d + m * j] * xa[p][j];
        for (int j = 0; j < k; ++j) ru += R[d + m * j] * nz.z[p * k + j];
        x[p][d] = C[d] + ta + ru;
      }
      carry[p] = 0.0;
    }
  }
};

template <int MAXM>
struct MngPsiDyn : MngDynBase<MAXM> {
  typedef MngDynBase<MAXM> B;
  using B::mp; using B::m; using B::tw; using B::ts;
  const double *p_ah, *p_Lt, *p_Ct, *p_yt, *p_Ht, *p_sc;
  double al0, scale;
  __device__ MngPsiDyn(const MngParams& p, uint32_t f) : B(p, f), scale(0.0) {
    al0 = p.approx_loglik[f];
    p_ah = p.alphahat + (size_t)f * p.fs_a;
    p_Lt = p.Lt + (size_t)f * p.fs_m; p_Ct = p.Ct + (size_t)f * p.fs_m;
    p_yt = p.ytilde + (size_t)f * p.fs_np; p_Ht = p.Htilde + (size_t)f * p.fs_np; p_sc = p.scales + (size_t)f * p.fs_n;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  template <int P>
  __device__ void init_all(const StepCtx& c, double (&x)[P][MAXM]) const {
    double z[P * MAXM];
    B::normals(c, c.i0 * (uint32_t)m, P * m, z);
    for (int p = 0; p < P; ++p)
      for (int d = 0; d < m; ++d) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += p_Lt[d + m * j] * z[p * m + j];
        x[p][d] = p_ah[d] + s;
      }
  }
  __device__ void begin_weight(int t) { tw = t; scale = p_sc[t]; }
  __device__ double logw(const double* x) const { return B::obs_sum(x, p_yt, p_Ht) - scale; }
  __device__ void begin_step(int t) { ts = t; }
  template <int P>
  __device__ void draw(const StepCtx& c, typename B::template Noise<P>& nz) const {
    B::normals(c, c.i0 * (uint32_t)m, P * m, nz.z);
  }
  template <int P>
  __device__ void propagate_all(const StepCtx&, const typename B::template Noise<P>& nz, const double (&xa)[P][MAXM],
                                double (&x)[P][MAXM], double (&carry)[P]) const {
    const double* ah0 = p_ah + (size_t)m * ts;
    const double* ah1 = p_ah + (size_t)m * (ts + 1);
    const double* Cm = p_Ct + (size_t)m * m * (ts + 1);
    const double* Lm = p_Lt + (size_t)m * m * (ts + 1);
    for (int p = 0; p < P; ++p) {
      double diff[MAXM];
      for (int j = 0; j < m; ++j) diff[j] = xa[p][j] - ah0[j];
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        for (int j = 0; j < m; ++j) cd += Cm[d + m * j] * diff[j];
        for (int j = 0; j < m; ++j) lu += Lm[d + m * j] * nz.z[p * m + j];
        x[p][d] = ah1[d] + cd + lu;
      }
      carry[p] = 0.0;
    }
  }
};

} // namespace pf