// models_ung.cuh -- device model policies for the ssm_ung family
// (ssm_ung / bsm_ng / svm / ar1_ng): linear-Gaussian state, exponential-family
// observation.  Formulas follow /root/reference/src/model_ssm_ung.cpp:
//   log_obs_density :447-493, log_weights :385-439,
//   bsf_filter propagation :663-670, psi_filter propagation :580-587.
#pragma once
#include "pf_device.cuh"

namespace pf {

struct UngParams {
  int n, distribution;
  int m_dyn, k_dyn;    // state / noise dimension of the run-time-m policies (models_ung_dyn.cuh)
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double* y;     // n
  const double* Z;     // m x (1|n)
  const double* T;     // m x m x (1|n)
  const double* R;     // m x k x (1|n)
  const double* a1;    // m
  const double* LP1;   // m x m lower Cholesky factor of P1 (positive-diagonal block)
  const double* D;     // (1|n)
  const double* C;     // m x (1|n)
  const double* xbeta; // n (zeros when no regressors)
  const double* u;     // n
  double phi;
  double loglik_const; // BSF: state-independent constants, added at the end (:690-710)
  // distribution 5 (ssm_ulg): HH_t = H_t^2 and norm_log_const(H_t), n each
  const double* HH;
  const double* hconst;
  // psi-APF (m = 1), per filter with stride `pstride`
  const double* phi_f;     // per-filter phi (post_correct) or null
  const double* alphahat;  // n+1
  const double* Lt;        // n+1
  const double* Ct;        // n+1
  const double* ytilde;    // n
  const double* Htilde;    // n
  const double* scales;    // n
  const double* approx_loglik; // per filter
  size_t pstride;
  const double* rHt;       // n: 1 / Htilde_t   (m = 1 pre-pass)
  const double* yphi2;     // n: (y_t / phi)^2  (m = 1 pre-pass)
  // psi-APF (m > 1): per-filter strides of alphahat / (Lt, Ct) / (ytilde, Htilde, scales) / xbeta;
  // all zero when the replicates of a job share one approximation
  size_t fs_a, fs_m, fs_n, fs_x;
  // bootstrap filter with one model per filter (is_correction_bsf, m = 1): 7 doubles per filter
  // {T, R, a1, sqrt(P1), C, phi, loglik_const}, or null
  const double* bsf_par;
  // parameter sets (parset_ung, src/parset_ng.h:11-40): per-filter strides of the model arrays, 0 = shared.
  // The psi-APF reads Z and D only (the rest enters through the approximation); the run-time-m
  // bootstrap filter reads all of them, with phi from phi_f and the constants from lconst_f.
  size_t fs_Z, fs_D, fs_T, fs_R, fs_C, fs_a1, fs_LP1;
  const double* lconst_f;
};

// log g(y_t | signal) up to state-independent constants, :447-493
__device__ __forceinline__ double ung_log_obs(int dist, double s, double y, double u, double phi,
                                              double yphi2) {
  switch (dist) {
    case 0: return -0.5 * (s + yphi2 * pf_exp(-s));
    case 1: return y * s - u * pf_exp(s);
    case 2: return y * s - u * pf_log1p(pf_exp(s));
    case 3: return y * s - (y + phi) * pf_log(phi + u * pf_exp(s));
    default: return -phi * (s + (y * pf_exp(-s) / u));
  }
}

// ---------------------------------------------------------- bootstrap filter
template <int M_, int K_>
struct UngBsf : GaussStep<UngBsf<M_, K_>, M_, M_, K_> {
  static constexpr int M = M_, KI = M_, KP = K_;
  typedef UngParams Params;
  const UngParams& mp;
  // per-step constants
  double y, u, yphi2, Dt, xb, HHt, hc;
  double Z[M_], T[M_ * M_], R[M_ * K_], C[M_];

  __device__ UngBsf(const UngParams& p, uint32_t) : mp(p), hc(0.0) {}
  __device__ double step_const() const { return hc; }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return mp.loglik_const; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) s += mp.LP1[d + M_ * j] * z[j];
      x[d] = mp.a1[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) {
    y = mp.y[t];
    u = mp.u[t];
    double q = y / mp.phi;
    yphi2 = q * q;
    Dt = mp.D[t * mp.Dtv];
    xb = mp.xbeta[t];
#pragma unroll
    for (int d = 0; d < M_; ++d) Z[d] = mp.Z[d + M_ * (t * mp.Ztv)];
    if (mp.distribution == 5) { HHt = mp.HH[t]; hc = mp.hconst[t]; }
  }
  __device__ double logw(const double* x) const {
    double s;
    if (mp.distribution == 0) {
      s = x[0];
    } else {
      double zs = 0.0;
#pragma unroll
      for (int d = 0; d < M_; ++d) zs += Z[d] * x[d];
      if (mp.distribution == 5) { // ssm_ulg::bsf_filter, src/model_ssm_ulg.cpp:577-580: mu = D + Z'a
        const double mu = Dt + zs;
        return -0.5 * ((y - mu) * (y - mu)) / HHt;
      }
      s = Dt + zs + xb;
    }
    return ung_log_obs(mp.distribution, s, y, u, mp.phi, yphi2);
  }
  __device__ void begin_step(int t) {
#pragma unroll
    for (int e = 0; e < M_ * M_; ++e) T[e] = mp.T[e + M_ * M_ * (t * mp.Ttv)];
#pragma unroll
    for (int e = 0; e < M_ * K_; ++e) R[e] = mp.R[e + M_ * K_ * (t * mp.Rtv)];
#pragma unroll
    for (int d = 0; d < M_; ++d) C[d] = mp.C[d + M_ * (t * mp.Ctv)];
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double ta = 0.0, ru = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) ta += T[d + M_ * j] * xa[j];
#pragma unroll
      for (int j = 0; j < K_; ++j) ru += R[d + M_ * j] * z[j];
      x[d] = C[d] + ta + ru;
    }
  }
};

// ------------------------------------- bootstrap filter, m = 1, one model per filter
// is_correction_bsf (src/approx_mcmc.cpp:447-518) runs ssm_ung::bsf_filter once per stored MCMC
// draw after update_model(theta_i): T, R, a1, P1, C and phi differ between the filters of a batch.
struct UngBsfDraw : GaussStep<UngBsfDraw, 1, 1, 1> {
  static constexpr int M = 1, KI = 1, KP = 1;
  typedef UngParams Params;
  const UngParams& mp;
  const double* xbp;
  double T, R, a1, L, C, phi, lconst;
  double y, u, yphi2, Dt, xb, Z0;

  __device__ UngBsfDraw(const UngParams& p, uint32_t f) : mp(p) {
    xbp = p.xbeta + (size_t)f * p.fs_x; // fs_x = n when beta is part of theta, else 0
    const double* q = p.bsf_par + (size_t)7 * f;
    T = q[0]; R = q[1]; a1 = q[2]; L = q[3]; C = q[4]; phi = q[5]; lconst = q[6];
  }
  __device__ double step_const() const { return 0.0; }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return lconst; }
  __device__ void init(const double* z, double* x) const { x[0] = a1 + L * z[0]; }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) {
    y = mp.y[t];
    u = mp.u[t];
    double q = y / phi;
    yphi2 = q * q;
    Dt = mp.D[t * mp.Dtv];
    xb = xbp[t];
    Z0 = mp.Z[t * mp.Ztv];
  }
  __device__ double logw(const double* x) const {
    const double s = (mp.distribution == 0) ? x[0] : (Dt + Z0 * x[0] + xb);
    return ung_log_obs(mp.distribution, s, y, u, phi, yphi2);
  }
  __device__ void begin_step(int) {}
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    x[0] = C + T * xa[0] + R * z[0];
  }
};

// ------------------------------------------------------------ psi-APF, m = 1
struct UngPsi1 : GaussStep<UngPsi1, 1, 1, 1> {
  static constexpr int M = 1, KI = 1, KP = 1;
  typedef UngParams Params;
  const UngParams& mp;
  const double *ah, *Lt, *Ct, *yt, *rHp, *yp2, *sc;
  double phi, al0;
  double y, u, yphi2, Dt, xb, Z0, ytil, rH, scale;
  double ah0, ah1, c1, l1;

  __device__ UngPsi1(const UngParams& p, uint32_t f) : mp(p) {
    const size_t o = (size_t)f * p.pstride;
    ah = p.alphahat + o; Lt = p.Lt + o; Ct = p.Ct + o;
    yt = p.ytilde + o; rHp = p.rHt + o; yp2 = p.yphi2 + o; sc = p.scales + o;
    phi = p.phi_f ? p.phi_f[f] : p.phi;
    al0 = p.approx_loglik[f];
    y = u = Dt = xb = Z0 = 0.0;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const { x[0] = ah[0] + Lt[0] * z[0]; }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) {
    // 1 / Htilde_t and (y_t / phi)^2 come from the pre-pass: no division per thread and step
    ytil = yt[t]; rH = rHp[t]; scale = sc[t]; yphi2 = yp2[t];
    if (mp.distribution != 0) {
      y = mp.y[t];
      u = mp.u[t];
      Dt = mp.D[t * mp.Dtv];
      xb = mp.xbeta[t];
      Z0 = mp.Z[t * mp.Ztv];
    }
  }
  __device__ double logw(const double* x) const {
    double s = (mp.distribution == 0) ? x[0] : (Dt + Z0 * x[0] + xb);
    double q = (ytil - s) * rH; // reciprocal once per time point, as in the oracle's log_weight_pf
    return (ung_log_obs(mp.distribution, s, y, u, phi, yphi2) + 0.5 * (q * q)) - scale;
  }
  __device__ void begin_step(int t) {
    ah0 = ah[t]; ah1 = ah[t + 1]; c1 = Ct[t + 1]; l1 = Lt[t + 1];
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    x[0] = ah1 + c1 * (xa[0] - ah0) + l1 * z[0];
  }
};

// ------------------------------------------------------------ psi-APF, m > 1
// alpha_{t+1} = ahat_{t+1} + Ct_{t+1} (alpha_anc - ahat_t) + L_{t+1} u, u ~ N(0, I_m)
// (src/model_ssm_ung.cpp:580-587); the Gaussian approximation (ahat, Ct, L, pseudo
// observations, scales) comes from the host (include/bssm_b200/ung_gauss_approx.h) and
// is shared by all replicate filters of a job (pstride = 0).
template <int M_>
struct UngPsiM : GaussStep<UngPsiM<M_>, M_, M_, M_> {
  static constexpr int M = M_, KI = M_, KP = M_;
  typedef UngParams Params;
  const UngParams& mp;
  double al0, phi;
  const double *p_ah, *p_Lt, *p_Ct, *p_yt, *p_Ht, *p_sc, *p_xb, *p_Z, *p_D;
  double y, u, yphi2, Dt, xb, ytil, rH, scale;
  double Z[M_], ah0[M_], ah1[M_], Cm[M_ * M_], Lm[M_ * M_];

  __device__ UngPsiM(const UngParams& p, uint32_t f) : mp(p) {
    al0 = p.approx_loglik[f];
    phi = p.phi_f ? p.phi_f[f] : p.phi;
    p_ah = p.alphahat + (size_t)f * p.fs_a;
    p_Lt = p.Lt + (size_t)f * p.fs_m; p_Ct = p.Ct + (size_t)f * p.fs_m;
    p_yt = p.ytilde + (size_t)f * p.fs_n; p_Ht = p.Htilde + (size_t)f * p.fs_n; p_sc = p.scales + (size_t)f * p.fs_n;
    p_xb = p.xbeta + (size_t)f * p.fs_x;
    p_Z = p.Z + (size_t)f * p.fs_Z; p_D = p.D + (size_t)f * p.fs_D;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) s += p_Lt[d + M_ * j] * z[j];
      x[d] = p_ah[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) {
    y = mp.y[t];
    u = mp.u[t];
    double q = y / phi;
    yphi2 = q * q;
    Dt = p_D[t * mp.Dtv];
    xb = p_xb[t];
#pragma unroll
    for (int d = 0; d < M_; ++d) Z[d] = p_Z[d + M_ * (t * mp.Ztv)];
    ytil = p_yt[t]; rH = 1.0 / p_Ht[t]; scale = p_sc[t];
  }
  __device__ double logw(const double* x) const {
    double s;
    if (mp.distribution == 0) {
      s = x[0];
    } else {
      double zs = 0.0;
#pragma unroll
      for (int d = 0; d < M_; ++d) zs += Z[d] * x[d];
      s = Dt + zs + xb;
    }
    double q = (ytil - s) * rH; // as UngPsi1
    return (ung_log_obs(mp.distribution, s, y, u, phi, yphi2) + 0.5 * (q * q)) - scale;
  }
  __device__ void begin_step(int t) {
#pragma unroll
    for (int d = 0; d < M_; ++d) { ah0[d] = p_ah[d + M_ * t]; ah1[d] = p_ah[d + M_ * (t + 1)]; }
#pragma unroll
    for (int e = 0; e < M_ * M_; ++e) {
      Cm[e] = p_Ct[e + M_ * M_ * (t + 1)];
      Lm[e] = p_Lt[e + M_ * M_ * (t + 1)];
    }
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    double diff[M_];
#pragma unroll
    for (int j = 0; j < M_; ++j) diff[j] = xa[j] - ah0[j];
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double cd = 0.0, lu = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) cd += Cm[d + M_ * j] * diff[j];
#pragma unroll
      for (int j = 0; j < M_; ++j) lu += Lm[d + M_ * j] * z[j];
      x[d] = ah1[d] + cd + lu;
    }
  }
};

// ------------------------------------------------- ssm_mng: bootstrap filter
// p observation series with their own exponential-family (or Gaussian) densities, one
// m-dimensional linear-Gaussian state: ssm_mng::bsf_filter src/model_ssm_mng.cpp:562-668,
// log_obs_density :407-442.  simsignal = D_t + Z_t alpha is a (p x m)(m x 1) product per
// particle; with m, p <= 8 this is a handful of fused loops per particle and stays on the
// FP64 pipe: there is no tile here that a tcgen05 MMA could usefully take (fp64, K = m <= 8).
struct MngParams {
  int n, p;
  int m_dyn, k_dyn; // state / noise dimensions for the run-time-m policies (models_ung_dyn.cuh)
  int Ztv, Ttv, Rtv, Dtv, Ctv;
  const double* y;     // n x p as R holds it (the reference transposes to p x n)
  const double* u;     // n x p
  const double* Z;     // p x m x (1|n)
  const double* T;     // m x m x (1|n)
  const double* R;     // m x k x (1|n)
  const double* a1;
  const double* LP1;
  const double* D;     // p x (1|n)
  const double* C;     // m x (1|n)
  const double* phi;   // p
  const int* distribution; // p (1 poisson, 2 binomial, 3 negbin, 4 gamma, 5 gaussian)
  double loglik_const;
  // psi-APF only: Gaussian approximation from the host (include/bssm_b200/mng_gauss_approx.h)
  const double* alphahat;      // m x (n+1)
  const double* Ct;            // m x m x (n+1)
  const double* Lt;            // m x m x (n+1)
  const double* ytilde;        // n x p pseudo-observations (same layout as y)
  const double* Htilde;        // n x p their standard deviations
  const double* scales;        // n
  const double* approx_loglik; // per filter
  // parameter sets (parset_mng, src/parset_ng.h:42-62; the *Draws policies): per-filter strides of the model
  // arrays and of the approximation, 0 = shared; lconst_f = per-filter loglik_const (phi enters it) or null
  size_t fs_Z, fs_T, fs_R, fs_C, fs_D, fs_a1, fs_LP1, fs_phi;
  size_t fs_a, fs_m, fs_np, fs_n; // alphahat / (Lt, Ct) / (ytilde, Htilde) / scales
  const double* lconst_f;
};
// the launch parameters with every per-filter array advanced to filter f
__device__ __forceinline__ void mng_load_draw(MngParams& loc, const MngParams& p, uint32_t f) {
  loc = p;
  loc.Z += (size_t)f * p.fs_Z; loc.T += (size_t)f * p.fs_T; loc.R += (size_t)f * p.fs_R; loc.C += (size_t)f * p.fs_C;
  loc.D += (size_t)f * p.fs_D; loc.a1 += (size_t)f * p.fs_a1; loc.LP1 += (size_t)f * p.fs_LP1;
  loc.phi += (size_t)f * p.fs_phi;
  if (p.lconst_f) loc.loglik_const = p.lconst_f[f];
  if (p.alphahat) {
    loc.alphahat += (size_t)f * p.fs_a; loc.Lt += (size_t)f * p.fs_m; loc.Ct += (size_t)f * p.fs_m;
    loc.ytilde += (size_t)f * p.fs_np; loc.Htilde += (size_t)f * p.fs_np; loc.scales += (size_t)f * p.fs_n;
  }
}

// log g(y_j | s) up to constants, src/model_ssm_mng.cpp:364-386 (shared by log_obs_density :407-442)
__device__ __forceinline__ double mng_log_obs(int dist, double s, double y, double u, double phi) {
  switch (dist) {
    case 1: return y * s - u * pf_exp(s);
    case 2: return y * s - u * pf_log1p(pf_exp(s));
    case 3: return y * s - (y + phi) * pf_log(phi + u * pf_exp(s));
    case 4: return -phi * (s + (y * pf_exp(-s) / u));
    default: { double q = (y - s) / phi; return -0.5 * (q * q); }
  }
}

template <int M_, int K_, bool PerFilter = false>
struct MngBsf : GaussStep<MngBsf<M_, K_, PerFilter>, M_, M_, K_> {
  static constexpr int M = M_, KI = M_, KP = K_;
  typedef MngParams Params;
  MngParams loc;
  const MngParams& mp;
  int tw;
  double T[M_ * M_], R[M_ * K_], C[M_];

  __device__ MngBsf(const MngParams& p, uint32_t f) : mp(PerFilter ? loc : p), tw(0) {
    if (PerFilter) mng_load_draw(loc, p, f);
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return mp.loglik_const; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) s += mp.LP1[d + M_ * j] * z[j];
      x[d] = mp.a1[d] + s;
    }
  }
  // weighted when at least one series is observed (:586, :620)
  __device__ bool observed(int t) const {
    for (int j = 0; j < mp.p; ++j)
      if (pf_isfinite(mp.y[t + (size_t)mp.n * j])) return true;
    return false;
  }
  __device__ void begin_weight(int t) { tw = t; }
  __device__ double logw(const double* x) const {
    double w = 0.0;
    const double* Zt = mp.Z + (size_t)mp.p * M_ * (tw * mp.Ztv);
    const double* Dt = mp.D + (size_t)mp.p * (tw * mp.Dtv);
    for (int j = 0; j < mp.p; ++j) {
      const double y = mp.y[tw + (size_t)mp.n * j];
      if (!pf_isfinite(y)) continue;
      double zs = 0.0;
#pragma unroll
      for (int d = 0; d < M_; ++d) zs += Zt[j + mp.p * d] * x[d];
      const double s = Dt[j] + zs;
      const double u = mp.u[tw + (size_t)mp.n * j], phi = mp.phi[j];
      w += mng_log_obs(mp.distribution[j], s, y, u, phi);
    }
    return w;
  }
  __device__ void begin_step(int t) {
#pragma unroll
    for (int e = 0; e < M_ * M_; ++e) T[e] = mp.T[e + M_ * M_ * (t * mp.Ttv)];
#pragma unroll
    for (int e = 0; e < M_ * K_; ++e) R[e] = mp.R[e + M_ * K_ * (t * mp.Rtv)];
#pragma unroll
    for (int d = 0; d < M_; ++d) C[d] = mp.C[d + M_ * (t * mp.Ctv)];
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double ta = 0.0, ru = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) ta += T[d + M_ * j] * xa[j];
#pragma unroll
      for (int j = 0; j < K_; ++j) ru += R[d + M_ * j] * z[j];
      x[d] = C[d] + ta + ru;
    }
  }
};

// ---------------------------------------------------- ssm_mng: psi-APF
// ssm_mng::psi_filter src/model_ssm_mng.cpp:461-559: the proposal of UngPsiM with the
// multivariate weights of log_weights (:357-393): per observed series
// log g(y_j | s_j) + 0.5 ((ytilde_j - s_j) / Htilde_j)^2, minus scales_t.
template <int M_, bool PerFilter = false>
struct MngPsi : GaussStep<MngPsi<M_, PerFilter>, M_, M_, M_> {
  static constexpr int M = M_, KI = M_, KP = M_;
  typedef MngParams Params;
  MngParams loc;
  const MngParams& mp;
  double al0, scale;
  int tw;
  double ah0[M_], ah1[M_], Cm[M_ * M_], Lm[M_ * M_];

  __device__ MngPsi(const MngParams& p, uint32_t f) : mp(PerFilter ? loc : p), tw(0) {
    if (PerFilter) mng_load_draw(loc, p, f);
    al0 = p.approx_loglik[f]; scale = 0.0;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) s += mp.Lt[d + M_ * j] * z[j];
      x[d] = mp.alphahat[d] + s;
    }
  }
  __device__ bool observed(int t) const {
    for (int j = 0; j < mp.p; ++j)
      if (pf_isfinite(mp.y[t + (size_t)mp.n * j])) return true;
    return false;
  }
  __device__ void begin_weight(int t) { tw = t; scale = mp.scales[t]; }
  __device__ double logw(const double* x) const {
    double w = 0.0;
    const double* Zt = mp.Z + (size_t)mp.p * M_ * (tw * mp.Ztv);
    const double* Dt = mp.D + (size_t)mp.p * (tw * mp.Dtv);
    for (int j = 0; j < mp.p; ++j) {
      const size_t e = tw + (size_t)mp.n * j;
      const double y = mp.y[e];
      if (!pf_isfinite(y)) continue;
      double zs = 0.0;
#pragma unroll
      for (int d = 0; d < M_; ++d) zs += Zt[j + mp.p * d] * x[d];
      const double s = Dt[j] + zs;
      w += mng_log_obs(mp.distribution[j], s, y, mp.u[e], mp.phi[j]);
      const double q = (mp.ytilde[e] - s) / mp.Htilde[e];
      w += 0.5 * (q * q);
    }
    return w - scale;
  }
  __device__ void begin_step(int t) {
#pragma unroll
    for (int d = 0; d < M_; ++d) { ah0[d] = mp.alphahat[d + M_ * t]; ah1[d] = mp.alphahat[d + M_ * (t + 1)]; }
#pragma unroll
    for (int e = 0; e < M_ * M_; ++e) {
      Cm[e] = mp.Ct[e + M_ * M_ * (t + 1)];
      Lm[e] = mp.Lt[e + M_ * M_ * (t + 1)];
    }
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    double diff[M_];
#pragma unroll
    for (int j = 0; j < M_; ++j) diff[j] = xa[j] - ah0[j];
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double cd = 0.0, lu = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) cd += Cm[d + M_ * j] * diff[j];
#pragma unroll
      for (int j = 0; j < M_; ++j) lu += Lm[d + M_ * j] * z[j];
      x[d] = ah1[d] + cd + lu;
    }
  }
};

} // namespace pf
