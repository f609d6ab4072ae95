// models_nlg_sde.cuh -- device functor registry for the model classes that the
// reference defines through user-supplied C++ function pointers compiled at
// run time by Rcpp::sourceCpp (ssm_nlg: /root/reference/src/model_ssm_nlg.h:10-21,
// ssm_sde: src/model_ssm_sde.h:8-11).  Host function pointers cannot be called
// from a kernel: each shipped example model is a device
// policy selected by an integer id.
//
//   BSSM_NLG_GROWTH  logistic growth model, R/cpp_example_models.R:336-511
//   BSSM_SDE_GBM     geometric Brownian motion, R/cpp_example_models.R:109-193
#pragma once
#include "pf_device.cuh"
#include "bssm_b200/pf_ekpf.h"

namespace pf {

// ------------------------------------------------------------ ssm_nlg: growth
// state (logit r, p); T_fn: r' unchanged, p <- K p e^{r dT} / (K + p (e^{r dT} - 1)),
// r = e^{r'}/(1 + e^{r'}); Z_fn = p; H = e^{theta0}; R = diag(e^{theta1}, e^{theta2}).
// Bootstrap filter: ssm_nlg::bsf_filter src/model_ssm_nlg.cpp:1066-1146, weights
// dmvnorm(y, Z_fn, H_fn, lwr = true, logd = true) for p = 1 (src/dmvnorm.cpp:9-12).
struct NlgParams {
  int n;
  const double* y;    // p x n with p = 1
  double a1[2];
  double LP1[4];      // psd_chol(P1), src/psd_chol.cpp:5-18
  double dT, K;       // known_params(0), (1)
  double H;           // exp(theta0)
  double R0, R1;      // exp(theta1), exp(theta2)
  double wconst;      // log(2 pi) + 2 log(H)
  double mu, rho;     // ar_exp: theta(0), theta(1); its R is R0, its a1 / sqrt(P1) are a1[0] / LP1[0]
  // EKF particle filter only: prior covariance and the EKF update of (a1, P1) with y_0 (host)
  double P1[4], att1[2], L1[4];
  // is_correction_bsf (src/approx_mcmc.cpp:447-518): one theta per filter.  NLG_DRAW_BLOCK doubles per
  // filter, (a1[2], LP1[4], H, R0, R1, wconst, mu, rho) = what update_model(theta_i) changes; else null
  const double* draw_blk;
  // user functors compiled at run time (user_models.cu, NlgFnBsf below): theta and known_params on the
  // device as the reference hands them to Z_fn / H_fn / T_fn / R_fn (src/model_ssm_nlg.h:10-21), and per
  // filter (or shared: stride 0) the initial distribution (a1 (m), psd_chol(P1) (m x m)) from a1_fn / P1_fn
  const double* theta_u;
  size_t theta_u_stride;
  const double* known_u;
  const double* init_u;
  size_t init_u_stride;
};
constexpr int NLG_DRAW_BLOCK = 12;
// PerFilter policies copy the launch parameters and overwrite the theta-dependent part with their draw's
__device__ __forceinline__ void nlg_load_draw(NlgParams& loc, const NlgParams& p, uint32_t f) {
  loc = p;
  const double* b = p.draw_blk + (size_t)NLG_DRAW_BLOCK * f;
  loc.a1[0] = b[0]; loc.a1[1] = b[1];
  loc.LP1[0] = b[2]; loc.LP1[1] = b[3]; loc.LP1[2] = b[4]; loc.LP1[3] = b[5];
  loc.H = b[6]; loc.R0 = b[7]; loc.R1 = b[8]; loc.wconst = b[9]; loc.mu = b[10]; loc.rho = b[11];
}

template <bool PerFilter>
struct NlgGrowthBsfT : GaussStep<NlgGrowthBsfT<PerFilter>, 2, 2, 2> {
  static constexpr int M = 2, KI = 2, KP = 2;
  typedef NlgParams Params;
  NlgParams loc;
  const NlgParams& mp;
  double y;
  __device__ NlgGrowthBsfT(const NlgParams& p, uint32_t f) : mp(PerFilter ? loc : p) {
    if (PerFilter) nlg_load_draw(loc, p, f);
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; } // normalised densities: no constant afterwards
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) s += mp.LP1[d + 2 * j] * z[j];
      x[d] = mp.a1[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; }
  __device__ double logw(const double* x) const {
    const double q = (y - x[1]) / mp.H;
    return -0.5 * (mp.wconst + q * q);
  }
  __device__ void begin_step(int) {}
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    const double e0 = pf_exp(xa[0]);
    const double r = e0 / (1.0 + e0);
    const double erdt = pf_exp(r * mp.dT);
    const double t1 = mp.K * xa[1] * erdt / (mp.K + xa[1] * (erdt - 1));
    // T_fn(...) + R_fn(...) * uk with R = diag(R0, R1) (the zero entries of R add +-0)
    x[0] = xa[0] + mp.R0 * z[0];
    x[1] = t1 + mp.R1 * z[1];
  }
};
typedef NlgGrowthBsfT<false> NlgGrowthBsf;
typedef NlgGrowthBsfT<true> NlgGrowthBsfDraws;

// ------------------------------------------------------------ ssm_nlg: ar_exp
// AR(1) state observed through its exponential (nlg_ar_exp, R/cpp_example_models.R:194-335):
// alpha_{t+1} = mu (1 - rho) + rho alpha_t + e^{theta2} eta, y_t ~ N(e^{alpha_t}, e^{2 theta3}).
// KIND 1: the nlg_sin_exp model (R/cpp_example_models.R:635-757), alpha_{t+1} = sin(alpha_t) + e^{theta1} eta
// with the same observation equation, a1 = 0, P1 = 1.  KIND 2: nlg_linear_gaussian (:512-634), the random walk
// alpha_{t+1} = alpha_t + eta observed directly, y_t ~ N(alpha_t, e^{2 theta0}) (mu = 0, rho = 1, R = 1 in NlgParams).
enum { NLG1_AR_EXP = 0, NLG1_SIN_EXP = 1, NLG1_LINEAR = 2 };
template <int KIND> __device__ __forceinline__ double nlg1_Z(double a) { return KIND == NLG1_LINEAR ? a : pf_exp(a); }
template <int KIND> __device__ __forceinline__ double nlg1_T(double a, double mu, double rho) {
  return KIND == NLG1_SIN_EXP ? pf_sin(a) : (mu * (1 - rho) + rho * a);
}
template <int KIND, bool PerFilter = false>
struct Nlg1Bsf : GaussStep<Nlg1Bsf<KIND, PerFilter>, 1, 1, 1> {
  static constexpr int M = 1, KI = 1, KP = 1;
  typedef NlgParams Params;
  NlgParams loc;
  const NlgParams& mp;
  double y;
  __device__ Nlg1Bsf(const NlgParams& p, uint32_t f) : mp(PerFilter ? loc : p) {
    if (PerFilter) nlg_load_draw(loc, p, f);
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const { x[0] = mp.a1[0] + mp.LP1[0] * z[0]; }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; }
  __device__ double logw(const double* x) const {
    const double q = (y - nlg1_Z<KIND>(x[0])) / mp.H;
    return -0.5 * (mp.wconst + q * q);
  }
  __device__ void begin_step(int) {}
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    const double tf = nlg1_T<KIND>(xa[0], mp.mu, mp.rho);
    x[0] = tf + mp.R0 * z[0]; // T_fn(...) + R_fn(...) * uk
  }
};
typedef Nlg1Bsf<NLG1_AR_EXP> NlgArExpBsf;
typedef Nlg1Bsf<NLG1_SIN_EXP> NlgSinExpBsf;
typedef Nlg1Bsf<NLG1_LINEAR> NlgLinearBsf;
typedef Nlg1Bsf<NLG1_AR_EXP, true> NlgArExpBsfDraws;
typedef Nlg1Bsf<NLG1_SIN_EXP, true> NlgSinExpBsfDraws;
typedef Nlg1Bsf<NLG1_LINEAR, true> NlgLinearBsfDraws;

// ------------------------------------------- ssm_nlg: functors from the caller
// Bootstrap filter of an ssm_nlg model whose Z_fn, H_fn, T_fn, R_fn come as device functions (p = 1):
// ssm_nlg::bsf_filter src/model_ssm_nlg.cpp:1066-1146 -- alpha_1 = a1 + psd_chol(P1) u, weights
// dmvnorm(y_t, Z_fn(t, alpha), H_fn(t, alpha), lwr = true, logd = true) (src/dmvnorm.cpp:9-12),
// alpha_{t+1} = T_fn(t, alpha) + R_fn(t, alpha) eta.  Fn is the adapter user_models.cu generates.
template <class Fn, int M_, int K_, bool PerFilter>
struct NlgFnBsf : GaussStep<NlgFnBsf<Fn, M_, K_, PerFilter>, M_, M_, K_> {
  static constexpr int M = M_, KI = M_, KP = K_;
  typedef NlgParams Params;
  const NlgParams& mp;
  const double *th, *ini;
  double y;
  int tw, ts;
  __device__ NlgFnBsf(const NlgParams& p, uint32_t f) : mp(p), y(0.0), tw(0), ts(0) {
    th = p.theta_u + (PerFilter ? (size_t)f * p.theta_u_stride : 0);
    ini = p.init_u + (PerFilter ? (size_t)f * p.init_u_stride : 0);
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M_; ++j) s += ini[M_ + d + M_ * j] * z[j];
      x[d] = ini[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; tw = t; }
  __device__ double logw(const double* x) const {
    const double H = Fn::H(tw, x, th, mp.known_u);
    const double q = (y - Fn::Z(tw, x, th, mp.known_u)) / H;
    return -0.5 * ((1.8378770664093453 + 2.0 * pf_log(H)) + q * q); // log(2 pi) + 2 log(H) + q^2
  }
  __device__ void begin_step(int t) { ts = t; }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    double Tx[M_], Rm[M_ * K_];
    Fn::T(ts, xa, th, mp.known_u, Tx);
    Fn::R(ts, xa, th, mp.known_u, Rm);
#pragma unroll
    for (int d = 0; d < M_; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < K_; ++j) s += Rm[d + M_ * j] * z[j];
      x[d] = Tx[d] + s; // T_fn(...) + R_fn(...) * uk
    }
  }
};
// a1_fn / P1_fn of every filter's theta: out = nf x (m + m m), a1 then P1 (column-major)
template <class Fn, int M_>
__global__ void nlg_user_init_kernel(const double* theta, size_t stride, const double* known, uint32_t nf, double* out) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  double* o = out + (size_t)f * (M_ + M_ * M_);
  Fn::a1(theta + (size_t)f * stride, known, o);
  Fn::P1(theta + (size_t)f * stride, known, o + M_);
}

// ------------------------------------------------- ssm_nlg: growth, psi-APF
// ssm_nlg::psi_filter src/model_ssm_nlg.cpp:971-1063 with log_weights :880-935.  The
// Gaussian approximation (EKF start, relinearisation sweeps, Kalman smoother, conditional
// covariances, scales) is host work (include/bssm_b200/nlg_gauss_approx.h); per time point
// the host packs the 32 doubles the weights need into `stepc`:
//   [0] y_t  [1] c_sd = log 2pi + 2 log H_t  [2] H_t  [3] c_var = log 2pi + log HH_t  [4] HH_t
//   [5] D_t  [6..7] Z_t  [8] scales_t
//   transition from t-1 (t >= 1): [10..11] C_{t-1}  [12..15] T_{t-1}
//   [16..19] rooti of RR_{t-1} (approximate)  [20] its constant
//   [21..24] rooti of R R' (true)  [25] its constant
struct NlgPsiParams {
  int n;
  const double* stepc;     // n x 32
  const double* alphahat;  // 2 x (n+1)
  const double* Lt;        // 2 x 2 x (n+1)
  const double* Ct;        // 2 x 2 x (n+1)
  const double* approx_loglik; // per filter
  double dT, K;
  // per-filter strides of stepc / alphahat / (Lt, Ct): 0 when replicates share one
  // approximation, non-zero for post_correct (one approximation per draw)
  size_t fstride_stepc, fstride_a, fstride_m;
};

struct NlgGrowthPsi : GaussStep<NlgGrowthPsi, 2, 2, 2> {
  static constexpr int M = 2, KI = 2, KP = 2;
  typedef NlgPsiParams Params;
  const NlgPsiParams& mp;
  const double *stepc, *alphahat, *Ltp, *Ctp;
  double al0;
  double y, c_sd, H, c_var, HH, D, Z0, Z1, scale;
  double ah0[2], ah1[2], Cm[4], Lm[4];
  double Ca[2], Ta[4], ra[4], ca, rt[4], ct;
  bool has_prev;

  __device__ NlgGrowthPsi(const NlgPsiParams& p, uint32_t f) : mp(p), has_prev(false) {
    al0 = p.approx_loglik[f];
    stepc = p.stepc + (size_t)f * p.fstride_stepc;
    alphahat = p.alphahat + (size_t)f * p.fstride_a;
    Ltp = p.Lt + (size_t)f * p.fstride_m;
    Ctp = p.Ct + (size_t)f * p.fstride_m;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const {
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) s += Ltp[d + 2 * j] * z[j];
      x[d] = alphahat[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(stepc[(size_t)32 * t]); }
  __device__ void begin_weight(int t) {
    const double* c = stepc + (size_t)32 * t;
    y = c[0]; c_sd = c[1]; H = c[2]; c_var = c[3]; HH = c[4]; D = c[5]; Z0 = c[6]; Z1 = c[7]; scale = c[8];
  }
  // observation part of log_weights (:888-896): dmvnorm(y; Z_fn, H, lwr) - dmvnorm(y; D + Z a, HH)
  __device__ double logw(const double* x) const {
    const double q = (y - x[1]) / H;
    const double tru = -0.5 * (c_sd + q * q);
    const double zs = Z0 * x[0] + Z1 * x[1];
    const double mean = D + zs;
    const double app = -0.5 * (c_var + (y - mean) * (y - mean) / HH);
    return tru - app;
  }
  __device__ double finish_logw(double w, double carry) const { return (w - carry) - scale; }
  __device__ void begin_step(int t) {
#pragma unroll
    for (int d = 0; d < 2; ++d) { ah0[d] = alphahat[d + 2 * t]; ah1[d] = alphahat[d + 2 * (t + 1)]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) { Cm[e] = Ctp[e + 4 * (t + 1)]; Lm[e] = Ltp[e + 4 * (t + 1)]; }
    // factors of the transition t -> t+1, used by the weights of column t+1 (if it is weighted)
    has_prev = (t + 1 < mp.n);
    if (has_prev) {
      const double* c = stepc + (size_t)32 * (t + 1);
      Ca[0] = c[10]; Ca[1] = c[11];
#pragma unroll
      for (int e = 0; e < 4; ++e) { Ta[e] = c[12 + e]; ra[e] = c[16 + e]; rt[e] = c[21 + e]; }
      ca = c[20]; ct = c[25];
    }
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    double diff[2];
    diff[0] = xa[0] - ah0[0]; diff[1] = xa[1] - ah0[1];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double cd = 0.0, lu = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) cd += Cm[d + 2 * j] * diff[j];
#pragma unroll
      for (int j = 0; j < 2; ++j) lu += Lm[d + 2 * j] * z[j];
      x[d] = ah1[d] + cd + lu;
    }
  }
  __device__ __forceinline__ double mvn(const double* r, double cst, const double* x, const double* mean) const {
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double z2 = 0.0;
#pragma unroll
      for (int l = 0; l < 2; ++l) z2 += r[i + 2 * l] * (x[l] - mean[l]);
      dot += z2 * z2;
    }
    return cst - 0.5 * dot;
  }
  // dmvnorm(a_t; C + T a_prev, RR) - dmvnorm(a_t; T_fn(a_prev), R R')   (:915-931)
  __device__ double trans_term(const double* xa, const double* x) const {
    if (!has_prev) return 0.0;
    double am[2], tm[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) s += Ta[d + 2 * j] * xa[j];
      am[d] = Ca[d] + s;
    }
    const double e0 = pf_exp(xa[0]);
    const double r = e0 / (1.0 + e0);
    const double erdt = pf_exp(r * mp.dT);
    tm[0] = xa[0];
    tm[1] = mp.K * xa[1] * erdt / (mp.K + xa[1] * (erdt - 1));
    return mvn(ra, ca, x, am) - mvn(rt, ct, x, tm);
  }
};

// ------------------------------------------------- ssm_nlg: ar_exp, psi-APF
// psi-APF for ar_exp: same stepc packing as the growth model with the m = 1 entries
//   [6] Z_t  [10] C_{t-1}  [12] T_{t-1}  [16] RR_{t-1} (approximate variance)  [20] log 2pi + log RR
//   [21] R R' (true variance)  [25] log 2pi + log(R R')  [26..27] (mu, rho) of the true transition
// (for one state dmvnorm takes its scalar branch, src/dmvnorm.cpp:9-16)
template <int KIND>
struct Nlg1Psi : GaussStep<Nlg1Psi<KIND>, 1, 1, 1> {
  static constexpr int M = 1, KI = 1, KP = 1;
  typedef NlgPsiParams Params;
  const NlgPsiParams& mp;
  const double *stepc, *alphahat, *Ltp, *Ctp;
  double al0;
  double y, c_sd, H, c_var, HH, D, Z0, scale;
  double ah0, ah1, Cm, Lm;
  double Ca, Ta, va, ca, vt, ct, mu, rho;
  bool has_prev;

  __device__ Nlg1Psi(const NlgPsiParams& p, uint32_t f) : mp(p), has_prev(false) {
    al0 = p.approx_loglik[f];
    stepc = p.stepc + (size_t)f * p.fstride_stepc;
    alphahat = p.alphahat + (size_t)f * p.fstride_a;
    Ltp = p.Lt + (size_t)f * p.fstride_m;
    Ctp = p.Ct + (size_t)f * p.fstride_m;
  }
  __device__ double loglik0() const { return al0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const { x[0] = alphahat[0] + Ltp[0] * z[0]; }
  __device__ bool observed(int t) const { return pf_isfinite(stepc[(size_t)32 * t]); }
  __device__ void begin_weight(int t) {
    const double* c = stepc + (size_t)32 * t;
    y = c[0]; c_sd = c[1]; H = c[2]; c_var = c[3]; HH = c[4]; D = c[5]; Z0 = c[6]; scale = c[8];
  }
  __device__ double logw(const double* x) const {
    const double q = (y - nlg1_Z<KIND>(x[0])) / H;
    const double tru = -0.5 * (c_sd + q * q);
    const double mean = D + Z0 * x[0];
    const double app = -0.5 * (c_var + (y - mean) * (y - mean) / HH);
    return tru - app;
  }
  __device__ double finish_logw(double w, double carry) const { return (w - carry) - scale; }
  __device__ void begin_step(int t) {
    ah0 = alphahat[t]; ah1 = alphahat[t + 1];
    Cm = Ctp[t + 1]; Lm = Ltp[t + 1];
    has_prev = (t + 1 < mp.n);
    if (has_prev) {
      const double* c = stepc + (size_t)32 * (t + 1);
      Ca = c[10]; Ta = c[12]; va = c[16]; ca = c[20]; vt = c[21]; ct = c[25]; mu = c[26]; rho = c[27];
    }
  }
  __device__ void propagate(const double* xa, const double* z, double* x) const {
    x[0] = ah1 + Cm * (xa[0] - ah0) + Lm * z[0];
  }
  // dmvnorm(a_t; C + T a_prev, RR) - dmvnorm(a_t; T_fn(a_prev), R R'), scalar branch (:915-931)
  __device__ double trans_term(const double* xa, const double* x) const {
    if (!has_prev) return 0.0;
    const double am = Ca + Ta * xa[0];
    const double tm = nlg1_T<KIND>(xa[0], mu, rho); // T_fn of the true model
    const double app = -0.5 * (ca + (x[0] - am) * (x[0] - am) / va);
    const double tru = -0.5 * (ct + (x[0] - tm) * (x[0] - tm) / vt);
    return app - tru;
  }
};
typedef Nlg1Psi<NLG1_AR_EXP> NlgArExpPsi;
typedef Nlg1Psi<NLG1_SIN_EXP> NlgSinExpPsi;
typedef Nlg1Psi<NLG1_LINEAR> NlgLinearPsi;

// ---------------------------------------------- ssm_nlg: EKF particle filter
// ssm_nlg::ekf_filter src/model_ssm_nlg.cpp:1150-1260 (van der Merwe et al.): every particle runs
// its own EKF update from its resampled ancestor -- at = T_fn(a), Pt = R R', (att, Ptt) =
// ekf_update_step(y_{t+1}; at, Pt) -- and draws alpha_{t+1} ~ N(att, Ptt); the weights are
// log g(y | alpha) + log N(alpha; at, Pt) - log N(alpha; att, Ptt).  The per-particle linear algebra
// (include/bssm_b200/pf_ekpf.h, M <= 2) stays in registers.  Model functions per registry id:
template <int MODEL> struct NlgDev;
template <> struct NlgDev<0> { // growth
  static constexpr int M = 2;
  static __device__ __forceinline__ void T_fn(const NlgParams& p, const double* x, double* an) {
    const double e0 = pf_exp(x[0]);
    const double r = e0 / (1.0 + e0);
    const double erdt = pf_exp(r * p.dT);
    an[0] = x[0];
    an[1] = p.K * x[1] * erdt / (p.K + x[1] * (erdt - 1));
  }
  static __device__ __forceinline__ void RR(const NlgParams& p, double* rr) {
    rr[0] = p.R0 * p.R0; rr[1] = 0.0; rr[2] = 0.0; rr[3] = p.R1 * p.R1;
  }
  static __device__ __forceinline__ double Z_fn(const double* x) { return x[1]; }
  static __device__ __forceinline__ void Z_gn(const double*, double* Zg) { Zg[0] = 0.0; Zg[1] = 1.0; }
};
template <> struct NlgDev<1> { // ar_exp
  static constexpr int M = 1;
  static __device__ __forceinline__ void T_fn(const NlgParams& p, const double* x, double* an) {
    an[0] = p.mu * (1 - p.rho) + p.rho * x[0];
  }
  static __device__ __forceinline__ void RR(const NlgParams& p, double* rr) { rr[0] = p.R0 * p.R0; }
  static __device__ __forceinline__ double Z_fn(const double* x) { return pf_exp(x[0]); }
  static __device__ __forceinline__ void Z_gn(const double* x, double* Zg) { Zg[0] = pf_exp(x[0]); }
};
template <> struct NlgDev<2> { // sin_exp
  static constexpr int M = 1;
  static __device__ __forceinline__ void T_fn(const NlgParams&, const double* x, double* an) { an[0] = pf_sin(x[0]); }
  static __device__ __forceinline__ void RR(const NlgParams& p, double* rr) { rr[0] = p.R0 * p.R0; }
  static __device__ __forceinline__ double Z_fn(const double* x) { return pf_exp(x[0]); }
  static __device__ __forceinline__ void Z_gn(const double* x, double* Zg) { Zg[0] = pf_exp(x[0]); }
};

template <> struct NlgDev<3> { // linear_gaussian
  static constexpr int M = 1;
  static __device__ __forceinline__ void T_fn(const NlgParams&, const double* x, double* an) { an[0] = x[0]; }
  static __device__ __forceinline__ void RR(const NlgParams& p, double* rr) { rr[0] = p.R0 * p.R0; }
  static __device__ __forceinline__ double Z_fn(const double* x) { return x[0]; }
  static __device__ __forceinline__ void Z_gn(const double*, double* Zg) { Zg[0] = 1.0; }
};

template <int MODEL>
struct NlgEkpf : GaussStep<NlgEkpf<MODEL>, NlgDev<MODEL>::M, NlgDev<MODEL>::M, NlgDev<MODEL>::M> {
  typedef NlgDev<MODEL> Dev;
  static constexpr int M = Dev::M, KI = Dev::M, KP = Dev::M;
  typedef NlgParams Params;
  const NlgParams& mp;
  double y, ynext, HH;
  bool first, last, upd;
  __device__ NlgEkpf(const NlgParams& p, uint32_t) : mp(p), y(0.0), ynext(0.0), first(false), last(false), upd(false) {
    HH = p.H * p.H;
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ void init(const double* z, double* x) const { // alpha_0 = att1 + L um (:1160-1172)
#pragma unroll
    for (int d = 0; d < M; ++d) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) s += mp.L1[d + M * j] * z[j];
      x[d] = mp.att1[d] + s;
    }
  }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; first = (t == 0); }
  __device__ double logw(const double* x) const {
    const double q = (y - Dev::Z_fn(x)) / mp.H;
    double w = -0.5 * (mp.wconst + q * q);                       // log_obs_density
    if (first) w += pf_dmvnorm_var<M>(x, mp.a1, mp.P1) - pf_dmvnorm_lwr<M>(x, mp.att1, mp.L1); // :1180-1183
    return w;
  }
  __device__ double finish_logw(double w, double carry) const { return w + carry; }
  __device__ void begin_step(int t) {
    last = !(t < mp.n - 1);
    ynext = last ? 0.0 : mp.y[t + 1];
    upd = !last && pf_isfinite(ynext);
  }
  template <int P>
  __device__ __forceinline__ void propagate_all(const StepCtx&, const typename NlgEkpf::template Noise<P>& nz,
                                                const double (&xa)[P][M], double (&x)[P][M], double (&carry)[P]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double at[M], Pt[M * M], att[M], Lc[M * M];
      Dev::T_fn(mp, xa[p], at);
      Dev::RR(mp, Pt);
      if (!last) {
        double Ptt[M * M];
        if (upd) {
          double Zg[M];
          Dev::Z_gn(at, Zg);
          pf_ekf_update<M>(ynext, at, Pt, Zg, Dev::Z_fn(at), HH, att, Ptt);
        } else {
#pragma unroll
          for (int d = 0; d < M; ++d) att[d] = at[d];
#pragma unroll
          for (int e = 0; e < M * M; ++e) Ptt[e] = Pt[e];
        }
        pf_psd_chol<M>(Ptt, Lc);
      } else { // the last column uses the covariance itself as the factor (:1222-1225)
#pragma unroll
        for (int d = 0; d < M; ++d) att[d] = at[d];
#pragma unroll
        for (int e = 0; e < M * M; ++e) Lc[e] = Pt[e];
      }
#pragma unroll
      for (int d = 0; d < M; ++d) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) s += Lc[d + M * j] * nz.z[p * M + j];
        x[p][d] = att[d] + s;
      }
      carry[p] = upd ? pf_dmvnorm_var<M>(x[p], at, Pt) - pf_dmvnorm_lwr<M>(x[p], att, Lc) : 0.0; // :1238-1240
    }
  }
};

// ------------------------------------------------------------------- ssm_sde
// Bootstrap filter ssm_sde::bsf_filter src/model_ssm_sde.cpp:21-82; every particle
// step is 2^L Milstein sub-steps (src/milstein.cpp:18-48) with fresh N(0, sqrt(dt))
// increments: sub-step s of particle i at column t uses normal (flat index i,
// sub = s) of the stream contract in pf_math.h.  The reference takes drift,
// diffusion, its derivative and the observation log-density as compiled function
// pointers (src/model_ssm_sde.h:9-16); here they are functors of a device registry:
//   SdeGbm        dx = mu x dt + sigma_x x dB, y_k ~ N(log x_k, sigma_y^2)
//                 (R/cpp_example_models.R:109-193)
//   SdePoissonOu  dx = rho (nu - x) dt + sigma dB, y_k ~ Poisson(exp(x_k))
//                 (R/cpp_example_models.R:24-108)
struct SdeParams {
  int n;
  const double* y;
  const double* lgy; // lgamma(y_t + 1) per time point (Poisson-OU), host-computed
  double x0;
  int positive;
  int nsub;         // 2^L
  double dt, sqdt;  // 1/2^L, sqrt(dt)
  double th0, th1, th2;
  double logsd;     // GBM: log(theta2)
  double e0, e2;    // Poisson-OU: exp(theta0), exp(theta2)
  // IS-correction (one theta per filter): 6 doubles per filter {th0, th1, th2, logsd, e0, e2}, or null
  const double* theta_f;
  // user functors compiled at run time (user_models.cu): the model's theta vector on the device, as the
  // reference hands `theta` to drift / diffusion / ddiffusion / log_obs_density (src/model_ssm_sde.h:8-11);
  // theta_u_stride = 0: one theta for the launch, else one per filter (PerFilter policies)
  const double* theta_u;
  size_t theta_u_stride;
};

struct SdeGbm {
  static __device__ __forceinline__ double drift(const SdeParams& mp, double x) { return mp.th0 * x; }
  static __device__ __forceinline__ double diffusion(const SdeParams& mp, double x) {
    const double tx = mp.th1 * x;
    return tx > 0.0 ? tx : 0.0;                       // std::max(0.0, theta1 * x)
  }
  static __device__ __forceinline__ double ddiffusion(const SdeParams& mp, double x) {
    return mp.th1 * (x > 0.0 ? 1.0 : 0.0);
  }
  // R::dnorm(y, log(x), theta2, log = TRUE): -(M_LN_SQRT_2PI + 0.5 q^2 + log(sd)) [UNVENDORED: R nmath]
  static __device__ __forceinline__ double log_obs(const SdeParams& mp, double y, double, double x) {
    const double q = (y - pf_log(x)) / mp.th2;
    return -(0.918938533204672741780329736406 + 0.5 * q * q + mp.logsd);
  }
};

struct SdePoissonOu {
  static __device__ __forceinline__ double drift(const SdeParams& mp, double x) { return mp.e0 * (mp.th1 - x); }
  static __device__ __forceinline__ double diffusion(const SdeParams& mp, double) { return mp.e2; }
  static __device__ __forceinline__ double ddiffusion(const SdeParams&, double) { return 0.0; }
  // R::dpois(y, exp(x), log = TRUE) = y x - exp(x) - lgamma(y + 1) [UNVENDORED: R nmath dpois.c
  // evaluates the same quantity through stirlerr / bd0]
  static __device__ __forceinline__ double log_obs(const SdeParams&, double y, double lgy, double x) {
    return (y * x - pf_exp(x)) - lgy;
  }
};

// PerFilter: every filter of the launch has its own theta (is_correction_bsf for ssm_sde,
// src/approx_mcmc.cpp:522-591); otherwise the parameters stay in the constant bank.
template <class Fn, bool PerFilter = false>
struct SdeBsf {
  static constexpr int M = 1;
  typedef SdeParams Params;
  SdeParams loc;
  const SdeParams& mp;
  double y, lgy;
  __device__ SdeBsf(const SdeParams& p, uint32_t f) : mp(PerFilter ? loc : p), y(0.0), lgy(0.0) {
    if (PerFilter) {
      loc = p;
      if (p.theta_f) {
        const double* th = p.theta_f + (size_t)6 * f;
        loc.th0 = th[0]; loc.th1 = th[1]; loc.th2 = th[2]; loc.logsd = th[3]; loc.e0 = th[4]; loc.e2 = th[5];
      }
      loc.theta_u = p.theta_u + (size_t)f * p.theta_u_stride;
    }
  }
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; lgy = mp.lgy ? mp.lgy[t] : 0.0; }
  __device__ double logw(const double* x) const { return Fn::log_obs(mp, y, lgy, x[0]); }
  __device__ void begin_step(int) {}
  __device__ __forceinline__ double milstein_step(double x, double dB) const {
    const double drift = Fn::drift(mp, x);
    const double diff = Fn::diffusion(mp, x);
    const double ddiff = Fn::ddiffusion(mp, x);
    x += drift * mp.dt + diff * dB + 0.5 * diff * ddiff * (dB * dB - mp.dt);
    if (mp.positive) x = fabs(x);
    return x;
  }
  template <int P> struct Noise {};
  template <int P>
  __device__ __forceinline__ void run(const StepCtx& c, double (&x)[P][1]) const {
    const uint32_t pair0 = c.i0 >> 1;
    for (int s = 0; s < mp.nsub; ++s) {
#pragma unroll
      for (int q = 0; q < P / 2; ++q) {
        double z0, z1;
        pf_normal_pair(c.seed, c.filter, c.t, (uint32_t)s, pair0 + q, &z0, &z1);
        x[2 * q][0] = milstein_step(x[2 * q][0], z0 * mp.sqdt);
        x[2 * q + 1][0] = milstein_step(x[2 * q + 1][0], z1 * mp.sqdt);
      }
    }
  }
  template <int P>
  __device__ __forceinline__ void init_all(const StepCtx& c, double (&x)[P][1]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) x[p][0] = mp.x0;
    run<P>(c, x);
  }
  template <int P>
  __device__ __forceinline__ void draw(const StepCtx&, Noise<P>&) const {}
  template <int P>
  __device__ __forceinline__ void propagate_all(const StepCtx& c, const Noise<P>&, const double (&xa)[P][1],
                                                double (&x)[P][1], double (&)[P]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) x[p][0] = xa[p][0];
    run<P>(c, x);
  }
  __device__ __forceinline__ double finish_logw(double w, double) const { return w; }
  __device__ __forceinline__ double step_const() const { return 0.0; }
};
typedef SdeBsf<SdeGbm> SdeGbmBsf;
typedef SdeBsf<SdePoissonOu> SdePoissonOuBsf;
typedef SdeBsf<SdeGbm, true> SdeGbmBsfDraws;
typedef SdeBsf<SdePoissonOu, true> SdePoissonOuBsfDraws;

} // namespace pf
