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
};

struct NlgGrowthBsf : GaussStep<NlgGrowthBsf, 2, 2, 2> {
  static constexpr int M = 2, KI = 2, KP = 2;
  typedef NlgParams Params;
  const NlgParams& mp;
  double y;
  __device__ NlgGrowthBsf(const NlgParams& p, uint32_t) : mp(p) {}
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

// --------------------------------------------------------------- ssm_sde: GBM
// dx = mu x dt + sigma_x x dB, y_k ~ N(log x_k, sigma_y^2).  Bootstrap filter:
// ssm_sde::bsf_filter src/model_ssm_sde.cpp:21-82; every particle step is 2^L
// Milstein sub-steps (src/milstein.cpp:18-48) with fresh N(0, sqrt(dt))
// increments: sub-step s of particle i at column t uses normal (flat index i,
// sub = s) of the stream contract in pf_math.h.
struct SdeParams {
  int n;
  const double* y;
  double x0;
  int positive;
  int nsub;         // 2^L
  double dt, sqdt;  // 1/2^L, sqrt(dt)
  double th0, th1, th2;
  double logsd;     // log(theta2)
};

struct SdeGbmBsf {
  static constexpr int M = 1;
  typedef SdeParams Params;
  const SdeParams& mp;
  double y;
  __device__ SdeGbmBsf(const SdeParams& p, uint32_t) : mp(p) {}
  __device__ double loglik0() const { return 0.0; }
  __device__ double loglik_end() const { return 0.0; }
  __device__ bool observed(int t) const { return pf_isfinite(mp.y[t]); }
  __device__ void begin_weight(int t) { y = mp.y[t]; }
  // R::dnorm(y, log(x), theta2, log = TRUE): -(M_LN_SQRT_2PI + 0.5 q^2 + log(sd)) [UNVENDORED: R nmath]
  __device__ double logw(const double* x) const {
    const double q = (y - pf_log(x[0])) / mp.th2;
    return -(0.918938533204672741780329736406 + 0.5 * q * q + mp.logsd);
  }
  __device__ void begin_step(int) {}
  __device__ __forceinline__ double milstein_step(double x, double dB) const {
    const double drift = mp.th0 * x;
    const double tx = mp.th1 * x;
    const double diff = tx > 0.0 ? tx : 0.0;           // std::max(0.0, theta1 * x)
    const double ddiff = mp.th1 * (x > 0.0 ? 1.0 : 0.0);
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
                                                double (&x)[P][1]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) x[p][0] = xa[p][0];
    run<P>(c, x);
  }
};

} // namespace pf
