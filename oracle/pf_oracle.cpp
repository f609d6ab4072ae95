/* pf_oracle.cpp -- CPU restatement of the helske/bssm particle-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (bssm_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as
 * the checker / the CPU baseline.
 *
 * PARITY: PARTLY PINNED.  The reference itself cannot be built here (no R,
 * Rcpp, RcppArmadillo, sitmo).  Pinned against the reference's
 * OWN golden vectors (tests/test_reference_golden.py, inputs regenerated with a
 * restatement of R's set.seed / rnorm stream, tests/r_rng.py):
 *   - the Gaussian approximation of an svm model -- model construction,
 *     laplace_iter, Kalman fast smoother, Newton-sweep semantics -- to the 15
 *     digits of tests/testthat/test_approx.R:38-53;
 *   - dmvnorm (the density of the ssm_nlg weights), test_basics.R:176-201;
 *   - the EKF and iterated-EKF log-likelihoods of the nlg_sin_exp model (the
 *     start of the ssm_nlg approximation and the EKPF proposal), to
 *     expect_equal's 1.5e-8, test_ekpf.R:64-69;
 *   - the GLM fixed points (Poisson, binomial, negative binomial), the
 *     ssm_mng = 2 x ssm_ung and the two-sweep nlg tests of test_approx.R:60-199;
 *   - the GBM data set of test_sde.R with its log-likelihood assertion, plus
 *     the closed-form likelihood of that model.
 * Also pinned: stratified_sample hand vectors, Philox Random123 vectors,
 * exact Kalman smoothers (numpy) for the general-m / multivariate
 * approximations and the LGSSM particle smoothers, data/poisson_series and
 * data/exchange decoded from the reference's own .rda (tests/test_oracle.py).
 * STILL UNPINNED: the particle loop's random numbers (sitmo + libstdc++
 * <random>) and Armadillo's summation order -- the reference's tests pin no
 * particle-filter output value, so indices / weights / log-likelihoods are
 * defined by this restatement.
 *
 * Every function keeps the reference's loop structure (scalar loops over
 * particles, sequential two-pointer stratified merge, particle-slowest alpha
 * cube m x (n+1) x N, column-major everything) and cites the file:line it
 * follows, relative to /root/reference.
 *
 * Two deliberate, documented substitutions:
 *   RNG   sitmo::prng_engine + std::normal_distribution (src/model_ssm_ung.cpp:
 *         536-546) -> the counter-based Philox stream of include/bssm_b200/
 *         pf_math.h, consumed in the reference's order (particle-major).
 *   libm  std::exp/log/log1p -> pf_exp/pf_log/pf_log1p (<= 4 ulp from glibc).
 * and one switch, `arith`:
 *   ORC_ARITH_REFERENCE  sum of weights = Armadillo's two-accumulator
 *         arrayops::accumulate [UNVENDORED, from memory], cumsum = sequential
 *         binary64 (arma::cumsum) -- verbatim src/stratified_sample.cpp:12.
 *   ORC_ARITH_REPRO      the order-independent fixed-point sum / cumsum of
 *         pf_math.h, which is what the CUDA kernels compute.
 * tests/test_oracle.py checks that both modes give identical ancestor indices
 * on the tested configurations and counts near-ties where they could differ.
 */
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <limits>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "bssm_b200/pf_math.h"
#include "bssm_b200.h"
#include "bssm_b200/ung_gauss_approx.h"
#include "bssm_b200/nlg_gauss_approx.h"
#include "bssm_b200/mng_gauss_approx.h"
#include "bssm_b200/pf_ekpf.h"

enum { ORC_ARITH_REFERENCE = 0, ORC_ARITH_REPRO = 1 };

static const double ORC_INF = std::numeric_limits<double>::infinity();

/* ---------------------------------------------------------------- helpers */

/* arma::accu on a contiguous vector: arrayops::accumulate, two accumulators
 * [UNVENDORED: RcppArmadillo, from memory of arrayops_meat.hpp] */
static double accu_arma(const double* x, size_t n) {
  double acc1 = 0.0, acc2 = 0.0;
  size_t j;
  for (j = 1; j < n; j += 2) {
    acc1 += x[j - 1];
    acc2 += x[j];
  }
  if ((j - 1) < n) acc1 += x[j - 1];
  return acc1 + acc2;
}

static double sum_weights(const double* w, size_t n, int arith) {
  if (arith == ORC_ARITH_REFERENCE) return accu_arma(w, n);
  int64_t hi = 0, lo = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!(w[i] == w[i])) return pf_nan();
    pf_fixsum t = pf_fixsum_term(w[i]);
    hi += t.hi;
    lo += t.lo;
  }
  return pf_fixsum_value(hi, lo);
}

extern "C" {

/* ------------------------------------------------------------ math probes */
void orc_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
  pf_u32x4 o = pf_philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
  for (int i = 0; i < 4; ++i) out[i] = o.v[i];
}
double orc_exp(double x) { return pf_exp(x); }
double orc_log(double x) { return pf_log(x); }
double orc_log1p(double x) { return pf_log1p(x); }
void orc_sincos2pi(double u, double* s, double* c) { pf_sincos2pi(u, s, c); }
void orc_exp_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_exp(x[i]); }
void orc_log_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_log(x[i]); }
void orc_log1p_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_log1p(x[i]); }
void orc_sin_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_sin(x[i]); }
void orc_cos_v(const double* x, double* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = pf_cos(x[i]); }
void orc_sincos2pi_v(const double* u, double* s, double* c, size_t n) {
  for (size_t i = 0; i < n; ++i) pf_sincos2pi(u[i], &s[i], &c[i]);
}
void orc_normals(uint32_t seed, uint32_t filter, uint32_t t, uint32_t sub, uint32_t count, double* z) {
  for (uint32_t f = 0; f < count; ++f) z[f] = pf_normal(seed, filter, t, sub, f);
}
void orc_resample_uniforms(uint32_t seed, uint32_t filter, uint32_t t, uint32_t N, double* r) {
  for (uint32_t j = 0; j < N; ++j) r[j] = pf_resample_uniform(seed, filter, t, j);
}

/* ------------------------------------------------------------------- a1 */
/* src/stratified_sample.cpp:9-28, verbatim: in-place cumsum, p[last] = 1,
 * two-pointer merge with alpha = 1.0/N (reciprocal multiply, :16-18). */
void orc_stratified_sample(double* p, uint32_t n_p, const double* r, uint32_t N, uint32_t* xp) {
  for (uint32_t k = 1; k < n_p; ++k) p[k] = p[k - 1] + p[k]; /* arma::cumsum, sequential */
  p[n_p - 1] = 1;
  unsigned int j = 0;
  double alpha = 1.0 / N;
  for (unsigned int k = 0; k < n_p && j < N; k++) {
    while (j < N && (r[j] + j) * alpha <= p[k]) {
      xp[j] = k;
      j++;
    }
  }
  while (j < N) {
    xp[j] = n_p - 1;
    j++;
  }
}

/* the same merge with the order-independent cumulative weights of pf_math.h:
 * w are UNNORMALISED weights in [0,1]; p_k = C_k * (1/S) (overwrites w). */
void orc_stratified_sample_repro(double* w, uint32_t n_p, const double* r, uint32_t N, uint32_t* xp) {
  int64_t hi = 0, lo = 0;
  std::vector<int64_t> H(n_p), Lo(n_p);
  for (uint32_t k = 0; k < n_p; ++k) {
    pf_fixsum t = pf_fixsum_term(w[k]);
    hi += t.hi; lo += t.lo;
    H[k] = hi; Lo[k] = lo;
  }
  const double invS = 1.0 / pf_fixsum_value(hi, lo);
  for (uint32_t k = 0; k < n_p; ++k) w[k] = pf_fixsum_value(H[k], Lo[k]) * invS;
  w[n_p - 1] = 1;
  unsigned int j = 0;
  double alpha = 1.0 / N;
  for (unsigned int k = 0; k < n_p && j < N; k++) {
    while (j < N && (r[j] + j) * alpha <= w[k]) {
      xp[j] = k;
      j++;
    }
  }
  while (j < N) {
    xp[j] = n_p - 1;
    j++;
  }
}

} /* extern "C" */

/* W holds the reference's normalized_weights in ORC_ARITH_REFERENCE mode and the
 * unnormalised weights (exp(logw - max), or 1 for an unweighted column) in
 * ORC_ARITH_REPRO mode, where the normalisation is applied to the cumulative sums */
static void resample(std::vector<double>& W, const std::vector<double>& r, uint32_t N,
                     uint32_t* xp, int arith) {
  if (arith == ORC_ARITH_REFERENCE) orc_stratified_sample(W.data(), N, r.data(), N, xp);
  else orc_stratified_sample_repro(W.data(), N, r.data(), N, xp);
}

/* ------------------------------------------------------ ssm_ung densities */

struct Ung {
  const bssm_ung_model* M;
  int n, m, k;
  std::vector<double> xbeta;
  /* Gaussian approximation (approx_model of src/model_ssm_ung.h:71) */
  std::vector<double> ay, aH, aHH, scales, mode;
  double approx_loglik = 0.0;
  explicit Ung(const bssm_ung_model* M_) : M(M_), n(M_->n), m(M_->m), k(M_->k) {
    xbeta.assign(n, 0.0);
    if (M->xbeta) for (int t = 0; t < n; ++t) xbeta[t] = M->xbeta[t];
    ay.assign(n, 0.0); aH.assign(n, 0.0); aHH.assign(n, 0.0); scales.assign(n, 0.0);
    mode.assign(n, 0.0);
  }
  const double* Zc(int t) const { return M->Z + (size_t)m * (t * M->Ztv); }
  const double* Tt(int t) const { return M->T + (size_t)m * m * (t * M->Ttv); }
  const double* Rt(int t) const { return M->R + (size_t)m * k * (t * M->Rtv); }
  const double* Cc(int t) const { return M->C + (size_t)m * (t * M->Ctv); }
  double Dt(int t) const { return M->D[t * M->Dtv]; }

  /* signal of particle state a (length m) at time t:
   * src/model_ssm_ung.cpp:395 (svm: first state) and :402-403 (others) */
  double signal(int t, const double* a) const {
    if (M->distribution == 0) return a[0];
    double zs = 0.0;
    const double* Z = Zc(t);
    for (int d = 0; d < m; ++d) zs += Z[d] * a[d];
    return Dt(t) + zs + xbeta[t];
  }
  /* src/model_ssm_ung.cpp:447-493, one particle */
  double log_obs_density(int t, double s) const {
    const double y = M->y[t], phi = M->phi, u = M->u[t];
    switch (M->distribution) {
      case 0: return -0.5 * (s + (y / phi) * (y / phi) * pf_exp(-s));
      case 1: return y * s - u * pf_exp(s);
      case 2: return y * s - u * pf_log1p(pf_exp(s));
      case 3: return y * s - (y + phi) * pf_log(phi + u * pf_exp(s));
      case 4: return -phi * (s + (y * pf_exp(-s) / u));
    }
    return 0.0;
  }
  /* src/model_ssm_ung.cpp:385-439: log g - log g~ (unnormalised), one particle */
  double log_weight(int t, double s) const {
    double q = (ay[t] - s) / aH[t];
    return log_obs_density(t, s) + 0.5 * (q * q);
  }
  /* the same inside the particle loop, where the kernels multiply by the reciprocal 1/H~_t
   * (one division per time point instead of one per particle): a declared 1-ulp deviation
   * from the reference's (approx_y - signal) / H of src/model_ssm_ung.cpp:398 */
  double log_weight_pf(int t, double s) const {
    const double rH = 1.0 / aH[t];
    double q = (ay[t] - s) * rH;
    return log_obs_density(t, s) + 0.5 * (q * q);
  }
};

/* state-independent constants: src/distr_consts.cpp:7-67 via the switch at
 * src/model_ssm_ung.cpp:690-710 (BSF) and :347-370 (psi) */
static double lchoose_(double n, double k) {
  return -std::log(n + 1) - std::lgamma(n - k + 1) - std::lgamma(k + 1) + std::lgamma(n + 2);
}
static double dist_const(const bssm_ung_model* M) {
  double res = 0.0;
  const double phi = M->phi;
  for (int t = 0; t < M->n; ++t) {
    const double y = M->y[t];
    if (!std::isfinite(y)) continue;
    const double u = M->u[t];
    switch (M->distribution) {
      case 0: res += -0.5 * std::log(2.0 * M_PI) - std::log(phi); break;
      case 1: res += y * std::log(u) - std::lgamma(y + 1); break;
      case 2: res += lchoose_(u, y); break;
      case 3: res += lchoose_(y + phi - 1, y) + phi * std::log(phi) + y * std::log(u); break;
      case 4: res += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(y) - phi * std::log(u); break;
    }
  }
  return res;
}

/* lower Cholesky of the positive-diagonal sub-block of P1,
 * src/model_ssm_ung.cpp:614-619 */
static void chol_P1(const double* P1, int m, std::vector<double>& L) {
  L.assign((size_t)m * m, 0.0);
  std::vector<int> nz;
  for (int d = 0; d < m; ++d) if (P1[d + (size_t)m * d] > 0) nz.push_back(d);
  int q = (int)nz.size();
  for (int j = 0; j < q; ++j) {
    for (int i = j; i < q; ++i) {
      double s = P1[nz[i] + (size_t)m * nz[j]];
      for (int l = 0; l < j; ++l) s -= L[nz[i] + (size_t)m * nz[l]] * L[nz[j] + (size_t)m * nz[l]];
      if (i == j) L[nz[i] + (size_t)m * nz[j]] = std::sqrt(s);
      else L[nz[i] + (size_t)m * nz[j]] = s / L[nz[j] + (size_t)m * nz[j]];
    }
  }
}

/* one normalisation step, src/model_ssm_ung.cpp:633-647 (and :672-683):
 * logw (length N) -> weights column (unnormalised exp(logw-max)), W normalised.
 * returns false when the filter must stop (sum_weights > 0 fails). */
static bool normalise(const double* logw, uint32_t N, double* wcol, std::vector<double>& W,
                      double* max_out, double* logmean_out, int arith) {
  double max_weight = logw[0];
  bool has_nan = !(logw[0] == logw[0]);
  for (uint32_t i = 1; i < N; ++i) {
    if (logw[i] > max_weight) max_weight = logw[i];
    if (!(logw[i] == logw[i])) has_nan = true;
  }
  /* a NaN log-weight gives a NaN sum in the reference (exp(NaN)), a non-finite
   * max gives exp(-inf - -inf) = NaN: both fail `sum_weights > 0.0` */
  if (has_nan || !std::isfinite(max_weight)) return false;
  for (uint32_t i = 0; i < N; ++i) wcol[i] = pf_exp_nonpos(logw[i] - max_weight);
  double sum = sum_weights(wcol, N, arith);
  if (!(sum > 0.0)) return false;
  if (arith == ORC_ARITH_REFERENCE) for (uint32_t i = 0; i < N; ++i) W[i] = wcol[i] / sum;
  else for (uint32_t i = 0; i < N; ++i) W[i] = wcol[i];
  *max_out = max_weight;
  *logmean_out = pf_log_normal(sum / N);
  return true;
}

extern "C" {

/* ------------------------------------------------------------------- a2 */
/* ssm_ung::bsf_filter, src/model_ssm_ung.cpp:611-712.
 * alpha m x (n+1) x N, weights N x (n+1), indices N x n; all pre-zeroed by the
 * caller (reference: arma::fill::zeros, src/R_bootstrap_filter.cpp:162-164). */
double orc_ung_bsf_filter(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, uint32_t filter,
                          int arith, double* alpha, double* weights, uint32_t* indices,
                          int32_t* stopped_at) {
  Ung g(M);
  const int n = g.n, m = g.m, k = g.k;
  const size_t N = nsim;
  *stopped_at = -1;
  std::vector<double> L_P1;
  chol_P1(M->P1, m, L_P1);
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };

  for (size_t i = 0; i < N; ++i) { /* :621-627 */
    std::vector<double> um(m);
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L_P1[d + (size_t)m * j] * um[j];
      A(d, 0, i) = M->a1[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) { /* :633-647 */
    for (size_t i = 0; i < N; ++i) logw[i] = g.log_obs_density(0, g.signal(0, &A(0, 0, i)));
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) { /* :648-688 */
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* Tm = g.Tt(t); const double* Rm = g.Rt(t); const double* Cv = g.Cc(t);
    for (size_t i = 0; i < N; ++i) {
      std::vector<double> uk(k);
      for (int j = 0; j < k; ++j) uk[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * k + j));
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += Tm[d + (size_t)m * j] * atmp[j + (size_t)m * i];
        for (int j = 0; j < k; ++j) ru += Rm[d + (size_t)m * j] * uk[j];
        A(d, t + 1, i) = Cv[d] + ta + ru;
      }
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && std::isfinite(M->y[t + 1])) {
      for (size_t i = 0; i < N; ++i) logw[i] = g.log_obs_density(t + 1, g.signal(t + 1, &A(0, t + 1, i)));
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  loglik += dist_const(M); /* :690-710 */
  return loglik;
}

} /* extern "C" */

/* ----------------------------------------------------- ssm_ulg filters */
extern "C" {
/* ssm_ulg::bsf_filter, src/model_ssm_ulg.cpp:557-643.  The model arrives as a bssm_ung_model
 * with distribution 5 and H (standard deviation, 1|n).  Like the reference, mu = D + Z'alpha
 * ignores xbeta and the normalising constant of every weighted step is added on the spot. */
double orc_ulg_bsf_filter(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                          double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  *stopped_at = -1;
  const int n = M->n, m = M->m, k = M->k;
  const size_t N = nsim;
  std::vector<double> L_P1;
  {
    /* psd_chol, src/psd_chol.cpp:5-18 */
    const double eps = std::numeric_limits<double>::epsilon();
    double mx = M->P1[0];
    for (int d = 1; d < m; ++d) mx = std::max(mx, M->P1[d + (size_t)m * d]);
    const double thr = std::max(eps, eps * m * mx);
    std::vector<double> sub((size_t)m * m, 0.0);
    for (int c = 0; c < m; ++c)
      for (int r = 0; r < m; ++r)
        if (M->P1[r + (size_t)m * r] > thr && M->P1[c + (size_t)m * c] > thr) sub[r + (size_t)m * c] = M->P1[r + (size_t)m * c];
    chol_P1(sub.data(), m, L_P1);
  }
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  auto H = [&](int t) { return M->H[t * M->Htv]; };
  auto norm_log_const = [](double sd) { return -0.5 * std::log(2.0 * M_PI) - std::log(sd); };
  for (size_t i = 0; i < N; ++i) {
    std::vector<double> um(m);
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L_P1[d + (size_t)m * j] * um[j];
      A(d, 0, i) = M->a1[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  auto weigh = [&](int t) {
    const double* Z = M->Z + (size_t)m * (t * M->Ztv);
    const double HH = H(t) * H(t);
    for (size_t i = 0; i < N; ++i) {
      double zs = 0.0;
      for (int d = 0; d < m; ++d) zs += Z[d] * A(d, t, i);
      double mu = M->D[t * M->Dtv] + zs;
      logw[i] = -0.5 * ((M->y[t] - mu) * (M->y[t] - mu)) / HH;
    }
  };
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) {
    weigh(0);
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean + norm_log_const(H(0));
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* Tm = M->T + (size_t)m * m * (t * M->Ttv);
    const double* Rm = M->R + (size_t)m * k * (t * M->Rtv);
    const double* Cv = M->C + (size_t)m * (t * M->Ctv);
    for (size_t i = 0; i < N; ++i) {
      std::vector<double> uk(k);
      for (int j = 0; j < k; ++j) uk[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * k + j));
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += Tm[d + (size_t)m * j] * atmp[j + (size_t)m * i];
        for (int j = 0; j < k; ++j) ru += Rm[d + (size_t)m * j] * uk[j];
        A(d, t + 1, i) = Cv[d] + ta + ru;
      }
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && std::isfinite(M->y[t + 1])) {
      weigh(t + 1);
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean + norm_log_const(H(t + 1));
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* ssm_ulg::psi_filter, src/model_ssm_ulg.cpp:646-672: alpha m x (n+1) x N */
void orc_ulg_psi_filter(const bssm_ung_model* M, uint32_t nsim, uint32_t seed, double* alpha) {
  const int n = M->n, m = M->m;
  const size_t mm = (size_t)m * m;
  pfga::UngApprox ga(M); /* smoother_ccov + conditional_cov of the model itself */
  for (int t = 0; t < n; ++t) {
    const double h = M->H[t * M->Htv];
    ga.ay[t] = M->y[t]; ga.aH[t] = h; ga.aHH[t] = h * h;
  }
  ga.smoother_and_conditional();
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  for (size_t i = 0; i < nsim; ++i) {
    for (int t = 0; t <= n; ++t) {
      std::vector<double> um(m);
      for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, 0, t, 0, (uint32_t)(i * m + j));
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        if (t > 0)
          for (int j = 0; j < m; ++j) cd += ga.Ct[d + (size_t)m * j + mm * t] * (A(j, t - 1, i) - ga.alphahat[j + (size_t)m * (t - 1)]);
        for (int j = 0; j < m; ++j) lu += ga.L[d + (size_t)m * j + mm * t] * um[j];
        A(d, t, i) = t > 0 ? ga.alphahat[d + (size_t)m * t] + cd + lu : ga.alphahat[d] + lu;
      }
    }
  }
}
} /* extern "C" */

/* ----------------------------------------- Gaussian approximation, m = 1 */
/* The scalar (m = 1) specialisation of the approx_model (class ssm_ulg) used by
 * the psi-APF.  Z, T, R, C, D may be time varying. */
struct Ulg1 {
  const Ung& g;
  explicit Ulg1(const Ung& g_) : g(g_) {}
  double Z(int t) const { return g.Zc(t)[0]; }
  double T(int t) const { return g.Tt(t)[0]; }
  double RR(int t) const { double r = g.Rt(t)[0]; return r * r; }
  double C(int t) const { return g.Cc(t)[0]; }
  double D(int t) const { return g.Dt(t); }
  double HH(int t) const { return g.aHH[t]; }
  double ytmp(int t) const { return g.M->xbeta ? g.ay[t] - g.xbeta[t] : g.ay[t]; }
  static constexpr double zero_tol = 1e-12; /* src/model_ssm_ulg.h:19 */

  /* ssm_ulg::log_likelihood, src/model_ssm_ulg.cpp:106-138 */
  double log_likelihood() const {
    const int n = g.n;
    double sumH = 0.0, sumR = 0.0;
    for (int t = 0; t < n; ++t) sumH += g.aH[t];
    int nR = g.M->Rtv ? n : 1;
    for (int t = 0; t < nR; ++t) sumR += g.Rt(t)[0];
    if (sumH + sumR < zero_tol) return -ORC_INF;
    double logLik = 0.0;
    double at = g.M->a1[0], Pt = g.M->P1[0];
    const double LOG2PI = std::log(2.0 * M_PI);
    for (int t = 0; t < n; ++t) {
      double F = Z(t) * Pt * Z(t) + HH(t);
      double yt = ytmp(t);
      if (std::isfinite(yt) && F > zero_tol) {
        double v = yt - D(t) - Z(t) * at;
        double K = Pt * Z(t) / F;
        at = C(t) + T(t) * (at + K * v);
        Pt = T(t) * (Pt - K * K * F) * T(t) + RR(t);
        logLik -= 0.5 * (LOG2PI + pf_log(F) + v * v / F);
      } else {
        at = C(t) + T(t) * at;
        Pt = T(t) * Pt * T(t) + RR(t);
      }
    }
    return logLik;
  }

  /* ssm_ulg::fast_smoother, src/model_ssm_ulg.cpp:252-305: returns at (n+1) */
  void fast_smoother(std::vector<double>& at) const {
    const int n = g.n;
    at.assign(n + 1, 0.0);
    std::vector<double> vt(n), Ft(n), Kt(n), rt(n);
    at[0] = g.M->a1[0];
    double Pt = g.M->P1[0];
    for (int t = 0; t < n; ++t) {
      Ft[t] = Z(t) * Pt * Z(t) + HH(t);
      double yt = ytmp(t);
      if (std::isfinite(yt) && Ft[t] > zero_tol) {
        Kt[t] = Pt * Z(t) / Ft[t];
        vt[t] = yt - D(t) - Z(t) * at[t];
        at[t + 1] = C(t) + T(t) * (at[t] + Kt[t] * vt[t]);
        double tmp = 1.0 - Kt[t] * Z(t);
        Pt = T(t) * (tmp * Pt * tmp + Kt[t] * HH(t) * Kt[t]) * T(t) + RR(t);
      } else {
        at[t + 1] = C(t) + T(t) * at[t];
        Pt = T(t) * Pt * T(t) + RR(t);
      }
    }
    rt[n - 1] = 0.0;
    for (int t = n - 1; t > 0; --t) {
      if (std::isfinite(ytmp(t)) && Ft[t] > zero_tol) {
        double L = T(t) * (1.0 - Kt[t] * Z(t));
        rt[t - 1] = Z(t) / Ft[t] * vt[t] + L * rt[t];
      } else {
        rt[t - 1] = T(t) * rt[t];
      }
    }
    if (std::isfinite(g.ay[0]) && Ft[0] > zero_tol) {
      double L = T(0) * (1.0 - Kt[0] * Z(0));
      at[0] = g.M->a1[0] + g.M->P1[0] * (Z(0) / Ft[0] * vt[0] + L * rt[0]);
    } else {
      at[0] = g.M->a1[0] + g.M->P1[0] * T(0) * rt[0];
    }
    for (int t = 0; t < n - 1; ++t) at[t + 1] = C(t) + T(t) * at[t] + RR(t) * rt[t];
  }

  /* ssm_ulg::smoother_ccov, src/model_ssm_ulg.cpp:419-473 */
  void smoother_ccov(std::vector<double>& at, std::vector<double>& Pt, std::vector<double>& ccov) const {
    const int n = g.n;
    at.assign(n + 1, 0.0); Pt.assign(n + 1, 0.0); ccov.assign(n + 1, 0.0);
    std::vector<double> vt(n), Ft(n), Kt(n);
    at[0] = g.M->a1[0];
    Pt[0] = g.M->P1[0];
    for (int t = 0; t < n; ++t) {
      Ft[t] = Z(t) * Pt[t] * Z(t) + HH(t);
      double yt = ytmp(t);
      if (std::isfinite(yt) && Ft[t] > zero_tol) {
        Kt[t] = Pt[t] * Z(t) / Ft[t];
        vt[t] = yt - D(t) - Z(t) * at[t];
        at[t + 1] = C(t) + T(t) * (at[t] + Kt[t] * vt[t]);
        double tmp = 1.0 - Kt[t] * Z(t);
        Pt[t + 1] = T(t) * (tmp * Pt[t] * tmp + Kt[t] * HH(t) * Kt[t]) * T(t) + RR(t);
      } else {
        at[t + 1] = C(t) + T(t) * at[t];
        Pt[t + 1] = T(t) * Pt[t] * T(t) + RR(t);
      }
      ccov[t] = Pt[t + 1];
    }
    double rt = 0.0, Nt = 0.0;
    for (int t = n - 1; t >= 0; --t) {
      if (std::isfinite(ytmp(t)) && Ft[t] > zero_tol) {
        double L = T(t) * (1.0 - Kt[t] * Z(t));
        ccov[t] = Pt[t] * L * (1.0 - Nt * ccov[t]);
        rt = Z(t) / Ft[t] * vt[t] + L * rt;
        Nt = Z(t) * Z(t) / Ft[t] + L * Nt * L;
      } else {
        ccov[t] = Pt[t] * T(t) * (1.0 - Nt * ccov[t]);
        rt = T(t) * rt;
        Nt = T(t) * Nt * T(t);
      }
      at[t] += Pt[t] * rt;
      Pt[t] -= Pt[t] * Nt * Pt[t];
    }
    ccov[n] = 0.0;
  }
};

/* conditional_cov, SVD branch (default use_svd = true, src/conditional_dist.h:6)
 * specialised to p = 1: src/conditional_dist.cpp:8-32.  For a 1x1 matrix
 * svd_econ gives U = 1, s = |V| ; `s > eps*p*s(0)` is false only for s = 0. */
static void conditional_cov_m1(std::vector<double>& Vt, std::vector<double>& Ct) {
  const int n1 = (int)Vt.size();
  const double eps = std::numeric_limits<double>::epsilon();
  for (int t = n1 - 1; t > 0; --t) {
    double s = std::fabs(Vt[t - 1]);
    double tmp = (s > eps * 1 * s) ? Ct[t - 1] * (1.0 / s) : 0.0;
    Vt[t] -= tmp * Ct[t - 1];
    Ct[t] = tmp;
    Vt[t] = std::sqrt(std::fabs(Vt[t]));
  }
  Vt[0] = std::sqrt(std::fabs(Vt[0]));
}

/* ssm_ung::laplace_iter, src/model_ssm_ung.cpp:302-342 */
static void laplace_iter(Ung& g, const std::vector<double>& signal) {
  const bssm_ung_model* M = g.M;
  const double phi = M->phi;
  for (int t = 0; t < g.n; ++t) {
    const double y = M->y[t], u = M->u[t], s = signal[t];
    switch (M->distribution) {
      case 0: {
        double tmp = y;
        if (std::fabs(tmp) < 1e-4) tmp = 1e-4;
        double q = tmp / phi;
        g.aHH[t] = 2.0 * pf_exp(s) / (q * q);
        g.ay[t] = s + 1.0 - 0.5 * g.aHH[t];
      } break;
      case 1: {
        g.aHH[t] = 1.0 / (pf_exp(s) * u);
        g.ay[t] = y * g.aHH[t] + s - 1.0;
      } break;
      case 2: {
        double e = pf_exp(s);
        g.aHH[t] = (1.0 + e) * (1.0 + e) / (u * e);
        g.ay[t] = y * g.aHH[t] + s - 1.0 - e;
      } break;
      case 3: {
        double e = pf_exp(s) * u;
        g.aHH[t] = (phi + e) * (phi + e) / (phi * e * (y + phi));
        g.ay[t] = s + (phi + e) * (y - e) / ((y + phi) * e);
      } break;
      case 4: {
        double e = pf_exp(s) * u;
        g.aHH[t] = e / (y * phi);
        g.ay[t] = s - e / y + 1;
      } break;
    }
    g.aH[t] = std::sqrt(g.aHH[t]);
  }
}

/* ssm_ung::update_scales, src/model_ssm_ung.cpp:241-292 */
static void update_scales(Ung& g) {
  for (int t = 0; t < g.n; ++t) {
    g.scales[t] = 0.0;
    if (std::isfinite(g.M->y[t])) g.scales[t] = g.log_weight(t, g.mode[t]);
  }
}

/* ssm_ung::approximate, src/model_ssm_ung.cpp:87-140 (approx_state < 1, local
 * approximation branch), m = 1.  mode starts from initial_mode (:174). */
static void approximate(Ung& g) {
  const bssm_ung_model* M = g.M;
  const int n = g.n;
  for (int t = 0; t < n; ++t) g.mode[t] = M->initial_mode[t];
  Ulg1 am(g);
  int i = 0;
  double diff = M->conv_tol + 1;
  std::vector<double> at, mode_new(n);
  while (i < M->max_iter && diff > M->conv_tol) {
    i++;
    laplace_iter(g, g.mode);
    am.fast_smoother(at);
    if (M->distribution == 0) {
      for (int t = 0; t < n; ++t) mode_new[t] = at[t];
    } else {
      for (int t = 0; t < n; ++t) mode_new[t] = g.xbeta[t] + g.Dt(t) + g.Zc(t)[0] * at[t];
    }
    std::vector<double> sq(n);
    for (int t = 0; t < n; ++t) sq[t] = (mode_new[t] - g.mode[t]) * (mode_new[t] - g.mode[t]);
    diff = accu_arma(sq.data(), n) / n;
    g.mode = mode_new;
  }
}

/* compute_const_term, src/model_ssm_ung.cpp:347-370 */
static double const_term(const Ung& g) {
  double c = dist_const(g.M);
  int nobs = 0;
  double sumlog = 0.0;
  for (int t = 0; t < g.n; ++t)
    if (std::isfinite(g.M->y[t])) { nobs++; sumlog += pf_log(g.aH[t]); }
  double norm = -0.5 * nobs * std::log(2 * M_PI) - sumlog;
  return c - norm;
}

struct PsiPre {
  std::vector<double> alphahat, L, Ct;
};

/* the head of ssm_ung::psi_filter, src/model_ssm_ung.cpp:514-533.
 * mode_in == NULL: approximate() from initial_mode, approx_loglik computed
 * (:519-528).  mode_in != NULL: approximate_for_is (:143-161): fixed mode, no
 * iterations, approx_loglik = 0. */
static void psi_prepass(Ung& g, const double* mode_in, PsiPre& pre) {
  Ulg1 am(g);
  if (mode_in) {
    for (int t = 0; t < g.n; ++t) g.mode[t] = mode_in[t];
    laplace_iter(g, g.mode);
    update_scales(g);
    g.approx_loglik = 0.0;
  } else {
    approximate(g);
    double gaussian_loglik = am.log_likelihood();
    update_scales(g);
    double ct = const_term(g);
    g.approx_loglik = gaussian_loglik + ct + accu_arma(g.scales.data(), g.n);
  }
  std::vector<double> Vt;
  am.smoother_ccov(pre.alphahat, Vt, pre.Ct);
  conditional_cov_m1(Vt, pre.Ct);
  pre.L = Vt;
}

extern "C" {

/* exposes the pre-pass arrays for parity tests of the device pre-pass kernel */
int orc_ung_psi_prepass(const bssm_ung_model* M, const double* mode_in, double* mode_out,
                        double* ytilde, double* Htilde, double* scales, double* alphahat,
                        double* L, double* Ct, double* approx_loglik) {
  if (M->m != 1) return 1;
  Ung g(M);
  PsiPre pre;
  psi_prepass(g, mode_in, pre);
  const int n = g.n;
  for (int t = 0; t < n; ++t) {
    if (mode_out) mode_out[t] = g.mode[t];
    if (ytilde) ytilde[t] = g.ay[t];
    if (Htilde) Htilde[t] = g.aH[t];
    if (scales) scales[t] = g.scales[t];
  }
  for (int t = 0; t <= n; ++t) {
    if (alphahat) alphahat[t] = pre.alphahat[t];
    if (L) L[t] = pre.L[t];
    if (Ct) Ct[t] = pre.Ct[t];
  }
  if (approx_loglik) *approx_loglik = g.approx_loglik;
  return 0;
}

/* ------------------------------------------------------------------- a3 */
/* ssm_ung::psi_filter, src/model_ssm_ung.cpp:511-609, m = 1. */
double orc_ung_psi_filter(const bssm_ung_model* M, const double* mode_in, uint32_t nsim,
                          uint32_t seed, uint32_t filter, int arith, double* alpha,
                          double* weights, uint32_t* indices, double* approx_loglik,
                          int32_t* stopped_at) {
  *stopped_at = -1;
  if (M->m != 1) return pf_nan();
  Ung g(M);
  PsiPre pre;
  psi_prepass(g, mode_in, pre);
  if (approx_loglik) *approx_loglik = g.approx_loglik;
  const int n = g.n;
  const size_t N = nsim;
  auto A = [&](int t, size_t i) -> double& { return alpha[t + (size_t)(n + 1) * i]; };

  for (size_t i = 0; i < N; ++i) { /* :538-544 */
    double um = pf_normal(seed, filter, 0, 0, (uint32_t)i);
    A(0, i) = pre.alphahat[0] + pre.L[0] * um;
  }
  std::vector<double> W(N), logw(N), r(N), atmp(N);
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) { /* :549-562 */
    for (size_t i = 0; i < N; ++i) { double a = A(0, i); logw[i] = g.log_weight_pf(0, g.signal(0, &a)) - g.scales[0]; }
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + g.approx_loglik + lmean; /* :561, evaluated left to right */
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    loglik = g.approx_loglik;
  }
  for (int t = 0; t < n; ++t) { /* :568-606 */
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i) atmp[i] = A(t, ind[i]);
    for (size_t i = 0; i < N; ++i) {
      double um = pf_normal(seed, filter, t + 1, 0, (uint32_t)i);
      A(t + 1, i) = pre.alphahat[t + 1] + pre.Ct[t + 1] * (atmp[i] - pre.alphahat[t]) + pre.L[t + 1] * um;
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && std::isfinite(M->y[t + 1])) {
      for (size_t i = 0; i < N; ++i) { double a = A(t + 1, i); logw[i] = g.log_weight_pf(t + 1, g.signal(t + 1, &a)) - g.scales[t + 1]; }
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* ssm_ung::psi_filter for any state dimension m (src/model_ssm_ung.cpp:511-609).  The
 * Gaussian approximation is the shared host implementation of
 * include/bssm_b200/ung_gauss_approx.h (for m = 1 it reproduces Ulg1 above bit for bit,
 * tests/test_oracle.py); the particle loop is restated here. */
double orc_ung_psi_filter_m(const bssm_ung_model* M, const double* mode_in, uint32_t nsim, uint32_t seed,
                            uint32_t filter, int arith, double* alpha, double* weights, uint32_t* indices,
                            double* approx_loglik, int32_t* stopped_at, double* alphahat_out, double* L_out,
                            double* Ct_out) {
  *stopped_at = -1;
  pfga::UngApprox ga(M);
  ga.run(mode_in);
  if (approx_loglik) *approx_loglik = ga.approx_loglik;
  const int n = M->n, m = M->m;
  const size_t N = nsim, mm = (size_t)m * m;
  if (alphahat_out) std::copy(ga.alphahat.begin(), ga.alphahat.end(), alphahat_out);
  if (L_out) std::copy(ga.L.begin(), ga.L.end(), L_out);
  if (Ct_out) std::copy(ga.Ct.begin(), ga.Ct.end(), Ct_out);
  Ung g(M);
  g.ay = ga.ay; g.aH = ga.aH; g.scales = ga.scales;
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  for (size_t i = 0; i < N; ++i) { /* :538-544 */
    std::vector<double> um(m);
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += ga.L[d + (size_t)m * j] * um[j];
      A(d, 0, i) = ga.alphahat[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) {
    for (size_t i = 0; i < N; ++i) logw[i] = g.log_weight_pf(0, g.signal(0, &A(0, 0, i))) - g.scales[0];
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + ga.approx_loglik + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    loglik = ga.approx_loglik;
  }
  for (int t = 0; t < n; ++t) { /* :568-606 */
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* ah0 = &ga.alphahat[(size_t)m * t];
    const double* ah1 = &ga.alphahat[(size_t)m * (t + 1)];
    const double* Ctm = &ga.Ct[mm * (t + 1)];
    const double* Lm = &ga.L[mm * (t + 1)];
    for (size_t i = 0; i < N; ++i) {
      std::vector<double> um(m), diff(m);
      for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * m + j));
      for (int j = 0; j < m; ++j) diff[j] = atmp[j + (size_t)m * i] - ah0[j];
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        for (int j = 0; j < m; ++j) cd += Ctm[d + (size_t)m * j] * diff[j];
        for (int j = 0; j < m; ++j) lu += Lm[d + (size_t)m * j] * um[j];
        A(d, t + 1, i) = ah1[d] + cd + lu;
      }
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && std::isfinite(M->y[t + 1])) {
      for (size_t i = 0; i < N; ++i) logw[i] = g.log_weight_pf(t + 1, g.signal(t + 1, &A(0, t + 1, i))) - g.scales[t + 1];
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* importance_sample_ng for the ssm_ung family (src/R_importance_sample.cpp:26-37) and the SPDK branch of
 * ssm_ung::log_likelihood (src/model_ssm_ung.cpp:199-211): ssm_ulg::simulate_states with nsim > 1
 * (src/model_ssm_ulg.cpp:141-217), fast_precomputing_smoother (:359-413) and fast_smoother with the stored
 * gains (:311-357), restated here with the reference's loop structure (trajectory after trajectory, the
 * smoother as three sweeps over vectors); only the Gaussian approximation itself (pseudo-observations,
 * scales, approx_loglik) is the shared ung_gauss_approx.h.  Normals: pf_sim_normal (pf_math.h).
 * alpha is m x (n+1) x nsim; weights = exp(w - max w); loglik[0] = SPDK estimate, loglik[1] = approx_loglik. */
static void psd_chol(const double* x, int m, std::vector<double>& cholx);
namespace {
struct SimSmoother {
  const bssm_ung_model* M;
  const pfga::UngApprox& ga;
  int n, m, k;
  std::vector<double> Ft, Kt, Lt, xbeta;
  static constexpr double zero_tol = 1e-12;
  SimSmoother(const bssm_ung_model* M_, const pfga::UngApprox& ga_) : M(M_), ga(ga_), n(M_->n), m(M_->m), k(M_->k) {
    Ft.assign(n, 0.0); Kt.assign((size_t)m * n, 0.0); Lt.assign((size_t)m * m * n, 0.0);
    xbeta.assign(n, 0.0);
    if (M->xbeta) for (int t = 0; t < n; ++t) xbeta[t] = M->xbeta[t];
  }
  const double* Z(int t) const { return M->Z + (size_t)m * (t * M->Ztv); }
  const double* T(int t) const { return M->T + (size_t)m * m * (t * M->Ttv); }
  const double* R(int t) const { return M->R + (size_t)m * k * (t * M->Rtv); }
  const double* C(int t) const { return M->C + (size_t)m * (t * M->Ctv); }
  double D(int t) const { return M->D[t * M->Dtv]; }
  void matvec(const double* A, const double* x, double* out, bool transpose) const {
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += (transpose ? A[j + m * i] : A[i + m * j]) * x[j];
      out[i] = s;
    }
  }
  /* the smoother of series y (n values): precompute = true fills Ft, Kt, Lt from the covariance recursion
   * (:359-413), false uses them (:311-357).  at is m x (n+1). */
  void smooth(const std::vector<double>& y, bool precompute, std::vector<double>& at) {
    const size_t mm = (size_t)m * m;
    at.assign((size_t)m * (n + 1), 0.0);
    std::vector<double> Pt(M->P1, M->P1 + mm), vt(n, 0.0), ytmp(n), tmp(m), ta(m);
    for (int i = 0; i < m; ++i) at[i] = M->a1[i];
    for (int t = 0; t < n; ++t) ytmp[t] = M->xbeta ? y[t] - xbeta[t] : y[t];
    for (int t = 0; t < n; ++t) {
      double* a = &at[(size_t)m * t];
      if (precompute) {
        /* F = Z' P Z + HH */
        std::vector<double> zp(m);
        for (int j = 0; j < m; ++j) { double s = 0.0; for (int i = 0; i < m; ++i) s += Z(t)[i] * Pt[i + m * j]; zp[j] = s; }
        double f = 0.0;
        for (int i = 0; i < m; ++i) f += zp[i] * Z(t)[i];
        Ft[t] = f + ga.aHH[t];
      }
      const bool used = std::isfinite(ytmp[t]) && Ft[t] > zero_tol;
      if (used) {
        if (precompute) {
          std::vector<double> pz(m);
          matvec(Pt.data(), Z(t), pz.data(), false);
          for (int i = 0; i < m; ++i) Kt[i + (size_t)m * t] = pz[i] / Ft[t];
        }
        double za = 0.0;
        for (int i = 0; i < m; ++i) za += Z(t)[i] * a[i];
        vt[t] = ytmp[t] - D(t) - za;
        for (int i = 0; i < m; ++i) tmp[i] = a[i] + Kt[i + (size_t)m * t] * vt[t];
      } else {
        for (int i = 0; i < m; ++i) tmp[i] = a[i];
      }
      matvec(T(t), tmp.data(), ta.data(), false);
      for (int i = 0; i < m; ++i) at[i + (size_t)m * (t + 1)] = C(t)[i] + ta[i];
      if (precompute) {
        /* Joseph form, :372-376 */
        std::vector<double> X(mm), A(mm), B(mm), J(mm), rr(mm);
        if (used) {
          for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) J[i + m * j] = (i == j ? 1.0 : 0.0) - Kt[i + (size_t)m * t] * Z(t)[j];
          pf_mm(J.data(), Pt.data(), A.data(), m, m, m);
          pf_mmt(A.data(), J.data(), B.data(), m, m, m);
          for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i)
              X[i + m * j] = B[i + m * j] + Kt[i + (size_t)m * t] * ga.aHH[t] * Kt[j + (size_t)m * t];
        } else {
          X = Pt;
        }
        pf_mm(T(t), X.data(), A.data(), m, m, m);
        pf_mmt(A.data(), T(t), B.data(), m, m, m);
        pf_mmt(R(t), R(t), rr.data(), m, k, m);
        for (size_t e = 0; e < mm; ++e) Pt[e] = B[e] + rr[e];
        pf_symmatu(Pt.data(), m);
      }
    }
    std::vector<double> rt((size_t)m * n, 0.0), lr(m), w(m), pw(m), J(mm);
    auto Lmat = [&](int t, double* Lm) {
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) J[i + m * j] = (i == j ? 1.0 : 0.0) - Kt[i + (size_t)m * t] * Z(t)[j];
      pf_mm(T(t), J.data(), Lm, m, m, m);
    };
    for (int t = n - 1; t > 0; --t) {
      if (std::isfinite(ytmp[t]) && Ft[t] > zero_tol) {
        if (precompute) Lmat(t, &Lt[mm * t]);
        matvec(&Lt[mm * t], &rt[(size_t)m * t], lr.data(), true);
        for (int i = 0; i < m; ++i) rt[i + (size_t)m * (t - 1)] = Z(t)[i] / Ft[t] * vt[t] + lr[i];
      } else {
        matvec(T(t), &rt[(size_t)m * t], &rt[(size_t)m * (t - 1)], true);
      }
    }
    if (std::isfinite(ytmp[0]) && Ft[0] > zero_tol) {
      std::vector<double> L0(mm);
      Lmat(0, L0.data());
      matvec(L0.data(), &rt[0], lr.data(), true);
      for (int i = 0; i < m; ++i) w[i] = Z(0)[i] / Ft[0] * vt[0] + lr[i];
      matvec(M->P1, w.data(), pw.data(), false);
    } else if (m == 1) {
      pw[0] = M->P1[0] * T(0)[0] * rt[0]; /* 1 x 1 operands associate from the left */
    } else {
      matvec(T(0), &rt[0], w.data(), true);
      matvec(M->P1, w.data(), pw.data(), false);
    }
    for (int i = 0; i < m; ++i) at[i] = M->a1[i] + pw[i];
    std::vector<double> rr(mm), rrr(m);
    for (int t = 0; t < n - 1; ++t) {
      matvec(T(t), &at[(size_t)m * t], ta.data(), false);
      pf_mmt(R(t), R(t), rr.data(), m, k, m);
      matvec(rr.data(), &rt[(size_t)m * t], rrr.data(), false);
      for (int i = 0; i < m; ++i) at[i + (size_t)m * (t + 1)] = C(t)[i] + ta[i] + rrr[i];
    }
  }
};
} // namespace

/* simulate_states (nsim > 1) + importance_weights for one approximated model: alpha m x (n+1) x nsim, w = the
 * sums over t of log_weights(t, alpha_i); `filter` = the random stream */
static void spdk_draws(const bssm_ung_model* M, const pfga::UngApprox& ga, uint32_t nsim, int use_antithetic,
                       uint32_t seed, uint32_t filter, double* alpha, std::vector<double>& w,
                       std::vector<double>& alphahat) {
  const int n = M->n, m = M->m, k = M->k;
  Ung g(M);
  g.ay = ga.ay; g.aH = ga.aH; g.scales = ga.scales;
  SimSmoother S(M, ga);
  std::vector<double> L_P1;
  S.smooth(ga.ay, true, alphahat);
  psd_chol(M->P1, m, L_P1);
  const size_t sl = (size_t)m * (n + 1);
  const uint32_t nsim2 = use_antithetic ? nsim / 2 : nsim;
  std::vector<double> aplus(sl), yplus(n), sm, um(m), uk(k);
  const uint32_t K1 = (uint32_t)k + 1;
  auto simulate = [&](uint32_t i) { /* :165-182 */
    for (int j = 0; j < m; ++j) um[j] = pf_sim_normal(seed, filter, 0u, i * (uint32_t)m + j);
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L_P1[d + (size_t)m * j] * um[j];
      aplus[d] = M->a1[d] + s;
    }
    for (int t = 0; t < n; ++t) {
      const double* a = &aplus[(size_t)m * t];
      yplus[t] = M->y[t];
      if (std::isfinite(M->y[t])) {
        double za = 0.0;
        for (int d = 0; d < m; ++d) za += S.Z(t)[d] * a[d];
        yplus[t] = S.xbeta[t] + S.D(t) + za + ga.aH[t] * pf_sim_normal(seed, filter, (uint32_t)(t + 1), i * K1);
      }
      for (int j = 0; j < k; ++j) uk[j] = pf_sim_normal(seed, filter, (uint32_t)(t + 1), i * K1 + 1 + j);
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += S.T(t)[d + m * j] * a[j];
        for (int j = 0; j < k; ++j) ru += S.R(t)[d + m * j] * uk[j];
        aplus[d + (size_t)m * (t + 1)] = S.C(t)[d] + ta + ru;
      }
    }
    S.smooth(yplus, false, sm);
  };
  for (uint32_t i = 0; i < nsim2; ++i) {
    simulate(i);
    double* ai = alpha + sl * i;
    for (size_t e = 0; e < sl; ++e) ai[e] = -sm[e] + aplus[e]; /* :184 */
    if (use_antithetic) {
      double* aj = alpha + sl * (i + nsim2);
      for (size_t e = 0; e < sl; ++e) aj[e] = alphahat[e] - ai[e]; /* :186 */
    }
    for (size_t e = 0; e < sl; ++e) ai[e] += alphahat[e]; /* :188 */
  }
  if (use_antithetic && 2 * nsim2 < nsim) { /* :190-217 */
    simulate(nsim2);
    double* ai = alpha + sl * (nsim - 1);
    for (size_t e = 0; e < sl; ++e) ai[e] = alphahat[e] - sm[e] + aplus[e];
  }
  /* importance_weights (src/model_ssm_ung.cpp:372-378) */
  w.assign(nsim, 0.0);
  for (int t = 0; t < n; ++t) {
    if (!std::isfinite(M->y[t])) continue;
    for (uint32_t i = 0; i < nsim; ++i) w[i] += g.log_weight_pf(t, g.signal(t, alpha + sl * i + (size_t)m * t));
  }
}

extern "C" int orc_importance_sample_ung(const bssm_ung_model* M, uint32_t nsim, int use_antithetic, uint32_t seed,
                                         double* alpha, double* weights, double* loglik, double* alphahat_out,
                                         double* ytilde_out, double* Htilde_out) {
  if (nsim < 2) return 1;
  const int n = M->n;
  pfga::UngApprox ga(M);
  ga.run_head(nullptr);
  std::vector<double> alphahat, w;
  spdk_draws(M, ga, nsim, use_antithetic, seed, 0u, alpha, w, alphahat);
  if (alphahat_out) std::copy(alphahat.begin(), alphahat.end(), alphahat_out);
  if (ytilde_out) std::copy(ga.ay.begin(), ga.ay.end(), ytilde_out);
  if (Htilde_out) std::copy(ga.aH.begin(), ga.aH.end(), Htilde_out);
  const double ssum = pfga::accu2(ga.scales.data(), (size_t)n);
  double maxw = -ORC_INF;
  for (uint32_t i = 0; i < nsim; ++i) { w[i] -= ssum; if (w[i] > maxw) maxw = w[i]; }
  for (uint32_t i = 0; i < nsim; ++i) w[i] = pf_exp(w[i] - maxw);
  if (weights) std::copy(w.begin(), w.end(), weights);
  if (loglik) {
    loglik[0] = ga.approx_loglik + pf_log(pfga::accu2(w.data(), nsim) / nsim) + maxw;
    loglik[1] = ga.approx_loglik;
  }
  return 0;
}
/* the normals of the simulation smoother, for the numpy pin of tests/test_oracle.py */
extern "C" double orc_sim_normal(uint32_t seed, uint32_t t, uint32_t f) { return pf_sim_normal(seed, 0u, t, f); }

/* ------------------------------------------------------------------ a10 */
/* filter_smoother, src/filter_smoother.cpp:5-18: alpha m x n1 x N in place */
void orc_filter_smoother(double* alpha, const uint32_t* indices, int m, int n1, uint32_t N) {
  std::vector<uint32_t> b(N), b2(N);
  for (uint32_t i = 0; i < N; ++i) b[i] = i;
  std::vector<double> alphatmp((size_t)m * N);
  for (int t = n1 - 2; t >= 0; t--) {
    for (uint32_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) alphatmp[d + (size_t)m * i] = alpha[d + (size_t)m * (t + (size_t)n1 * i)];
    const uint32_t* col = indices + (size_t)N * t;
    for (uint32_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) alpha[d + (size_t)m * (t + (size_t)n1 * i)] = alphatmp[d + (size_t)m * col[b[i]]];
    for (uint32_t i = 0; i < N; ++i) b2[i] = col[b[i]];
    b.swap(b2);
  }
}

/* ------------------------------------------------------------------ a11 */
/* summary, src/summary.cpp:4-17: sequential Welford over particles */
void orc_summary(const double* x, int m, int n1, uint32_t N, double* mean_x, double* cov_x) {
  std::fill(mean_x, mean_x + (size_t)m * n1, 0.0);
  std::fill(cov_x, cov_x + (size_t)m * m * n1, 0.0);
  std::vector<double> diff((size_t)m * n1);
  for (uint32_t i = 0; i < N; ++i) {
    const double* xi = x + (size_t)m * n1 * i;
    for (size_t e = 0; e < (size_t)m * n1; ++e) {
      diff[e] = xi[e] - mean_x[e];
      mean_x[e] += diff[e] / (i + 1);
    }
    for (int t = 0; t < n1; ++t)
      for (int c = 0; c < m; ++c)
        for (int r = 0; r < m; ++r)
          cov_x[r + (size_t)m * (c + (size_t)m * t)] +=
              diff[r + (size_t)m * t] * (xi[c + (size_t)m * t] - mean_x[c + (size_t)m * t]);
  }
  for (size_t e = 0; e < (size_t)m * m * n1; ++e) cov_x[e] /= N;
}

/* filter_summary, src/summary.cpp:42-79.  weights is taken by value in the
 * reference (normalised copy); here `weights` is read-only. at m x n1,
 * att m x (n1-1), Pt m x m x n1, Ptt m x m x (n1-1). */
void orc_filter_summary(const double* alpha, int m, int n1, uint32_t N, const double* weights,
                        double* at, double* att, double* Pt, double* Ptt) {
  const int n = n1 - 1;
  std::fill(at, at + (size_t)m * n1, 0.0);
  std::fill(att, att + (size_t)m * n, 0.0);
  std::vector<double> w((size_t)N * n1);
  for (int t = 0; t < n1; ++t) {
    double s = accu_arma(weights + (size_t)N * t, N);
    for (uint32_t i = 0; i < N; ++i) w[i + (size_t)N * t] = weights[i + (size_t)N * t] / s;
  }
  auto A = [&](int d, int t, uint32_t i) { return alpha[d + (size_t)m * (t + (size_t)n1 * i)]; };
  for (int t = 0; t < n; ++t)
    for (uint32_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) {
        att[d + (size_t)m * t] += A(d, t, i) * w[i + (size_t)N * t];
        at[d + (size_t)m * t] += A(d, t, i);
      }
  for (uint32_t i = 0; i < N; ++i)
    for (int d = 0; d < m; ++d) at[d + (size_t)m * n] += A(d, n, i);
  for (size_t e = 0; e < (size_t)m * n1; ++e) at[e] /= N;
  std::fill(Pt, Pt + (size_t)m * m * n1, 0.0);
  std::fill(Ptt, Ptt + (size_t)m * m * n, 0.0);
  for (int t = 0; t < n1; ++t)
    for (uint32_t i = 0; i < N; ++i)
      for (int c = 0; c < m; ++c)
        for (int r = 0; r < m; ++r) {
          Pt[r + (size_t)m * (c + (size_t)m * t)] +=
              (A(r, t, i) - at[r + (size_t)m * t]) * (A(c, t, i) - at[c + (size_t)m * t]);
          if (t < n)
            Ptt[r + (size_t)m * (c + (size_t)m * t)] += w[i + (size_t)N * t] *
                (A(r, t, i) - att[r + (size_t)m * t]) * (A(c, t, i) - att[c + (size_t)m * t]);
        }
  for (size_t e = 0; e < (size_t)m * m * n1; ++e) Pt[e] /= N;
}

/* CPU baseline driver: `n_filters` independent replicate psi-APF / BSF filters,
 * OpenMP-parallel over filters with schedule(static) -- the reference's only
 * parallel pattern (over independent filters, src/approx_mcmc.cpp:364-369);
 * each filter is the single-threaded loop above with per-filter heap
 * allocation of alpha / weights / indices (src/approx_mcmc.cpp:379-381). */
void orc_ung_filter_batch(const bssm_ung_model* M, int method, const double* mode_in, uint32_t nsim,
                          uint32_t seed, uint32_t first_filter, uint32_t n_filters, int arith,
                          int threads, double* logliks) {
  const size_t n = M->n, m = M->m, N = nsim;
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t f = 0; f < (int64_t)n_filters; ++f) {
    std::vector<double> alpha(m * (n + 1) * N, 0.0), weights(N * (n + 1), 0.0);
    std::vector<uint32_t> indices(N * n, 0u);
    int32_t st;
    double al;
    if (method == 1)
      logliks[f] = orc_ung_psi_filter(M, mode_in, nsim, seed, first_filter + (uint32_t)f, arith, alpha.data(),
                                      weights.data(), indices.data(), &al, &st);
    else
      logliks[f] = orc_ung_bsf_filter(M, nsim, seed, first_filter + (uint32_t)f, arith, alpha.data(),
                                      weights.data(), indices.data(), &st);
  }
}

/* ------------------------------------------------------------------ a12 */
/* ssm_mng::log_obs_density, src/model_ssm_mng.cpp:407-442, one particle.  y, u are n x p. */
static double mng_log_obs(const bssm_mng_model* M, int t, const double* alpha) {
  const int n = M->n, m = M->m, p = M->p;
  const double* Z = M->Z + (size_t)p * m * (t * M->Ztv);
  const double* D = M->D + (size_t)p * (t * M->Dtv);
  double w = 0.0;
  for (int j = 0; j < p; ++j) {
    const double y = M->y[t + (size_t)n * j];
    if (!std::isfinite(y)) continue;
    double zs = 0.0;
    for (int d = 0; d < m; ++d) zs += Z[j + (size_t)p * d] * alpha[d];
    const double s = D[j] + zs; /* simsignal = D + Z alpha */
    const double u = M->u[t + (size_t)n * j], phi = M->phi[j];
    switch (M->distribution[j]) {
      case 1: w += y * s - u * pf_exp(s); break;
      case 2: w += y * s - u * pf_log1p(pf_exp(s)); break;
      case 3: w += y * s - (y + phi) * pf_log(phi + u * pf_exp(s)); break;
      case 4: w += -phi * (s + (y * pf_exp(-s) / u)); break;
      case 5: { double q = (y - s) / phi; w += -0.5 * (q * q); } break;
    }
  }
  return w;
}

/* ssm_mng::bsf_filter, src/model_ssm_mng.cpp:562-668 */
double orc_mng_bsf_filter(const bssm_mng_model* M, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                          double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  *stopped_at = -1;
  const int n = M->n, m = M->m, k = M->k, p = M->p;
  const size_t N = nsim;
  std::vector<double> L_P1;
  chol_P1(M->P1, m, L_P1);
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  auto any_obs = [&](int t) {
    int na = 0;
    for (int j = 0; j < p; ++j) if (!std::isfinite(M->y[t + (size_t)n * j])) na++;
    return na < p;
  };
  for (size_t i = 0; i < N; ++i) {
    std::vector<double> um(m);
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L_P1[d + (size_t)m * j] * um[j];
      A(d, 0, i) = M->a1[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  double loglik = 0.0;
  if (any_obs(0)) {
    for (size_t i = 0; i < N; ++i) logw[i] = mng_log_obs(M, 0, &A(0, 0, i));
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* Tm = M->T + (size_t)m * m * (t * M->Ttv);
    const double* Rm = M->R + (size_t)m * k * (t * M->Rtv);
    const double* Cv = M->C + (size_t)m * (t * M->Ctv);
    for (size_t i = 0; i < N; ++i) {
      std::vector<double> uk(k);
      for (int j = 0; j < k; ++j) uk[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * k + j));
      for (int d = 0; d < m; ++d) {
        double ta = 0.0, ru = 0.0;
        for (int j = 0; j < m; ++j) ta += Tm[d + (size_t)m * j] * atmp[j + (size_t)m * i];
        for (int j = 0; j < k; ++j) ru += Rm[d + (size_t)m * j] * uk[j];
        A(d, t + 1, i) = Cv[d] + ta + ru;
      }
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && any_obs(t + 1)) {
      for (size_t i = 0; i < N; ++i) logw[i] = mng_log_obs(M, t + 1, &A(0, t + 1, i));
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  /* constants added term by term, :636-666 */
  for (int i = 0; i < p; ++i) {
    const double phi = M->phi[i];
    for (int t = 0; t < n; ++t) {
      const double y = M->y[t + (size_t)n * i], u = M->u[t + (size_t)n * i];
      if (!std::isfinite(y)) continue;
      switch (M->distribution[i]) {
        case 1: loglik += -std::lgamma(y + 1) + y * std::log(u); break;
        case 2: loglik += lchoose_(u, y); break;
        case 3: loglik += lchoose_(y + phi - 1, y) + phi * std::log(phi) + y * std::log(u); break;
        case 4: loglik += phi * std::log(phi) - std::lgamma(phi) + (phi - 1) * std::log(y) - phi * std::log(u); break;
        case 5: break; /* added below as count * norm_log_const(phi) */
      }
    }
    if (M->distribution[i] == 5) {
      int cnt = 0;
      for (int t = 0; t < n; ++t) if (std::isfinite(M->y[t + (size_t)n * i])) cnt++;
      loglik += cnt * (-0.5 * std::log(2.0 * M_PI) - std::log(phi));
    }
  }
  return loglik;
}

/* ssm_mng::log_weights, src/model_ssm_mng.cpp:357-393, one particle: per observed
 * series log g(y_j | s_j) + 0.5 ((ytilde_j - s_j) / Htilde_j)^2 */
static double mng_log_weight(const bssm_mng_model* M, const pfga::MngApprox& ga, int t, const double* alpha) {
  const int n = M->n, m = M->m, p = M->p;
  const double* Z = M->Z + (size_t)p * m * (t * M->Ztv);
  const double* D = M->D + (size_t)p * (t * M->Dtv);
  double w = 0.0;
  for (int j = 0; j < p; ++j) {
    const double y = M->y[t + (size_t)n * j];
    if (!std::isfinite(y)) continue;
    double zs = 0.0;
    for (int d = 0; d < m; ++d) zs += Z[j + (size_t)p * d] * alpha[d];
    const double s = D[j] + zs;
    const double u = M->u[t + (size_t)n * j], phi = M->phi[j];
    switch (M->distribution[j]) {
      case 1: w += y * s - u * pf_exp(s); break;
      case 2: w += y * s - u * pf_log1p(pf_exp(s)); break;
      case 3: w += y * s - (y + phi) * pf_log(phi + u * pf_exp(s)); break;
      case 4: w -= phi * (s + (y * pf_exp(-s) / u)); break;
      case 5: { double q = (y - s) / phi; w -= 0.5 * (q * q); } break;
    }
    const double q = (ga.ay[j + (size_t)p * t] - s) / ga.aH[j + (size_t)p * t];
    w += 0.5 * (q * q);
  }
  return w;
}

/* ssm_mng::psi_filter, src/model_ssm_mng.cpp:461-559.  mode_in (p x n) selects
 * approximate_for_is (:121-137); NULL runs approximate() from initial_mode. */
double orc_mng_psi_filter(const bssm_mng_model* M, const double* mode_in, uint32_t nsim, uint32_t seed,
                          uint32_t filter, int arith, double* alpha, double* weights, uint32_t* indices,
                          double* approx_loglik, int32_t* stopped_at) {
  *stopped_at = -1;
  pfga::MngApprox ga(M);
  ga.run(mode_in);
  if (approx_loglik) *approx_loglik = ga.approx_loglik;
  const int n = M->n, m = M->m, p = M->p;
  const size_t N = nsim, mm = (size_t)m * m;
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  auto any_obs = [&](int t) {
    int na = 0;
    for (int j = 0; j < p; ++j) if (!std::isfinite(M->y[t + (size_t)n * j])) na++;
    return na < p;
  };
  for (size_t i = 0; i < N; ++i) { /* :490-496 */
    std::vector<double> um(m);
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += ga.L[d + (size_t)m * j] * um[j];
      A(d, 0, i) = ga.alphahat[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  double loglik = 0.0;
  if (any_obs(0)) { /* :502-514 */
    for (size_t i = 0; i < N; ++i) logw[i] = mng_log_weight(M, ga, 0, &A(0, 0, i)) - ga.scales[0];
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + ga.approx_loglik + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    loglik = ga.approx_loglik;
  }
  for (int t = 0; t < n; ++t) { /* :521-556 */
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* ah0 = &ga.alphahat[(size_t)m * t];
    const double* ah1 = &ga.alphahat[(size_t)m * (t + 1)];
    const double* Ctm = &ga.Ct[mm * (t + 1)];
    const double* Lm = &ga.L[mm * (t + 1)];
    for (size_t i = 0; i < N; ++i) {
      std::vector<double> um(m), diff(m);
      for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * m + j));
      for (int j = 0; j < m; ++j) diff[j] = atmp[j + (size_t)m * i] - ah0[j];
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        for (int j = 0; j < m; ++j) cd += Ctm[d + (size_t)m * j] * diff[j];
        for (int j = 0; j < m; ++j) lu += Lm[d + (size_t)m * j] * um[j];
        A(d, t + 1, i) = ah1[d] + cd + lu;
      }
    }
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && any_obs(t + 1)) {
      for (size_t i = 0; i < N; ++i) logw[i] = mng_log_weight(M, ga, t + 1, &A(0, t + 1, i)) - ga.scales[t + 1];
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* the Gaussian approximation alone (checked against an independent numpy Kalman
 * smoother in tests/test_oracle.py): alphahat m x (n+1), mode p x n */
double orc_mng_approx(const bssm_mng_model* M, double* alphahat, double* mode, double* scales) {
  pfga::MngApprox ga(M);
  ga.run(nullptr);
  std::copy(ga.alphahat.begin(), ga.alphahat.end(), alphahat);
  std::copy(ga.mode.begin(), ga.mode.end(), mode);
  std::copy(ga.scales.begin(), ga.scales.end(), scales);
  return ga.approx_loglik;
}

/* ------------------------------------------------------------------ a13 */
/* psd_chol, src/psd_chol.cpp:5-18 */
static void psd_chol(const double* x, int m, std::vector<double>& cholx) {
  const double eps = std::numeric_limits<double>::epsilon();
  double mx = x[0];
  for (int d = 1; d < m; ++d) mx = std::max(mx, x[d + (size_t)m * d]);
  const double thr = std::max(eps, eps * m * mx);
  std::vector<int> nz;
  for (int d = 0; d < m; ++d) if (x[d + (size_t)m * d] > thr) nz.push_back(d);
  cholx.assign((size_t)m * m, 0.0);
  int q = (int)nz.size();
  for (int j = 0; j < q; ++j)
    for (int i = j; i < q; ++i) {
      double s = x[nz[i] + (size_t)m * nz[j]];
      for (int l = 0; l < j; ++l) s -= cholx[nz[i] + (size_t)m * nz[l]] * cholx[nz[j] + (size_t)m * nz[l]];
      if (i == j) cholx[nz[i] + (size_t)m * nz[j]] = std::sqrt(s);
      else cholx[nz[i] + (size_t)m * nz[j]] = s / cholx[nz[j] + (size_t)m * nz[j]];
    }
}

/* the growth model's user functions, R/cpp_example_models.R:350-428 */
struct Growth {
  const double* theta; const double* known;
  void a1_fn(double* a1) const { a1[0] = known[2]; a1[1] = known[3]; }
  void P1_fn(double* P1) const { P1[0] = known[4]; P1[1] = 0; P1[2] = 0; P1[3] = known[5]; }
  double H_fn() const { return pf_exp(theta[0]); }
  void R_fn(double* R) const { R[0] = pf_exp(theta[1]); R[1] = 0; R[2] = 0; R[3] = pf_exp(theta[2]); }
  double Z_fn(const double* alpha) const { return alpha[1]; }
  void T_fn(const double* alpha, double* an) const {
    double dT = known[0], K = known[1];
    an[0] = alpha[0];
    double r = pf_exp(alpha[0]) / (1.0 + pf_exp(alpha[0]));
    an[1] = K * alpha[1] * pf_exp(r * dT) / (K + alpha[1] * (pf_exp(r * dT) - 1));
  }
};

/* dmvnorm(x, mean, sigma, lwr = true, logd = true) for one series, src/dmvnorm.cpp:9-12 */
static double dmvnorm1(double x, double mean, double sigma) {
  double q = (x - mean) / sigma;
  return -0.5 * (std::log(2.0 * M_PI) + 2.0 * pf_log(sigma) + q * q);
}

/* nlg_ar_exp user functions, R/cpp_example_models.R:200-262: theta = (mu, rho, log sigma, log H) */
struct ArExp {
  const double* theta;
  int kind; /* model_id: 1 = nlg_ar_exp;
             * 2 = nlg_sin_exp, R/cpp_example_models.R:640-690: a1 = 0, P1 = 1, theta = (log H, log R), T_fn = sin;
             * 3 = nlg_linear_gaussian, :519-580: a1 = 0, P1 = 1, theta = (log H), R = 1, Z_fn = T_fn = identity */
  double a1_fn() const { return kind != 1 ? 0.0 : theta[0]; }
  double P1_fn() const {
    if (kind != 1) return 1.0;
    double e = pf_exp(theta[2]);
    return e * e / (1 - theta[1] * theta[1]); /* pow(exp(.), 2) */
  }
  double H_fn() const { return kind != 1 ? pf_exp(theta[0]) : pf_exp(theta[3]); }
  double R_fn() const { return kind == 3 ? 1.0 : (kind == 2 ? pf_exp(theta[1]) : pf_exp(theta[2])); }
  double Z_fn(double alpha) const { return kind == 3 ? alpha : pf_exp(alpha); }
  double T_fn(double alpha) const {
    return kind == 3 ? alpha : (kind == 2 ? pf_sin(alpha) : theta[0] * (1 - theta[1]) + theta[1] * alpha);
  }
};

/* ssm_nlg::bsf_filter, src/model_ssm_nlg.cpp:1066-1146 for the one-state ar_exp model */
static double nlg_ar_exp_bsf_filter(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                                    double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  ArExp fn{M->theta, M->model_id};
  const int n = M->n;
  const size_t N = nsim;
  const double a1 = fn.a1_fn(), P1 = fn.P1_fn();
  std::vector<double> L_P1;
  psd_chol(&P1, 1, L_P1);
  auto A = [&](int t, size_t i) -> double& { return alpha[t + (size_t)(n + 1) * i]; };
  for (size_t i = 0; i < N; ++i) A(0, i) = a1 + L_P1[0] * pf_normal(seed, filter, 0, 0, (uint32_t)i);
  std::vector<double> W(N), logw(N), r(N), atmp(N);
  double loglik = 0.0;
  auto weigh = [&](int t) {
    const double H = fn.H_fn();
    for (size_t i = 0; i < N; ++i) logw[i] = dmvnorm1(M->y[t], fn.Z_fn(A(t, i)), H);
  };
  if (std::isfinite(M->y[0])) {
    weigh(0);
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i) atmp[i] = A(t, ind[i]);
    const double R = fn.R_fn();
    for (size_t i = 0; i < N; ++i)
      A(t + 1, i) = fn.T_fn(atmp[i]) + R * pf_normal(seed, filter, t + 1, 0, (uint32_t)i);
    double* wcol = weights + N * (t + 1);
    if (t < (n - 1) && std::isfinite(M->y[t + 1])) {
      weigh(t + 1);
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* ssm_nlg::bsf_filter, src/model_ssm_nlg.cpp:1066-1146 (growth model, p = 1, m = k = 2) */
double orc_nlg_bsf_filter(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                          double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  *stopped_at = -1;
  if ((M->model_id == BSSM_NLG_AR_EXP || M->model_id == BSSM_NLG_SIN_EXP || M->model_id == BSSM_NLG_LINEAR_GAUSSIAN) && M->p == 1)
    return nlg_ar_exp_bsf_filter(M, nsim, seed, filter, arith, alpha, weights, indices, stopped_at);
  if (M->model_id != BSSM_NLG_GROWTH || M->p != 1) return pf_nan();
  Growth fn{M->theta, M->known_params};
  const int n = M->n, m = 2, k = 2;
  const size_t N = nsim;
  double a1[2], P1[4];
  fn.a1_fn(a1); fn.P1_fn(P1);
  std::vector<double> L_P1;
  psd_chol(P1, m, L_P1);
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  for (size_t i = 0; i < N; ++i) {
    double um[2];
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += L_P1[d + (size_t)m * j] * um[j];
      A(d, 0, i) = a1[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N);
  double loglik = 0.0;
  auto weigh = [&](int t) {
    const double H = fn.H_fn();
    for (size_t i = 0; i < N; ++i) logw[i] = dmvnorm1(M->y[t], fn.Z_fn(&A(0, t, i)), H);
  };
  if (std::isfinite(M->y[0])) {
    weigh(0);
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    double R[4];
    fn.R_fn(R);
    for (size_t i = 0; i < N; ++i) {
      double uk[2], Tv[2];
      for (int j = 0; j < k; ++j) uk[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * k + j));
      fn.T_fn(&atmp[(size_t)m * i], Tv);
      /* R is diagonal: the off-diagonal zeros contribute +-0 */
      A(0, t + 1, i) = Tv[0] + R[0] * uk[0];
      A(1, t + 1, i) = Tv[1] + R[3] * uk[1];
    }
    double* wcol = weights + N * (t + 1);
    if (t < (n - 1) && std::isfinite(M->y[t + 1])) {
      weigh(t + 1);
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

/* ssm_nlg::psi_filter, src/model_ssm_nlg.cpp:971-1063, with log_weights :880-935 (the
 * per-particle dmvnorm branch: time_varying = TRUE is the R default, R/models.R:1547).
 * The Gaussian approximation is include/bssm_b200/nlg_gauss_approx.h. */
static double nlg_psi_filter_impl(const bssm_nlg_model* M, const double* mode_in, uint32_t max_iter, double conv_tol,
                                  uint32_t iekf_iter, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                                  double* alpha, double* weights, uint32_t* indices, double* approx_loglik,
                                  int32_t* stopped_at, double* mode_out) {
  *stopped_at = -1;
  if (M->model_id < BSSM_NLG_GROWTH || M->model_id > BSSM_NLG_LINEAR_GAUSSIAN || M->p != 1) return pf_nan();
  pfga::NlgApprox ga(M);
  ga.max_iter = (int)max_iter;
  ga.conv_tol = conv_tol;
  ga.iekf_iter = (int)iekf_iter;
  ga.run(mode_in);
  if (!ga.finite_mode) return -ORC_INF;
  if (approx_loglik) *approx_loglik = ga.approx_loglik;
  if (mode_out) std::copy(ga.mode.begin(), ga.mode.end(), mode_out);
  const pfga::NlgFns& fn = ga.fn;
  const int n = M->n, m = ga.m; /* 1 or 2 */
  const size_t N = nsim, mm = (size_t)m * m;
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  for (size_t i = 0; i < N; ++i) {
    double um[2];
    for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < m; ++d) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += ga.L[d + (size_t)m * j] * um[j];
      A(d, 0, i) = ga.alphahat[d] + s;
    }
  }
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N, 0.0);
  /* log_weights(t, alpha, alpha_prev) for one particle */
  auto log_weight = [&](int t, const double* a, const double* aprev) {
    double w = 0.0;
    if (std::isfinite(M->y[t])) {
      double za = 0.0;
      for (int d = 0; d < m; ++d) za += ga.Z[d + (size_t)m * t] * a[d];
      w = pfga::dnorm1_sd(M->y[t], fn.Z_fn(t, a), fn.H_fn(t, a)) -
          pfga::dnorm1_var(M->y[t], ga.D[t] + za, ga.H[t] * ga.H[t]);
    }
    if (t > 0) {
      double mean[2], amean[2], tmpv[2];
      fn.T_fn(t - 1, aprev, mean);
      pf_mm(&ga.T[mm * (t - 1)], aprev, tmpv, m, m, 1);
      for (int d = 0; d < m; ++d) amean[d] = ga.C[d + (size_t)m * (t - 1)] + tmpv[d];
      w -= ga.f_approx[t - 1].logpdf(a, amean) - ga.f_true[t - 1].logpdf(a, mean);
    }
    return w;
  };
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) {
    for (size_t i = 0; i < N; ++i) logw[i] = log_weight(0, &A(0, 0, i), &atmp[(size_t)m * i]) - ga.scales[0];
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + ga.approx_loglik + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    loglik = ga.approx_loglik;
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      for (int d = 0; d < m; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
    const double* ah0 = &ga.alphahat[(size_t)m * t];
    const double* ah1 = &ga.alphahat[(size_t)m * (t + 1)];
    const double* Ctm = &ga.Ct[mm * (t + 1)];
    const double* Lm = &ga.L[mm * (t + 1)];
    for (size_t i = 0; i < N; ++i) {
      double um[2], diff[2];
      for (int j = 0; j < m; ++j) um[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * m + j));
      for (int j = 0; j < m; ++j) diff[j] = atmp[j + (size_t)m * i] - ah0[j];
      for (int d = 0; d < m; ++d) {
        double cd = 0.0, lu = 0.0;
        for (int j = 0; j < m; ++j) cd += Ctm[d + (size_t)m * j] * diff[j];
        for (int j = 0; j < m; ++j) lu += Lm[d + (size_t)m * j] * um[j];
        A(d, t + 1, i) = ah1[d] + cd + lu;
      }
    }
    double* wcol = weights + N * (t + 1);
    if (t < (n - 1) && std::isfinite(M->y[t + 1])) {
      for (size_t i = 0; i < N; ++i)
        logw[i] = log_weight(t + 1, &A(0, t + 1, i), &atmp[(size_t)m * i]) - ga.scales[t + 1];
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

double orc_nlg_psi_filter(const bssm_nlg_model* M, const double* mode_in, uint32_t max_iter, double conv_tol,
                          uint32_t nsim, uint32_t seed, uint32_t filter, int arith, double* alpha,
                          double* weights, uint32_t* indices, double* approx_loglik, int32_t* stopped_at,
                          double* mode_out) {
  return nlg_psi_filter_impl(M, mode_in, max_iter, conv_tol, 0, nsim, seed, filter, arith, alpha, weights, indices,
                             approx_loglik, stopped_at, mode_out);
}
/* the same with an iterated-EKF start (iekf_iter > 0, src/model_ssm_nlg.cpp:325-356) */
double orc_nlg_psi_filter_iekf(const bssm_nlg_model* M, uint32_t max_iter, double conv_tol, uint32_t iekf_iter,
                               uint32_t nsim, uint32_t seed, uint32_t filter, int arith, double* alpha,
                               double* weights, uint32_t* indices, double* approx_loglik, int32_t* stopped_at,
                               double* mode_out) {
  return nlg_psi_filter_impl(M, nullptr, max_iter, conv_tol, iekf_iter, nsim, seed, filter, arith, alpha, weights,
                             indices, approx_loglik, stopped_at, mode_out);
}

/* ssm_nlg::ekf log-likelihood (src/model_ssm_nlg.cpp:284-378), p = 1: the known answers of
 * tests/testthat/test_ekpf.R:64-69 pin the (iterated) EKF that starts the psi-APF approximation;
 * att_out (m x n, optional) receives the filtered means */
double orc_nlg_ekf_loglik(const bssm_nlg_model* M, uint32_t iekf_iter, double* att_out) {
  pfga::NlgApprox A(M);
  A.iekf_iter = (int)iekf_iter;
  std::vector<double> at, att;
  double ll;
  A.ekf(at, att, &ll);
  if (att_out) std::memcpy(att_out, att.data(), att.size() * sizeof(double));
  return ll;
}

/* ------------------------------------------------------------------ a14 */
/* GBM user functions, R/cpp_example_models.R:129-171 */
struct Gbm {
  const double* theta;
  double drift(double x) const { return theta[0] * x; }
  double diffusion(double x) const { return std::max(0.0, theta[1] * x); }
  double ddiffusion(double x) const { return theta[1] * (x > 0.0); }
  /* R::dnorm(y, log(alpha), theta(2), 1)  [UNVENDORED: R nmath dnorm.c:
   * -(M_LN_SQRT_2PI + 0.5 * x * x + log(sigma)) with x = (y - mu) / sigma] */
  double log_obs(double y, double alpha) const {
    double x = (y - pf_log(alpha)) / theta[2];
    return -(0.918938533204672741780329736406 + 0.5 * x * x + pf_log(theta[2]));
  }
};

/* a model given by its four function pointers, as the reference takes them (src/model_ssm_sde.h:8-11):
 * the tests compile the SAME source text the engine hands to NVRTC with g++ and pass the addresses */
typedef double (*orc_sde_fn)(double x, const double* theta);
typedef double (*orc_sde_obs_fn)(double y, double x, const double* theta);
static orc_sde_fn g_user_drift = nullptr, g_user_diffusion = nullptr, g_user_ddiffusion = nullptr;
static orc_sde_obs_fn g_user_log_obs = nullptr;
struct UserSde {
  const double* theta;
  double drift(double x) const { return g_user_drift(x, theta); }
  double diffusion(double x) const { return g_user_diffusion(x, theta); }
  double ddiffusion(double x) const { return g_user_ddiffusion(x, theta); }
  double log_obs(double y, double alpha) const { return g_user_log_obs(y, alpha, theta); }
};

/* Poisson-OU user functions, R/cpp_example_models.R:41-86: theta = (log rho, nu, log sigma) */
struct PoissonOu {
  const double* theta;
  double drift(double x) const { return pf_exp(theta[0]) * (theta[1] - x); }
  double diffusion(double) const { return pf_exp(theta[2]); }
  double ddiffusion(double) const { return 0.0; }
  /* R::dpois(y, exp(alpha), 1) [UNVENDORED: R nmath dpois.c, via stirlerr / bd0] = y alpha - exp(alpha) - lgamma(y + 1) */
  double log_obs(double y, double alpha) const { return (y * alpha - pf_exp(alpha)) - std::lgamma(y + 1.0); }
};

/* milstein + milstein_worker, src/milstein.cpp:18-48; dB(k) = N(0, sqrt(dt)) drawn
 * as sqrt(dt) * z with z the (particle i, sub-step k) normal of the Philox stream */
} /* extern "C": the templates below need C++ linkage */

/* ssm_nlg::ekf_filter, src/model_ssm_nlg.cpp:1150-1260 (EKF-based particle filter), p = 1, M states.
 * The per-particle linear algebra is include/bssm_b200/pf_ekpf.h (shared with the device policy,
 * like pf_math.h); the loop structure and the model functions (pfga::NlgFns) are restated here. */
template <int M>
static double nlg_ekf_filter(const bssm_nlg_model* Mo, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                             double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  *stopped_at = -1;
  pfga::NlgFns fn{Mo->model_id, Mo->theta, Mo->known_params};
  const int n = Mo->n, m = M;
  const size_t N = nsim;
  double a1[2] = {0, 0}, P1[4] = {0, 0, 0, 0}, att1[2], Ptt1[4], L1[4]; /* sized for the largest registry model */
  fn.a1_fn(a1);
  fn.P1_fn(P1);
  auto update = [&](int t, const double* at, const double* Pt, double* att, double* Ptt) { /* ekf_update_step */
    if (std::isfinite(Mo->y[t])) {
      double Zg[2];
      fn.Z_gn(t, at, Zg);
      const double H = fn.H_fn(t, at);
      pf_ekf_update<M>(Mo->y[t], at, Pt, Zg, fn.Z_fn(t, at), H * H, att, Ptt);
    } else {
      for (int d = 0; d < M; ++d) att[d] = at[d];
      for (int e = 0; e < M * M; ++e) Ptt[e] = Pt[e];
    }
  };
  update(0, a1, P1, att1, Ptt1);
  pf_psd_chol<M>(Ptt1, L1);
  auto A = [&](int d, int t, size_t i) -> double& { return alpha[d + (size_t)m * (t + (size_t)(n + 1) * i)]; };
  for (size_t i = 0; i < N; ++i) {
    double um[2];
    for (int j = 0; j < M; ++j) um[j] = pf_normal(seed, filter, 0, 0, (uint32_t)(i * m + j));
    for (int d = 0; d < M; ++d) {
      double sacc = 0.0;
      for (int j = 0; j < M; ++j) sacc += L1[d + M * j] * um[j];
      A(d, 0, i) = att1[d] + sacc;
    }
  }
  auto log_obs = [&](int t, const double* a) { return pfga::dnorm1_sd(Mo->y[t], fn.Z_fn(t, a), fn.H_fn(t, a)); };
  std::vector<double> W(N), logw(N), r(N), atmp((size_t)m * N), attv((size_t)m * N), Lv((size_t)m * m * N);
  double loglik = 0.0;
  if (std::isfinite(Mo->y[0])) {
    for (size_t i = 0; i < N; ++i) {
      double w = log_obs(0, &A(0, 0, i));
      w += pf_dmvnorm_var<M>(&A(0, 0, i), a1, P1) - pf_dmvnorm_lwr<M>(&A(0, 0, i), att1, L1);
      logw[i] = w;
    }
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i) {
      for (int d = 0; d < M; ++d) atmp[d + (size_t)m * i] = A(d, t, ind[i]);
      double Rt[4] = {0, 0, 0, 0}, Pt[4], at[2] = {0, 0}, Ptt[4];
      fn.R_fn(t, &atmp[(size_t)m * i], Rt);
      pf_mmt(Rt, Rt, Pt, M, M, M);
      fn.T_fn(t, &atmp[(size_t)m * i], at);
      if (t < (n - 1)) {
        update(t + 1, at, Pt, &attv[(size_t)m * i], Ptt);
        pf_psd_chol<M>(Ptt, &Lv[(size_t)m * m * i]);
      } else { /* :1222-1225: the covariance itself takes the place of its factor */
        for (int d = 0; d < M; ++d) attv[d + (size_t)m * i] = at[d];
        for (int e = 0; e < M * M; ++e) Lv[e + (size_t)m * m * i] = Pt[e];
      }
    }
    for (size_t i = 0; i < N; ++i) {
      double um[2];
      for (int j = 0; j < M; ++j) um[j] = pf_normal(seed, filter, t + 1, 0, (uint32_t)(i * m + j));
      for (int d = 0; d < M; ++d) {
        double sacc = 0.0;
        for (int j = 0; j < M; ++j) sacc += Lv[d + M * j + (size_t)m * m * i] * um[j];
        A(d, t + 1, i) = attv[d + (size_t)m * i] + sacc;
      }
    }
    double* wcol = weights + N * (t + 1);
    if (t < (n - 1) && std::isfinite(Mo->y[t + 1])) {
      for (size_t i = 0; i < N; ++i) {
        double Rt[4] = {0, 0, 0, 0}, RR[4], mean[2] = {0, 0};
        fn.R_fn(t, &atmp[(size_t)m * i], Rt);
        pf_mmt(Rt, Rt, RR, M, M, M);
        fn.T_fn(t, &atmp[(size_t)m * i], mean);
        double w = log_obs(t + 1, &A(0, t + 1, i));
        w += pf_dmvnorm_var<M>(&A(0, t + 1, i), mean, RR) -
             pf_dmvnorm_lwr<M>(&A(0, t + 1, i), &attv[(size_t)m * i], &Lv[(size_t)m * m * i]);
        logw[i] = w;
      }
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

template <class Fn>
static double milstein(double x, unsigned int L, const Fn& fn, bool positive, uint32_t seed, uint32_t filter,
                       uint32_t t, uint32_t i) {
  unsigned int n = 1u << L;
  double dt = 1.0 / n;
  double sd = std::sqrt(dt);
  for (unsigned int k = 0; k < n; k++) {
    double dB = pf_normal(seed, filter, t, k, i) * sd;
    x += fn.drift(x) * dt + fn.diffusion(x) * dB + 0.5 * fn.diffusion(x) * fn.ddiffusion(x) * (dB * dB - dt);
    if (positive) x = std::abs(x);
  }
  return x;
}

/* ssm_sde::bsf_filter, src/model_ssm_sde.cpp:21-82 (alpha is 1 x (n+1) x N) */
template <class Fn>
static double sde_bsf_filter(const Fn& fn, const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed,
                             uint32_t filter, int arith, double* alpha, double* weights, uint32_t* indices,
                             int32_t* stopped_at) {
  *stopped_at = -1;
  const int n = M->n;
  const size_t N = nsim;
  auto A = [&](int t, size_t i) -> double& { return alpha[t + (size_t)(n + 1) * i]; };
  for (size_t i = 0; i < N; ++i) A(0, i) = milstein(M->x0, L, fn, M->positive, seed, filter, 0, (uint32_t)i);
  std::vector<double> W(N), logw(N), r(N);
  double loglik = 0.0;
  if (std::isfinite(M->y[0])) {
    for (size_t i = 0; i < N; ++i) logw[i] = fn.log_obs(M->y[0], A(0, i));
    double mxw, lmean;
    if (!normalise(logw.data(), nsim, weights, W, &mxw, &lmean, arith)) { *stopped_at = 0; return -ORC_INF; }
    loglik = mxw + lmean;
  } else {
    for (size_t i = 0; i < N; ++i) { weights[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
  }
  for (int t = 0; t < n; ++t) {
    for (size_t i = 0; i < N; ++i) r[i] = pf_resample_uniform(seed, filter, t, (uint32_t)i);
    uint32_t* ind = indices + N * t;
    resample(W, r, nsim, ind, arith);
    for (size_t i = 0; i < N; ++i)
      A(t + 1, i) = milstein(A(t, ind[i]), L, fn, M->positive, seed, filter, t + 1, (uint32_t)i);
    double* wcol = weights + N * (t + 1);
    if ((t < (n - 1)) && std::isfinite(M->y[t + 1])) {
      for (size_t i = 0; i < N; ++i) logw[i] = fn.log_obs(M->y[t + 1], A(t + 1, i));
      double mxw, lmean;
      if (!normalise(logw.data(), nsim, wcol, W, &mxw, &lmean, arith)) { *stopped_at = t + 1; return -ORC_INF; }
      loglik += mxw + lmean;
    } else {
      for (size_t i = 0; i < N; ++i) { wcol[i] = 1.0; W[i] = (arith == ORC_ARITH_REFERENCE) ? 1.0 / nsim : 1.0; }
    }
  }
  return loglik;
}

extern "C" {

double orc_nlg_ekf_filter(const bssm_nlg_model* M, uint32_t nsim, uint32_t seed, uint32_t filter, int arith,
                          double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  if (M->p != 1) { *stopped_at = -1; return pf_nan(); }
  if (M->model_id == BSSM_NLG_GROWTH)
    return nlg_ekf_filter<2>(M, nsim, seed, filter, arith, alpha, weights, indices, stopped_at);
  if (M->model_id == BSSM_NLG_AR_EXP || M->model_id == BSSM_NLG_SIN_EXP || M->model_id == BSSM_NLG_LINEAR_GAUSSIAN)
    return nlg_ekf_filter<1>(M, nsim, seed, filter, arith, alpha, weights, indices, stopped_at);
  *stopped_at = -1;
  return pf_nan();
}

double orc_sde_bsf_filter(const bssm_sde_model* M, uint32_t nsim, uint32_t L, uint32_t seed, uint32_t filter,
                          int arith, double* alpha, double* weights, uint32_t* indices, int32_t* stopped_at) {
  if (M->model_id == BSSM_SDE_GBM)
    return sde_bsf_filter(Gbm{M->theta}, M, nsim, L, seed, filter, arith, alpha, weights, indices, stopped_at);
  if (M->model_id == BSSM_SDE_POISSON_OU)
    return sde_bsf_filter(PoissonOu{M->theta}, M, nsim, L, seed, filter, arith, alpha, weights, indices, stopped_at);
  if (M->model_id >= BSSM_USER_MODEL_BASE && g_user_drift && g_user_diffusion && g_user_ddiffusion && g_user_log_obs)
    return sde_bsf_filter(UserSde{M->theta}, M, nsim, L, seed, filter, arith, alpha, weights, indices, stopped_at);
  *stopped_at = -1;
  return pf_nan();
}
/* the function pointers every model_id >= BSSM_USER_MODEL_BASE runs with */
void orc_sde_set_user(void* drift, void* diffusion, void* ddiffusion, void* log_obs) {
  g_user_drift = (orc_sde_fn)drift; g_user_diffusion = (orc_sde_fn)diffusion;
  g_user_ddiffusion = (orc_sde_fn)ddiffusion; g_user_log_obs = (orc_sde_obs_fn)log_obs;
}

/* ------------------------------------------------------------------ a15 */
/* update_model(theta): svm src/model_svm.cpp:12-29, ar1_ng src/model_ar1_ng.cpp:12-36.
 * Writes the scalars of the per-draw model into caller storage. */
struct DrawModel {
  bssm_ung_model M;
  double T, R, a1, P1, C;
  std::vector<double> xb; /* xbeta of this draw when beta is part of theta */
};
static void update_model(DrawModel& dm, const bssm_ung_model* base, const bssm_theta_map* map,
                         const double* new_theta, uint32_t n_par) {
  dm.M = *base;
  dm.T = base->T[0]; dm.R = base->R[0]; dm.a1 = base->a1[0]; dm.P1 = base->P1[0]; dm.C = base->C[0];
  if (map->model_type == 3) {
    if (map->svm_type == 0) {
      dm.M.phi = new_theta[2];
    } else {
      dm.a1 = new_theta[2];
      dm.C = new_theta[2] * (1.0 - new_theta[0]);
    }
    dm.T = new_theta[0];
    dm.R = new_theta[1];
    dm.P1 = new_theta[1] * new_theta[1] / (1 - new_theta[0] * new_theta[0]);
  } else {
    double rho = new_theta[0];
    double sigma = pf_exp(new_theta[1]);
    dm.T = rho;
    dm.R = sigma;
    double RR = sigma * sigma;
    if (map->mu_est) {
      dm.a1 = new_theta[2];
      dm.C = new_theta[2] * (1.0 - rho);
    }
    dm.P1 = RR / (1.0 - rho * rho);
    if (map->phi_est) dm.M.phi = pf_exp(new_theta[2 + map->mu_est]);
  }
  dm.M.T = &dm.T; dm.M.R = &dm.R; dm.M.a1 = &dm.a1; dm.M.P1 = &dm.P1; dm.M.C = &dm.C;
  dm.M.Ttv = 0; dm.M.Rtv = 0; dm.M.Ctv = 0;
  if (map->n_xreg > 0) {
    /* beta = new_theta.subvec(n_elem - xreg.n_cols, n_elem - 1); compute_xbeta()
     * (src/model_ar1_ng.cpp:28-31, src/model_ssm_ung.h:83: xbeta = xreg * beta) */
    const size_t n = (size_t)base->n;
    dm.xb.assign(n, 0.0);
    for (size_t t = 0; t < n; ++t) {
      double sum = 0.0;
      for (int j = 0; j < map->n_xreg; ++j) sum += map->xreg[t + n * j] * new_theta[n_par - map->n_xreg + j];
      dm.xb[t] = sum;
    }
    dm.M.xbeta = dm.xb.data();
  }
}

/* approx_mcmc::is_correction_psi, src/approx_mcmc.cpp:353-427, with expand()
 * (:49-88) for is_type 3; m = 1.  OpenMP over draws with schedule(static) as
 * in the reference (:365-369).  Deliberate difference: draw d uses the Philox
 * stream `filter = d` instead of a per-thread engine (:367), so the result
 * does not depend on `threads`.  The output_type 2 merge runs in draw order
 * after the parallel loop (the reference's omp critical order is
 * nondeterministic for threads > 1).
 * weights, posterior: n_draws; alpha: (n+1) x m x n_draws (output_type 1);
 * alphahat (n+1) x m, Vt m x m x (n+1) (output_type 2). */
int orc_postcorrection_ung(const bssm_ung_model* M, const bssm_theta_map* map, int output_type,
                           uint32_t nsim, uint32_t seed, int threads, int is_type, const uint32_t* counts,
                           const double* theta, uint32_t n_par, const double* modes, uint32_t n_stored,
                           int arith, double* weights, double* posterior, double* alpha, double* alphahat,
                           double* Vt, double* loglik_out) {
  if (M->m != 1) return 1;
  const size_t n = M->n, n1 = n + 1;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(n1 * nd, 0.0); Vt_all.assign(n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    DrawModel dm;
    update_model(dm, M, map, theta + (size_t)src[i] * n_par, n_par);
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    int32_t st;
    double al;
    double ll = orc_ung_psi_filter(&dm.M, modes + n * src[i], nsimc, seed, (uint32_t)i, arith, alpha_i.data(),
                                   weights_i.data(), indices.data(), &al, &st);
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = pf_exp(ll);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), 1, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t) alpha[n1 * i + t] = alpha_i[t + n1 * b];
      } else {
        orc_summary(alpha_i.data(), 1, (int)n1, nsimc, &ah_all[n1 * i], &Vt_all[n1 * i]);
      }
    }
  }
  if (output_type == 2) { /* :397-424, in draw order */
    std::vector<double> ah(n1, 0.0), V(n1, 0.0), Valpha(n1, 0.0);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        double diff = ah_all[n1 * i + t] - ah[t];
        ah[t] += wnew / sum_w * diff;
        double diff2 = ah_all[n1 * i + t] - ah[t];
        Valpha[t] += wnew * diff * diff2;
        V[t] += wnew / sum_w * (Vt_all[n1 * i + t] - V[t]);
      }
    }
    for (size_t t = 0; t < n1; ++t) { alphahat[t] = ah[t]; Vt[t] = V[t] + Valpha[t] / sum_w; }
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + 0.0 + pf_log(weights[i]);
  return 0;
}

/* converged mode of ssm_ung::approximate() (src/model_ssm_ung.cpp:87-140) for any m: what
 * run_mcmc stores per draw in mcmc_output$modes */
void orc_ung_approx_mode_m(const bssm_ung_model* M, double* mode) {
  pfga::UngApprox ga(M);
  ga.approximate();
  std::copy(ga.mode.begin(), ga.mode.end(), mode);
}

/* is_correction_psi<bsm_ng> (src/approx_mcmc.cpp:353-427) for any state dimension m:
 * bsm_ng::update_model (src/model_bsm_ng.cpp:17-51) per draw, then the general-m
 * psi-APF with approximate_for_is at the stored mode.  Same conventions as
 * orc_postcorrection_ung; alpha (n+1) x m x n_draws, alphahat (n+1) x m, Vt m x m x (n+1). */
int orc_postcorrection_bsm(const bssm_ung_model* M, const bssm_theta_map* map, int output_type, uint32_t nsim,
                           uint32_t seed, int threads, int is_type, const uint32_t* counts, const double* theta,
                           uint32_t n_par, const double* modes, uint32_t n_stored, int arith, double* weights,
                           double* posterior, double* alpha, double* alphahat, double* Vt, double* loglik_out) {
  if (map->model_type != 2) return 1;
  const size_t n = M->n, n1 = n + 1, m = M->m, k = M->k;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(m * n1 * nd, 0.0); Vt_all.assign(m * m * n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    const double* new_theta = theta + (size_t)src[i] * n_par;
    /* model.update_model(theta_i), src/model_bsm_ng.cpp:17-45 */
    bssm_ung_model Mi = *M;
    std::vector<double> R(M->R, M->R + m * k), P1(M->P1, M->P1 + m * m), xb;
    const int slope = map->slope, seasonal = map->seasonal, noise = map->noise;
    const int level_est = map->level_est, slope_est = map->slope_est, seasonal_est = map->seasonal_est;
    if (level_est) R[0 + m * 0] = pf_exp(new_theta[0]);
    if (slope_est) R[1 + m * 1] = pf_exp(new_theta[level_est]);
    if (seasonal_est) R[(1 + slope) + m * (1 + slope)] = pf_exp(new_theta[level_est + slope_est]);
    if (noise) {
      R[(m - 1) + m * (1 + slope + seasonal)] = pf_exp(new_theta[level_est + slope_est + seasonal_est]);
      const double r = R[(m - 1) + m * (1 + slope + seasonal)];
      P1[(m - 1) + m * (m - 1)] = r * r;
    }
    if (map->phi_est) Mi.phi = pf_exp(new_theta[level_est + slope_est + seasonal_est + noise]);
    if (map->n_xreg > 0) {
      xb.assign(n, 0.0);
      for (size_t t = 0; t < n; ++t) {
        double sum = 0.0;
        for (int j = 0; j < map->n_xreg; ++j) sum += map->xreg[t + n * j] * new_theta[n_par - map->n_xreg + j];
        xb[t] = sum;
      }
      Mi.xbeta = xb.data();
    }
    Mi.R = R.data(); Mi.Rtv = 0; Mi.P1 = P1.data();
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(m * n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    int32_t st;
    double al;
    double ll = orc_ung_psi_filter_m(&Mi, modes + n * src[i], nsimc, seed, (uint32_t)i, arith, alpha_i.data(),
                                     weights_i.data(), indices.data(), &al, &st, nullptr, nullptr, nullptr);
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = pf_exp(ll);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), (int)m, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t)
          for (size_t d = 0; d < m; ++d) alpha[m * n1 * i + t + n1 * d] = alpha_i[d + m * (t + n1 * b)];
      } else {
        orc_summary(alpha_i.data(), (int)m, (int)n1, nsimc, &ah_all[m * n1 * i], &Vt_all[m * m * n1 * i]);
      }
    }
  }
  if (output_type == 2) { /* :397-424, in draw order */
    std::vector<double> ah(m * n1, 0.0), V(m * m * n1, 0.0), Valpha(m * m * n1, 0.0), diff(m), diff2(m);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        for (size_t d = 0; d < m; ++d) {
          diff[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
          ah[d + m * t] += wnew / sum_w * diff[d];
          diff2[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
        }
        for (size_t c = 0; c < m; ++c)
          for (size_t r = 0; r < m; ++r) {
            Valpha[r + m * (c + m * t)] += wnew * diff[r] * diff2[c];
            V[r + m * (c + m * t)] += wnew / sum_w * (Vt_all[m * m * n1 * i + r + m * (c + m * t)] - V[r + m * (c + m * t)]);
          }
      }
    }
    for (size_t t = 0; t < n1; ++t)
      for (size_t d = 0; d < m; ++d) alphahat[t + n1 * d] = ah[d + m * t]; /* alphahat.t() */
    for (size_t e = 0; e < m * m * n1; ++e) Vt[e] = V[e] + Valpha[e] / sum_w;
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + 0.0 + pf_log(weights[i]);
  return 0;
}

/* is_correction_psi<ssm_nlg> (src/approx_mcmc.cpp:353-427), growth model: theta is
 * n_par x n_stored, modes m x n x n_stored.  Same conventions as orc_postcorrection_ung. */
/* shared draw loop of the two ssm_nlg correction phases: modes != null is is_correction_psi (psi-APF around the
 * stored mode, weight = exp(loglik)), approx_loglik != null is is_correction_bsf (bootstrap filter,
 * weight = exp(loglik - approx_loglik_i)), src/approx_mcmc.cpp:345-425 and :447-518 */
static int nlg_correction(const bssm_nlg_model* M, int output_type, uint32_t nsim, uint32_t seed, int threads,
                          int is_type, const uint32_t* counts, const double* theta, uint32_t n_par,
                          const double* modes, const double* approx_loglik, uint32_t n_stored, int arith,
                          double* weights, double* posterior, double* alpha, double* alphahat, double* Vt,
                          double* loglik_out) {
  if (M->model_id < BSSM_NLG_GROWTH || M->model_id > BSSM_NLG_LINEAR_GAUSSIAN || M->p != 1) return 1;
  const size_t n = M->n, n1 = n + 1, m = M->model_id == BSSM_NLG_GROWTH ? 2 : 1;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(m * n1 * nd, 0.0); Vt_all.assign(m * m * n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    bssm_nlg_model Mi = *M;
    Mi.theta = theta + (size_t)src[i] * n_par; /* model.update_model(theta_i) */
    Mi.n_theta = (int32_t)n_par;
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(m * n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    int32_t st;
    double al;
    double ll = modes ? orc_nlg_psi_filter(&Mi, modes + m * n * src[i], 100, 1e-8, nsimc, seed, (uint32_t)i, arith,
                                           alpha_i.data(), weights_i.data(), indices.data(), &al, &st, nullptr)
                      : orc_nlg_bsf_filter(&Mi, nsimc, seed, (uint32_t)i, arith, alpha_i.data(), weights_i.data(),
                                           indices.data(), &st);
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = modes ? pf_exp(ll) : pf_exp(ll - approx_loglik[src[i]]);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), (int)m, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t)
          for (size_t d = 0; d < m; ++d) alpha[m * n1 * i + t + n1 * d] = alpha_i[d + m * (t + n1 * b)];
      } else {
        orc_summary(alpha_i.data(), (int)m, (int)n1, nsimc, &ah_all[m * n1 * i], &Vt_all[m * m * n1 * i]);
      }
    }
  }
  if (output_type == 2) {
    std::vector<double> ah(m * n1, 0.0), V(m * m * n1, 0.0), Valpha(m * m * n1, 0.0);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        double diff[2], diff2[2];
        for (size_t d = 0; d < m; ++d) {
          diff[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
          ah[d + m * t] += wnew / sum_w * diff[d];
          diff2[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
        }
        for (size_t c = 0; c < m; ++c)
          for (size_t r = 0; r < m; ++r) {
            Valpha[r + m * (c + m * t)] += wnew * diff[r] * diff2[c];
            V[r + m * (c + m * t)] += wnew / sum_w * (Vt_all[m * m * n1 * i + r + m * (c + m * t)] - V[r + m * (c + m * t)]);
          }
      }
    }
    for (size_t t = 0; t < n1; ++t)
      for (size_t d = 0; d < m; ++d) alphahat[t + n1 * d] = ah[d + m * t]; /* alphahat.t() */
    for (size_t e = 0; e < m * m * n1; ++e) Vt[e] = V[e] + Valpha[e] / sum_w;
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + 0.0 + pf_log(weights[i]);
  return 0;
}
int orc_postcorrection_nlg(const bssm_nlg_model* M, int output_type, uint32_t nsim, uint32_t seed, int threads,
                           int is_type, const uint32_t* counts, const double* theta, uint32_t n_par,
                           const double* modes, uint32_t n_stored, int arith, double* weights, double* posterior,
                           double* alpha, double* alphahat, double* Vt, double* loglik_out) {
  if (!modes) return 1;
  return nlg_correction(M, output_type, nsim, seed, threads, is_type, counts, theta, n_par, modes, nullptr, n_stored,
                        arith, weights, posterior, alpha, alphahat, Vt, loglik_out);
}
/* is_correction_bsf<ssm_nlg>, src/approx_mcmc.cpp:442-518: update_model(theta_i), bsf_filter,
 * weight = exp(loglik - approx_loglik_i), posterior = prior (0 here) + log(weight); draw d uses Philox stream d */
int orc_is_correction_bsf_nlg(const bssm_nlg_model* M, int output_type, uint32_t nsim, uint32_t seed, int threads,
                              int is_type, const uint32_t* counts, const double* theta, uint32_t n_par,
                              const double* approx_loglik, uint32_t n_stored, int arith, double* weights,
                              double* posterior, double* alpha, double* alphahat, double* Vt, double* loglik_out) {
  if (!approx_loglik) return 1;
  return nlg_correction(M, output_type, nsim, seed, threads, is_type, counts, theta, n_par, nullptr, approx_loglik,
                        n_stored, arith, weights, posterior, alpha, alphahat, Vt, loglik_out);
}

/* is_correction_bsf, src/approx_mcmc.cpp:447-518 (svm / ar1_ng, m = 1): bootstrap filter per stored
 * draw after update_model(theta_i); weight = exp(loglik - approx_loglik), posterior = log(weight). */
int orc_is_correction_bsf_ung(const bssm_ung_model* M, const bssm_theta_map* map, int output_type, uint32_t nsim,
                              uint32_t seed, int threads, int is_type, const uint32_t* counts, const double* theta,
                              uint32_t n_par, const double* approx_loglik, uint32_t n_stored, int arith,
                              double* weights, double* posterior, double* alpha, double* alphahat, double* Vt,
                              double* loglik_out) {
  if (M->m != 1) return 1;
  const size_t n = M->n, n1 = n + 1;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(n1 * nd, 0.0); Vt_all.assign(n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    DrawModel dm;
    update_model(dm, M, map, theta + (size_t)src[i] * n_par, n_par);
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    int32_t st;
    double ll = orc_ung_bsf_filter(&dm.M, nsimc, seed, (uint32_t)i, arith, alpha_i.data(), weights_i.data(),
                                   indices.data(), &st);
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = pf_exp(ll - approx_loglik[src[i]]);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), 1, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t) alpha[n1 * i + t] = alpha_i[t + n1 * b];
      } else {
        orc_summary(alpha_i.data(), 1, (int)n1, nsimc, &ah_all[n1 * i], &Vt_all[n1 * i]);
      }
    }
  }
  if (output_type == 2) {
    std::vector<double> ah(n1, 0.0), V(n1, 0.0), Valpha(n1, 0.0);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        double diff = ah_all[n1 * i + t] - ah[t];
        ah[t] += wnew / sum_w * diff;
        double diff2 = ah_all[n1 * i + t] - ah[t];
        Valpha[t] += wnew * diff * diff2;
        V[t] += wnew / sum_w * (Vt_all[n1 * i + t] - V[t]);
      }
    }
    for (size_t t = 0; t < n1; ++t) { alphahat[t] = ah[t]; Vt[t] = V[t] + Valpha[t] / sum_w; }
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + pf_log(weights[i]);
  return 0;
}

/* is_correction_bsf<ssm_sde>, src/approx_mcmc.cpp:522-591: bootstrap filter on the fine mesh L_f per
 * stored draw, weight = exp(loglik - approx_loglik), posterior = approx_loglik + log(weight) (prior 0).
 * Draw d uses Philox stream d (the reference re-seeds two engines per thread, :537-538). */
int orc_is_correction_bsf_sde(const bssm_sde_model* M, uint32_t L_f, int output_type, uint32_t nsim, uint32_t seed,
                              int threads, int is_type, const uint32_t* counts, const double* theta,
                              uint32_t n_par, const double* approx_loglik, uint32_t n_stored, int arith,
                              double* weights, double* posterior, double* alpha, double* alphahat, double* Vt,
                              double* loglik_out) {
  const size_t n = M->n, n1 = n + 1;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(n1 * nd, 0.0); Vt_all.assign(n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    bssm_sde_model Mi = *M;
    Mi.theta = theta + (size_t)src[i] * n_par; /* model.theta = theta_storage.col(i) */
    Mi.n_theta = (int32_t)n_par;
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    int32_t st;
    double ll = orc_sde_bsf_filter(&Mi, nsimc, L_f, seed, (uint32_t)i, arith, alpha_i.data(), weights_i.data(),
                                   indices.data(), &st);
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = pf_exp(ll - approx_loglik[src[i]]);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), 1, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t) alpha[n1 * i + t] = alpha_i[t + n1 * b];
      } else {
        orc_summary(alpha_i.data(), 1, (int)n1, nsimc, &ah_all[n1 * i], &Vt_all[n1 * i]);
      }
    }
  }
  if (output_type == 2) {
    std::vector<double> ah(n1, 0.0), V(n1, 0.0), Valpha(n1, 0.0);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        double diff = ah_all[n1 * i + t] - ah[t];
        ah[t] += wnew / sum_w * diff;
        double diff2 = ah_all[n1 * i + t] - ah[t];
        Valpha[t] += wnew * diff * diff2;
        V[t] += wnew / sum_w * (Vt_all[n1 * i + t] - V[t]);
      }
    }
    for (size_t t = 0; t < n1; ++t) { alphahat[t] = ah[t]; Vt[t] = V[t] + Valpha[t] / sum_w; }
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + approx_loglik[src[i]] + pf_log(weights[i]);
  return 0;
}

/* ------------------------------------------------ parameter sets (parset_ng) */
} /* extern "C" */
/* the draw loop shared by is_correction_psi / is_correction_bsf of the ssm_ung and ssm_mng specialisations
 * (src/approx_mcmc.cpp:685-804, :806-921, :1038-1157, :1159-1273): filter(c, i, nsimc, alpha_i, weights_i,
 * indices) runs draw i (stored column c) and returns its log-likelihood; weight = exp(loglik [- approx_loglik_c]),
 * posterior = prior (0) + log(weight) */
template <class Filter>
static int parset_correction(size_t n, size_t m, int output_type, uint32_t nsim, uint32_t seed, int threads,
                             int is_type, const uint32_t* counts, uint32_t n_stored, const double* approx_loglik,
                             double* weights, double* posterior, double* alpha, double* alphahat, double* Vt,
                             double* loglik_out, Filter&& filter) {
  const size_t n1 = n + 1;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
  std::vector<double> ah_all, Vt_all;
  if (output_type == 2) { ah_all.assign(m * n1 * nd, 0.0); Vt_all.assign(m * m * n1 * nd, 0.0); }
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(m * n1 * nsimc, 0.0), weights_i((size_t)nsimc * n1, 0.0);
    std::vector<uint32_t> indices((size_t)nsimc * n, 0u);
    const double ll = filter(src[i], (uint32_t)i, nsimc, alpha_i.data(), weights_i.data(), indices.data());
    if (loglik_out) loglik_out[i] = ll;
    weights[i] = pf_exp(approx_loglik ? ll - approx_loglik[src[i]] : ll);
    if (output_type != 3) {
      orc_filter_smoother(alpha_i.data(), indices.data(), (int)m, (int)n1, nsimc);
      if (output_type == 1) {
        uint32_t b = (uint32_t)(pf_pick_uniform(seed, (uint32_t)i) * (double)nsimc);
        if (b > nsimc - 1) b = nsimc - 1;
        for (size_t t = 0; t < n1; ++t)
          for (size_t d = 0; d < m; ++d) alpha[m * n1 * i + t + n1 * d] = alpha_i[d + m * (t + n1 * b)];
      } else {
        orc_summary(alpha_i.data(), (int)m, (int)n1, nsimc, &ah_all[m * n1 * i], &Vt_all[m * m * n1 * i]);
      }
    }
  }
  if (output_type == 2) {
    std::vector<double> ah(m * n1, 0.0), V(m * m * n1, 0.0), Valpha(m * m * n1, 0.0), diff(m), diff2(m);
    double sum_w = 0.0;
    for (int64_t i = 0; i < nd; ++i) {
      double wnew = weights[i] * cnt[i];
      sum_w += wnew;
      for (size_t t = 0; t < n1; ++t) {
        for (size_t d = 0; d < m; ++d) {
          diff[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
          ah[d + m * t] += wnew / sum_w * diff[d];
          diff2[d] = ah_all[m * n1 * i + d + m * t] - ah[d + m * t];
        }
        for (size_t c = 0; c < m; ++c)
          for (size_t r = 0; r < m; ++r) {
            Valpha[r + m * (c + m * t)] += wnew * diff[r] * diff2[c];
            V[r + m * (c + m * t)] += wnew / sum_w * (Vt_all[m * m * n1 * i + r + m * (c + m * t)] - V[r + m * (c + m * t)]);
          }
      }
    }
    for (size_t t = 0; t < n1; ++t)
      for (size_t d = 0; d < m; ++d) alphahat[t + n1 * d] = ah[d + m * t]; /* alphahat.t() */
    for (size_t e = 0; e < m * m * n1; ++e) Vt[e] = V[e] + Valpha[e] / sum_w;
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = 0.0 + 0.0 + pf_log(weights[i]);
  return 0;
}

/* parset_ung::update(model, i), src/parset_ng.cpp:84-124 (xb: storage for compute_xbeta) */
static void parset_update_ung(const bssm_ung_model* M, const bssm_parset* ps, uint32_t c, bssm_ung_model& Mi,
                              std::vector<double>& xb) {
  Mi = *M;
  const size_t n = M->n, m = M->m, k = M->k;
  if (ps->phi) Mi.phi = ps->phi[c];
  if (ps->Z) Mi.Z = ps->Z + m * (M->Ztv ? n : 1) * c;
  if (ps->T) Mi.T = ps->T + m * m * (M->Ttv ? n : 1) * c;
  if (ps->R) Mi.R = ps->R + m * k * (M->Rtv ? n : 1) * c;
  if (ps->a1) Mi.a1 = ps->a1 + m * c;
  if (ps->P1) Mi.P1 = ps->P1 + m * m * c;
  if (ps->D) Mi.D = ps->D + (M->Dtv ? n : 1) * c;
  if (ps->C) Mi.C = ps->C + m * (M->Ctv ? n : 1) * c;
  if (ps->beta) {
    xb.assign(n, 0.0);
    for (size_t t = 0; t < n; ++t) {
      double sum = 0.0;
      for (int j = 0; j < ps->n_xreg; ++j) sum += ps->xreg[t + n * j] * ps->beta[(size_t)ps->n_xreg * c + j];
      xb[t] = sum;
    }
    Mi.xbeta = xb.data();
  }
}
/* parset_mng::update(model, i), src/parset_ng.cpp:185-215 */
static void parset_update_mng(const bssm_mng_model* M, const bssm_parset* ps, uint32_t c, bssm_mng_model& Mi) {
  Mi = *M;
  const size_t n = M->n, m = M->m, k = M->k, p = M->p;
  if (ps->phi) Mi.phi = ps->phi + p * c;
  if (ps->Z) Mi.Z = ps->Z + p * m * (M->Ztv ? n : 1) * c;
  if (ps->T) Mi.T = ps->T + m * m * (M->Ttv ? n : 1) * c;
  if (ps->R) Mi.R = ps->R + m * k * (M->Rtv ? n : 1) * c;
  if (ps->a1) Mi.a1 = ps->a1 + m * c;
  if (ps->P1) Mi.P1 = ps->P1 + m * m * c;
  if (ps->D) Mi.D = ps->D + p * (M->Dtv ? n : 1) * c;
  if (ps->C) Mi.C = ps->C + m * (M->Ctv ? n : 1) * c;
}
extern "C" {

/* approx_mcmc::is_correction_spdk (src/approx_mcmc.cpp:606-678), output_type 3, for parameter sets: per stored
 * draw parset_ung::update, approximate_for_is at the stored mode, simulate_states(nsimc), weight =
 * mean(exp(importance_weights - accu(scales))) (:636-639); posterior = approx_loglik + log(weight) (:676-677).
 * Draw i uses the Philox stream `filter = i` instead of a per-thread engine (:621), as in is_correction_psi. */
int orc_is_correction_spdk_ung_parset(const bssm_ung_model* M, const bssm_parset* ps, uint32_t nsim, uint32_t seed,
                                      int threads, int is_type, const uint32_t* counts, const double* modes,
                                      const double* approx_loglik, uint32_t n_stored, double* weights,
                                      double* posterior) {
  const size_t n = M->n, m = M->m;
  std::vector<uint32_t> src, cnt;
  for (uint32_t i = 0; i < n_stored; ++i) {
    if (is_type == 3) for (uint32_t c = 0; c < counts[i]; ++c) { src.push_back(i); cnt.push_back(1u); }
    else { src.push_back(i); cnt.push_back(counts[i]); }
  }
  const int64_t nd = (int64_t)src.size();
#ifdef _OPENMP
  if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t i = 0; i < nd; ++i) {
    bssm_ung_model Mi;
    std::vector<double> xb;
    parset_update_ung(M, ps, src[i], Mi, xb);
    pfga::UngApprox ga(&Mi);
    ga.run_head(modes + n * src[i]);
    unsigned int nsimc = nsim;
    if (is_type == 1) nsimc *= cnt[i];
    std::vector<double> alpha_i(m * (n + 1) * nsimc), w, ah;
    spdk_draws(&Mi, ga, nsimc, 1, seed, (uint32_t)i, alpha_i.data(), w, ah);
    const double ssum = pfga::accu2(ga.scales.data(), n);
    for (uint32_t j = 0; j < nsimc; ++j) w[j] = pf_exp(w[j] - ssum);
    weights[i] = pfga::accu2(w.data(), nsimc) / nsimc;
  }
  for (int64_t i = 0; i < nd; ++i) posterior[i] = (approx_loglik ? approx_loglik[src[i]] : 0.0) + pf_log(weights[i]);
  return 0;
}

/* is_correction_psi<ssm_ung> (src/approx_mcmc.cpp:685-804; modes != null) and is_correction_bsf<ssm_ung>
 * (:806-921; approx_loglik != null) with the parameter sets of src/parset_ng.cpp */
int orc_correction_ung_parset(const bssm_ung_model* M, const bssm_parset* ps, int output_type, uint32_t nsim,
                              uint32_t seed, int threads, int is_type, const uint32_t* counts, const double* modes,
                              const double* approx_loglik, uint32_t n_stored, int arith, double* weights,
                              double* posterior, double* alpha, double* alphahat, double* Vt, double* loglik_out) {
  if ((modes != nullptr) == (approx_loglik != nullptr)) return 1;
  const size_t n = M->n;
  return parset_correction(
      n, M->m, output_type, nsim, seed, threads, is_type, counts, n_stored, approx_loglik, weights, posterior, alpha,
      alphahat, Vt, loglik_out,
      [&](uint32_t c, uint32_t i, uint32_t nsimc, double* alpha_i, double* weights_i, uint32_t* indices) {
        bssm_ung_model Mi;
        std::vector<double> xb;
        parset_update_ung(M, ps, c, Mi, xb);
        int32_t st;
        double al;
        if (modes)
          return orc_ung_psi_filter_m(&Mi, modes + n * c, nsimc, seed, i, arith, alpha_i, weights_i, indices, &al, &st,
                                      nullptr, nullptr, nullptr);
        return orc_ung_bsf_filter(&Mi, nsimc, seed, i, arith, alpha_i, weights_i, indices, &st);
      });
}

/* is_correction_psi<ssm_mng> (src/approx_mcmc.cpp:1038-1157) and is_correction_bsf<ssm_mng> (:1159-1273);
 * modes is p x n x n_stored (mode_storage, :29) */
int orc_correction_mng_parset(const bssm_mng_model* M, const bssm_parset* ps, int output_type, uint32_t nsim,
                              uint32_t seed, int threads, int is_type, const uint32_t* counts, const double* modes,
                              const double* approx_loglik, uint32_t n_stored, int arith, double* weights,
                              double* posterior, double* alpha, double* alphahat, double* Vt, double* loglik_out) {
  if ((modes != nullptr) == (approx_loglik != nullptr)) return 1;
  const size_t n = M->n, p = M->p;
  return parset_correction(
      n, M->m, output_type, nsim, seed, threads, is_type, counts, n_stored, approx_loglik, weights, posterior, alpha,
      alphahat, Vt, loglik_out,
      [&](uint32_t c, uint32_t i, uint32_t nsimc, double* alpha_i, double* weights_i, uint32_t* indices) {
        bssm_mng_model Mi;
        parset_update_mng(M, ps, c, Mi);
        int32_t st;
        double al;
        if (modes)
          return orc_mng_psi_filter(&Mi, modes + p * n * c, nsimc, seed, i, arith, alpha_i, weights_i, indices, &al, &st);
        return orc_mng_bsf_filter(&Mi, nsimc, seed, i, arith, alpha_i, weights_i, indices, &st);
      });
}

/* dmvnorm(x, mean, sigma, lwr = false, logd = true) for a full-rank m x m covariance
 * (src/dmvnorm.cpp:4-60), the density the ssm_nlg weights are built from: test hook for the
 * reference's own known-answer test (tests/testthat/test_basics.R:176-201) */
double orc_dmvnorm_logpdf(const double* x, const double* mean, const double* sigma, int m) {
  pfga::MvnFactor f;
  f.set(sigma, m);
  return f.logpdf(x, mean);
}
double orc_dnorm1(double x, double mean, double s, int lwr) {
  return lwr ? pfga::dnorm1_sd(x, mean, s) : pfga::dnorm1_var(x, mean, s);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} /* extern "C" */
