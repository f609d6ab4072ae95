/* bssm_b200.h -- C ABI of the B200 particle-filter engine.
 *
 * This is the drop-in boundary for the particle-filter hot path of helske/bssm
 * (SURVEY.md section 8b).  Each entry point replaces one of the reference's
 * `.Call` targets; the citation after "replaces:" is the reference function
 * whose argument meaning, output layout and error behaviour it keeps.  All
 * arrays are column-major binary64 exactly as R / Armadillo hold them, so an R
 * shim can pass REAL(x) pointers without copying (r/bssm_b200_shim.c).
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; pointers are HOST
 *     pointers unless the name ends in _dev;
 *   - outputs are caller-allocated; a NULL output pointer means "not wanted";
 *   - return value: 0 = ok, >0 = error (message via bssm_b200_last_error());
 *     a filter that stops prematurely because all weights vanished is NOT an
 *     error: status 0, loglik = -inf, stopped_at = the time index, later
 *     columns left zero (reference: src/model_ssm_ung.cpp:556-560 and the
 *     warning at src/R_bootstrap_filter.cpp:166-167);
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns BSSM_B200_ERR_NO_DEVICE.
 */
#ifndef BSSM_B200_H
#define BSSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* a resident batch of filters (device-side inputs and outputs), see bssm_b200_job_* below */
typedef struct bssm_b200_job bssm_b200_job;

enum {
  BSSM_B200_OK = 0,
  BSSM_B200_ERR_ARG = 1,
  BSSM_B200_ERR_NO_DEVICE = 2,
  BSSM_B200_ERR_CUDA = 3,
  BSSM_B200_ERR_UNSUPPORTED = 4
};

/* distribution codes, R/bootstrap_filter.R:106-109 */
enum {
  BSSM_DIST_SVM = 0, BSSM_DIST_POISSON = 1, BSSM_DIST_BINOMIAL = 2,
  BSSM_DIST_NEGBIN = 3, BSSM_DIST_GAMMA = 4,
  /* univariate linear-Gaussian model (ssm_ulg / bsm_lg / ar1_lg): the `gaussian = TRUE`
   * branch of bsf / bsf_smoother (src/R_bootstrap_filter.cpp:24-105) */
  BSSM_DIST_GAUSSIAN = 5
};
/* sampling_method, R/run_mcmc.R:452-454 */
enum { BSSM_METHOD_PSI = 1, BSSM_METHOD_BSF = 2, BSSM_METHOD_EKPF = 4 }; /* sampling_method codes of R/run_mcmc.R */
/* post-processing a job runs after the filter */
enum {
  BSSM_POST_NONE = 0,           /* filter only */
  BSSM_POST_FILTER_SUMMARY = 1, /* + filter_summary: at, att, Pt, Ptt (src/summary.cpp:42-79) */
  BSSM_POST_SMOOTHER = 2,       /* + filter_smoother backtrack + summary: alphahat, Vt */
  /* used by the post-correction driver (src/approx_mcmc.cpp:388-414) */
  BSSM_POST_PICK = 3,           /* + one uniformly chosen smoothed trajectory per filter */
  BSSM_POST_SMOOTHER_MERGE = 4  /* + per-filter summaries kept for the weighted merge */
};
/* what a job keeps on the device / hands back */
enum {
  BSSM_OUT_ALPHA = 1,    /* particle cube m x (n+1) x N            */
  BSSM_OUT_WEIGHTS = 2,  /* N x (n+1), unnormalised exp(logw - max) */
  BSSM_OUT_INDICES = 4,  /* N x n uint32 ancestors (parity observable) */
  BSSM_OUT_ALL = 7
};

/* Univariate non-Gaussian observation / linear-Gaussian state model:
 * the fields the reference reads from the R list in
 * src/model_ssm_ung.cpp:9-32.  xbeta = xreg %*% beta (compute_xbeta,
 * src/model_ssm_ung.h:83) is formed by the caller; NULL = no regressors. */
typedef struct {
  int32_t n, m, k;
  int32_t distribution;
  int32_t Ztv, Ttv, Rtv, Dtv, Ctv; /* 1 = last dimension is n, 0 = constant */
  const double* y;     /* n, non-finite = missing */
  const double* Z;     /* m x (1|n) */
  const double* T;     /* m x m x (1|n) */
  const double* R;     /* m x k x (1|n) */
  const double* a1;    /* m */
  const double* P1;    /* m x m */
  const double* D;     /* (1|n) */
  const double* C;     /* m x (1|n) */
  const double* xbeta; /* n or NULL */
  const double* u;     /* n */
  double phi;
  /* Gaussian approximation controls (psi-APF only) */
  int32_t max_iter;
  int32_t local_approx;
  double conv_tol;
  const double* initial_mode; /* n */
  /* BSSM_DIST_GAUSSIAN only: observation standard deviation H (1|n), src/model_ssm_ulg.cpp:9-37 */
  const double* H;
  int32_t Htv;
} bssm_ung_model;

typedef struct {
  /* particle filter raw outputs */
  double* alpha;      /* m x (n+1) x N */
  double* weights;    /* N x (n+1) */
  uint32_t* indices;  /* N x n */
  /* filter_summary outputs (src/summary.cpp:42-79), bsf only */
  double* at;         /* m x (n+1) */
  double* att;        /* m x n */
  double* Pt;         /* m x m x (n+1) */
  double* Ptt;        /* m x m x n */
  /* smoother outputs (src/summary.cpp:4-17), *_smoother only */
  double* alphahat;   /* m x (n+1) */
  double* Vt;         /* m x m x (n+1) */
  /* scalars written by the callee */
  double loglik;
  double approx_loglik; /* psi only: log-likelihood of the Gaussian approximation */
  int32_t stopped_at;   /* -1, or the time index at which all weights vanished */
  float device_ms;      /* CUDA-event time of the kernels of this call */
} bssm_filter_out;

/* ---- library management ------------------------------------------------ */
const char* bssm_b200_last_error(void);
int bssm_b200_device_count(void);
int bssm_b200_set_device(int device);
/* frees cached device state; hook for .onUnload (reference: R/zzz.R:1-3) */
void bssm_b200_shutdown(void);
const char* bssm_b200_version(void);

/* ---- unit hook ---------------------------------------------------------- */
/* replaces: stratified_sample(p, r, N)  src/stratified_sample.cpp:9-28
 * (itself .Call-exported, R/RcppExports.R:228).  p (length n_p) is consumed as
 * the normalised weights; unlike the reference it is NOT mutated.  r: N
 * uniforms.  xp: N ancestor indices (0-based).  cumsum_mode: 0 = reproducible
 * fixed-point cumsum (what the filters use), see include/bssm_b200/pf_math.h. */
int bssm_b200_stratified_sample(const double* p, uint32_t n_p, const double* r,
                                uint32_t N, uint32_t* xp);

/* Test hooks: evaluate the deterministic math / random stream of
 * include/bssm_b200/pf_math.h ON THE DEVICE, so tests can check that host and
 * device agree bit for bit.  op: 0 exp, 1 log, 2 log1p, 3 sin(2 pi x), 4 cos(2 pi x),
 * 5 sin(x), 6 cos(x). */
int bssm_b200_math_probe(int op, const double* x, double* y, uint32_t n);
/* z: `count` normals of (seed, filter, t) in flat order; r: `count` resampling uniforms */
int bssm_b200_rng_probe(uint32_t seed, uint32_t filter, uint32_t t, uint32_t count,
                        double* z, double* r);

/* ---- ssm_ung family (ssm_ung / bsm_ng / svm / ar1_ng) -------------------- */
/* replaces: bsf(model_, nsim, seed, gaussian=FALSE, model_type)
 *           src/R_bootstrap_filter.cpp:15-241 -> ssm_ung::bsf_filter
 *           src/model_ssm_ung.cpp:611-712 + filter_summary src/summary.cpp:42-79 */
int bssm_b200_bsf(const bssm_ung_model* model, uint32_t nsim, uint32_t seed,
                  bssm_filter_out* out);
/* replaces: bsf_smoother(...) src/R_bootstrap_filter.cpp:245-457
 *           (bsf_filter + filter_smoother src/filter_smoother.cpp:5-18 +
 *            summary src/summary.cpp:4-17) */
int bssm_b200_bsf_smoother(const bssm_ung_model* model, uint32_t nsim, uint32_t seed,
                           bssm_filter_out* out);
/* replaces: psi_smoother(model_, nsim, seed, model_type) src/R_psi.cpp:52-171
 *           -> ssm_ung::psi_filter src/model_ssm_ung.cpp:511-609 (including the
 *           Gaussian approximation, smoother_ccov src/model_ssm_ulg.cpp:419-473
 *           and conditional_cov src/conditional_dist.cpp:4-32), then
 *           filter_smoother + summary.  m = 1 in this round. */
int bssm_b200_psi_smoother(const bssm_ung_model* model, uint32_t nsim, uint32_t seed,
                           bssm_filter_out* out);
/* replaces: gaussian_psi_smoother(model_, nsim, seed, model_type) src/R_psi.cpp:16-49 ->
 *           ssm_ulg::psi_filter src/model_ssm_ulg.cpp:646-672: nsim independent draws from the
 *           smoothing distribution of a linear-Gaussian model (no weights, no resampling).
 *           model->distribution must be BSSM_DIST_GAUSSIAN; alpha is m x (n+1) x nsim. */
int bssm_b200_gaussian_psi_smoother(const bssm_ung_model* model, uint32_t nsim, uint32_t seed,
                                    double* alpha);
/* importance_sample_ng(model_, nsim, use_antithetic, seed, model_type) for the ssm_ung family
 * (src/R_importance_sample.cpp:9-75, RcppExports.cpp:1300) and the "spdk" method of
 * nongaussian_loglik (src/R_loglik.cpp:41-87 -> ssm_ung::log_likelihood, src/model_ssm_ung.cpp:199-211):
 * nsim >= 2 draws of the states from the smoothing distribution of the approximating Gaussian model
 * (simulation smoother with one antithetic variable, src/model_ssm_ulg.cpp:141-217) and their
 * importance weights.  alpha (m x (n+1) x nsim, or NULL), weights (nsim, exp(w - max w), or NULL),
 * loglik (2 values or NULL: the SPDK estimate approx_loglik + log(mean(weights)) + max w, and
 * approx_loglik).  Like the reference's simulate_states() declaration the smoother always pairs the
 * draws when use_antithetic != 0 (draw i + nsim/2 mirrors draw i).  State dimension m <= 16. */
int bssm_b200_importance_sample_ung(const bssm_ung_model* model, uint32_t nsim, int use_antithetic,
                                    uint32_t seed, double* alpha, double* weights, double* loglik);

/* psi-APF filter only (no backtrack/summary): the body of
 * ssm_ung::log_likelihood(method=1) src/model_ssm_ung.cpp:164-236 */
int bssm_b200_psi_filter(const bssm_ung_model* model, uint32_t nsim, uint32_t seed,
                         bssm_filter_out* out);

/* ---- ssm_mng: multivariate observations ------------------------------------ */
/* The fields the reference reads in src/model_ssm_mng.cpp:7-30.  y and u are n x p
 * column-major exactly as R holds them (the reference transposes them on entry). */
typedef struct {
  int32_t n, m, k, p;
  int32_t Ztv, Ttv, Rtv, Dtv, Ctv;
  const int32_t* distribution; /* p: 1 poisson, 2 binomial, 3 negative binomial, 4 gamma, 5 gaussian */
  const double* y;     /* n x p */
  const double* u;     /* n x p */
  const double* Z;     /* p x m x (1|n) */
  const double* T;     /* m x m x (1|n) */
  const double* R;     /* m x k x (1|n) */
  const double* a1;    /* m */
  const double* P1;    /* m x m */
  const double* D;     /* p x (1|n) */
  const double* C;     /* m x (1|n) */
  const double* phi;   /* p */
  /* psi-APF only (src/model_ssm_mng.cpp:19-22): Laplace iteration control and start */
  int32_t max_iter, local_approx;
  double conv_tol;
  const double* initial_mode; /* n x p as R holds it, or NULL for the bootstrap filter */
} bssm_mng_model;

/* replaces: bsf(model_, nsim, seed, gaussian = FALSE, model_type = 0)
 *           src/R_bootstrap_filter.cpp:130-155 -> ssm_mng::bsf_filter
 *           src/model_ssm_mng.cpp:562-668 + filter_summary.  p <= 16, m <= 16. */
int bssm_b200_bsf_mng(const bssm_mng_model* model, uint32_t nsim, uint32_t seed, bssm_filter_out* out);
/* replaces: bsf_smoother(..., model_type = 0) src/R_bootstrap_filter.cpp:360-385 */
int bssm_b200_bsf_smoother_mng(const bssm_mng_model* model, uint32_t nsim, uint32_t seed,
                               bssm_filter_out* out);
int bssm_b200_job_create_mng(const bssm_mng_model* model, uint32_t nsim, uint32_t seed,
                             uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                             struct bssm_b200_job** job);
/* replaces: psi_smoother(model_, nsim, seed, model_type = 0) src/R_psi.cpp:57-78 ->
 *           ssm_mng::psi_filter src/model_ssm_mng.cpp:461-559 + filter_smoother + summary.
 *           m <= 16, p <= 16; needs initial_mode / max_iter / conv_tol. */
int bssm_b200_psi_smoother_mng(const bssm_mng_model* model, uint32_t nsim, uint32_t seed,
                               bssm_filter_out* out);
/* psi-APF filter only: ssm_mng::log_likelihood(method = 1) src/model_ssm_mng.cpp:138-172
 * (nongaussian_loglik, src/R_loglik.cpp:49-55) */
int bssm_b200_psi_filter_mng(const bssm_mng_model* model, uint32_t nsim, uint32_t seed,
                             bssm_filter_out* out);
int bssm_b200_job_create_mng_psi(const bssm_mng_model* model, uint32_t nsim, uint32_t seed,
                                 uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                                 struct bssm_b200_job** job);

/* ---- ssm_nlg and ssm_sde: device functor registry ------------------------- */
/* The reference defines these model classes through C++ function pointers that
 * R compiles at run time (externalptr arguments of bsf_nlg / bsf_sde,
 * src/R_bootstrap_filter.cpp:461-481, src/R_sde.cpp:35-49).  Host function
 * pointers cannot run in a kernel, so the shipped example models are device
 * policies selected by id; theta / known_params keep their meaning. */
enum {
  BSSM_NLG_GROWTH = 0, /* logistic growth model, R/cpp_example_models.R:336-511: m = k = 2, p = 1,
                        * theta = (log H, log R1, log R2), known_params = (dT, K, a11, a12, P11, P12) */
  BSSM_NLG_AR_EXP = 1  /* AR(1) state observed through exp(alpha), R/cpp_example_models.R:194-335:
                        * m = k = p = 1, theta = (mu, rho, log sigma, log H), no known_params */
  ,
  BSSM_NLG_SIN_EXP = 2 /* sine transition observed through exp(alpha), R/cpp_example_models.R:635-757:
                        * m = k = p = 1, theta = (log H, log R), a1 = 0, P1 = 1 */
  ,
  BSSM_NLG_LINEAR_GAUSSIAN = 3 /* random walk observed with noise, R/cpp_example_models.R:512-634:
                        * m = k = p = 1, theta = (log H), a1 = 0, P1 = 1, R = 1, Z(alpha) = T(alpha) = alpha */
};
enum {
  BSSM_SDE_GBM = 0,       /* geometric Brownian motion, R/cpp_example_models.R:109-193 */
  BSSM_SDE_POISSON_OU = 1 /* Ornstein-Uhlenbeck state, Poisson observations, R/cpp_example_models.R:24-108:
                           * theta = (log rho, nu, log sigma) */
};

typedef struct {
  int32_t model_id;           /* BSSM_NLG_* */
  int32_t n, p, m, k;         /* time points, series, states, etas */
  const double* y;            /* p x n (R passes t(model$y), R/bootstrap_filter.R:147) */
  const double* theta;        int32_t n_theta;
  const double* known_params; int32_t n_known;
} bssm_nlg_model;

typedef struct {
  int32_t model_id;           /* BSSM_SDE_* */
  int32_t n;
  const double* y;            /* n */
  const double* theta;        int32_t n_theta;
  double x0;
  int32_t positive;
} bssm_sde_model;

/* replaces: bsf_nlg(...) src/R_bootstrap_filter.cpp:461-507 -> ssm_nlg::bsf_filter
 *           src/model_ssm_nlg.cpp:1066-1146 + filter_summary */
int bssm_b200_bsf_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed, bssm_filter_out* out);
/* replaces: bsf_smoother_nlg(...) src/R_bootstrap_filter.cpp:509-551 */
int bssm_b200_bsf_smoother_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                               bssm_filter_out* out);
/* replaces: ekpf(...) and ekpf_smoother(...) src/R_ekpf.cpp:7-100 -> ssm_nlg::ekf_filter
 *           src/model_ssm_nlg.cpp:1150-1260 (EKF-based particle filter: every particle proposes from
 *           its own EKF update) + filter_summary / filter_smoother + summary.  Registry models. */
int bssm_b200_ekpf_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed, bssm_filter_out* out);
int bssm_b200_ekpf_smoother_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                                bssm_filter_out* out);
int bssm_b200_job_create_nlg_ekpf(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                                  uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                                  bssm_b200_job** job);

/* replaces: psi_smoother_nlg(..., nsim, seed, max_iter, conv_tol, iekf_iter) src/R_psi.cpp:173-221
 *           -> ssm_nlg::approximate + psi_filter (src/model_ssm_nlg.cpp:62-150, :971-1063),
 *           filter_smoother, summary.  p = 1 models of the registry; iekf_iter > 0 iterates the EKF start (src/model_ssm_nlg.cpp:325-356).
 * bssm_b200_psi_filter_nlg is the same without backtrack / summary. */
int bssm_b200_psi_smoother_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                               uint32_t max_iter, double conv_tol, uint32_t iekf_iter,
                               bssm_filter_out* out);
int bssm_b200_psi_filter_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                             uint32_t max_iter, double conv_tol, uint32_t iekf_iter,
                             bssm_filter_out* out);
int bssm_b200_job_create_nlg_psi(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                                 uint32_t max_iter, double conv_tol, uint32_t n_filters,
                                 uint32_t first_filter_id, int out_flags, int post,
                                 struct bssm_b200_job** job);
/* replaces: the compiled function pointers of an ssm_sde model -- drift, diffusion, ddiffusion,
 *           log_obs_density (src/model_ssm_sde.h:8-11; written by the user and compiled with
 *           Rcpp::sourceCpp, vignettes/sde_model.Rmd).  A host pointer cannot be called from a
 *           kernel: the four functions are given as CUDA C++ SOURCE with the scalar signatures
 *             __device__ double drift(double x, const double* theta);
 *             __device__ double diffusion(double x, const double* theta);
 *             __device__ double ddiffusion(double x, const double* theta);
 *             __device__ double log_obs_density(double y, double x, const double* theta);
 *           and compiled at run time (NVRTC) into the filter kernels the registry models use.
 *           The returned id (>= BSSM_USER_MODEL_BASE) goes into bssm_sde_model.model_id; theta
 *           then has any length >= 1.  The source may call the deterministic math of
 *           include/bssm_b200/pf_math.h (pf_exp, pf_log, pf_log1p, pf_sin, pf_cos).  A source that
 *           does not compile returns BSSM_B200_ERR_ARG with the compiler log in last_error. */
#define BSSM_USER_MODEL_BASE 1000
int bssm_b200_sde_register(const char* cuda_source, int32_t* model_id);
/* replaces: the function pointers of an ssm_nlg model that the bootstrap filter calls -- Z_fn, H_fn,
 *           T_fn, R_fn, a1_fn, P1_fn (src/model_ssm_nlg.h:10-21; arma::vec / arma::mat returns),
 *           as CUDA C++ source with the outputs written through a pointer (p = 1, 1 <= k <= m <= 4):
 *             __device__ double Z_fn(int t, const double* alpha, const double* theta, const double* known_params);
 *             __device__ double H_fn(int t, const double* alpha, const double* theta, const double* known_params); // sd
 *             __device__ void T_fn(int t, const double* alpha, const double* theta, const double* known_params, double* out);
 *             __device__ void R_fn(int t, const double* alpha, const double* theta, const double* known_params, double* out);
 *             __device__ void a1_fn(const double* theta, const double* known_params, double* out);
 *             __device__ void P1_fn(const double* theta, const double* known_params, double* out);
 *           (matrices column-major; known_tv_params go to the tail of known_params, indexed with t).
 *           The id is valid for bssm_b200_bsf_nlg, bssm_b200_bsf_smoother_nlg, bssm_b200_job_create_nlg
 *           (+ job_set_theta_nlg) and bssm_b200_is_correction_bsf_nlg; the psi-APF and the EKF particle
 *           filter need the host-side Gaussian approximation and stay with the registry models. */
int bssm_b200_nlg_register(const char* cuda_source, int32_t m, int32_t k, int32_t* model_id);

/* replaces: bsf_sde(y, x0, positive, ..., theta, nsim, L, seed) src/R_sde.cpp:35-72
 *           -> ssm_sde::bsf_filter src/model_ssm_sde.cpp:21-82 (Milstein, 2^L sub-steps).
 * Like the reference, att and Ptt have n+1 columns here with the last one zero. */
int bssm_b200_bsf_sde(const bssm_sde_model* model, uint32_t nsim, uint32_t L, uint32_t seed,
                      bssm_filter_out* out);
/* replaces: bsf_smoother_sde(...) src/R_sde.cpp:75-111 */
int bssm_b200_bsf_smoother_sde(const bssm_sde_model* model, uint32_t nsim, uint32_t L, uint32_t seed,
                               bssm_filter_out* out);
/* replaces: loglik_sde(...) src/R_sde.cpp:10-32 (log-likelihood only) */
int bssm_b200_loglik_sde(const bssm_sde_model* model, uint32_t nsim, uint32_t L, uint32_t seed,
                         double* loglik);
/* resident jobs of replicate filters, as bssm_b200_job_create_ung */
int bssm_b200_job_create_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                             uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                             struct bssm_b200_job** job);
int bssm_b200_job_create_sde(const bssm_sde_model* model, uint32_t nsim, uint32_t L, uint32_t seed,
                             uint32_t n_filters, uint32_t first_filter_id, int out_flags, int post,
                             struct bssm_b200_job** job);

/* Mode of the Gaussian approximation: the Newton / Laplace iterations of
 * ssm_ung::approximate (src/model_ssm_ung.cpp:87-140) started from
 * model->initial_mode, m = 1.  This is what run_mcmc stores per draw in
 * mcmc_output$modes and post_correct consumes.  Host computation (O(n) per
 * sweep, sequential in t); exported so callers can build post-correction
 * inputs without R. */
int bssm_b200_approximate_mode(const bssm_ung_model* model, double* mode);

/* ---- IS-MCMC post-correction ---------------------------------------------- */
/* How theta maps onto the model (the C++ update_model of the model class):
 * svm src/model_svm.cpp:12-29, ar1_ng src/model_ar1_ng.cpp:12-36,
 * bsm_ng src/model_bsm_ng.cpp:17-51 (any m <= 16: seasonal components included). */
typedef struct {
  int32_t model_type; /* R/model_type.R:1-16: 2 = bsm_ng, 3 = svm, 4 = ar1_ng */
  int32_t svm_type;   /* svm: 0 = theta[2] is sigma (phi), 1 = theta[2] is mu */
  int32_t mu_est;     /* ar1_ng */
  int32_t phi_est;    /* ar1_ng, bsm_ng */
  /* bsm_ng (src/model_bsm_ng.cpp:4-51): which components exist and which
   * standard deviations are in theta (fixed(i) == 0) */
  int32_t slope, seasonal, noise;
  int32_t level_est, slope_est, seasonal_est;
  /* bsm_ng and ar1_ng: the regression coefficients are the last n_xreg entries
   * of theta; update_model recomputes xbeta = xreg * beta for every theta
   * (src/model_bsm_ng.cpp:46-49, src/model_ar1_ng.cpp:28-31) */
  int32_t n_xreg;
  const double* xreg; /* n x n_xreg column-major, or NULL */
} bssm_theta_map;

typedef struct {
  double* weights;    /* n_draws: exp(loglik_i), src/approx_mcmc.cpp:385 */
  double* posterior;  /* n_draws: prior (0) + approx_loglik (0) + log(weight), :425-426 */
  double* alpha;      /* output_type 1: (n+1) x m x n_draws, one trajectory per draw */
  double* alphahat;   /* output_type 2: (n+1) x m */
  double* Vt;         /* output_type 2: m x m x (n+1) */
  double* loglik;     /* optional, n_draws: the psi-APF log-likelihood estimates */
  double sum_w;       /* output_type 2: sum of w_i * count_i over the draws of this call */
  float device_ms;    /* CUDA-event time of all kernels of this call */
  int32_t launches;
} bssm_postcorr_out;

/* replaces: postcorrection_nongaussian(model_, model_type, output_type, nsim,
 *           seed, n_threads, is_type, counts, theta, modes)
 *           src/R_postcorrection.cpp:150-248 -> approx_mcmc::is_correction_psi
 *           src/approx_mcmc.cpp:353-427 (+ expand :49-88 for is_type 3).
 * theta is n_par x n_stored, modes is n x n_stored (p = 1), counts n_stored.
 * One independent psi-APF per (expanded) draw; draw d uses random stream
 * `filter = d`, so results do not depend on how draws are sharded -- unlike the
 * reference, whose per-thread engines make results depend on `threads`
 * (src/approx_mcmc.cpp:367).  n_threads is the number of GPUs of this process to use (the
 * reference's OpenMP fan-out, src/approx_mcmc.cpp:364-369, becomes one host thread per device):
 * contiguous static shards of the draw list, then ONE ncclAllGather of the per-draw
 * log-likelihoods over a communicator created once per process (bssm_b200_last_gather() says how
 * the last call gathered).  n_threads <= 1, one visible device or output_type 2 (the summary merge
 * is a sequential recursion over the draws) run on the current device.  [draw_begin, draw_end)
 * selects the shard of the (expanded, for is_type 3) draw list this call computes -- 0, 0 = all
 * draws; outputs are indexed from draw_begin.  One process per GPU (torchrun) works too: each
 * rank calls this with its shard, then one gather of `weights` (bssm_b200/distributed.py). */
const char* bssm_b200_last_gather(void);
/* destroys the persistent NCCL communicator (also done by bssm_b200_shutdown) */
void bssm_b200_release_comm(void);
int bssm_b200_postcorrection_nongaussian(const bssm_ung_model* model, const bssm_theta_map* map,
                                         int output_type, uint32_t nsim, uint32_t seed,
                                         uint32_t n_threads, int is_type, const uint32_t* counts,
                                         const double* theta, uint32_t n_par, const double* modes,
                                         uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                         bssm_postcorr_out* out);

/* ---- models whose update_fn lives in R: pre-evaluated parameter sets ---------- */
/* parset_ung / parset_mng (src/parset_ng.h:11-62, src/parset_ng.cpp:5-215).  ssm_ung and ssm_mng
 * models map theta onto the model through an R closure (update_fn); the reference evaluates it
 * for every stored draw BEFORE its OpenMP region and keeps the results in a parset, and
 * parset::update(model, i) swaps them in per draw.  This struct is that parset: for each
 * component that update_fn returns, n_stored consecutive slices, each with the shape of the
 * model's own array (time-varying or not as the model says); NULL = not returned by update_fn
 * (the model's own array is used for every draw).
 *   ssm_ung: phi n_stored; Z m x (1|n); T m x m x (1|n); R m x k x (1|n); a1 m; P1 m x m;
 *            D (1|n); C m x (1|n); beta n_xreg (then xreg n x n_xreg is needed: compute_xbeta)
 *   ssm_mng: phi p; Z p x m x (1|n); T, R, a1, P1, C as above; D p x (1|n); no beta */
typedef struct {
  const double* phi;
  const double* Z;
  const double* T;
  const double* R;
  const double* a1;
  const double* P1;
  const double* D;
  const double* C;
  const double* beta;
  const double* xreg; /* n x n_xreg column-major, with beta */
  int32_t n_xreg;
} bssm_parset;

/* replaces: postcorrection_nongaussian(..., model_type = 1) -> approx_mcmc::is_correction_psi<ssm_ung>
 *           (src/approx_mcmc.cpp:685-804) and, with approx_loglik instead of modes, the ssm_ung
 *           specialisation of is_correction_bsf (:806-921).  Any m <= 16.  modes is n x n_stored.  Other arguments
 *           and outputs as bssm_b200_postcorrection_nongaussian. */
int bssm_b200_postcorrection_ung_parset(const bssm_ung_model* model, const bssm_parset* parset,
                                        int output_type, uint32_t nsim, uint32_t seed, uint32_t n_threads,
                                        int is_type, const uint32_t* counts, const double* modes,
                                        uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                        bssm_postcorr_out* out);
int bssm_b200_is_correction_bsf_ung_parset(const bssm_ung_model* model, const bssm_parset* parset,
                                           int output_type, uint32_t nsim, uint32_t seed, uint32_t n_threads,
                                           int is_type, const uint32_t* counts, const double* approx_loglik,
                                           uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                           bssm_postcorr_out* out);
/* approx_mcmc::is_correction_spdk (src/approx_mcmc.cpp:606-678, ssm_ung with parameter sets :923-1036; run_mcmc(mcmc_type = "is*",
 * sampling_method = "spdk")) for ssm_ung models whose stored draws are given as the reference's own parameter
 * sets (bssm_parset = parset_ung, src/parset_ng.h:11-40), output_type 3: for every stored draw the model,
 * its approximation at the stored mode (modes: n x n_stored) and nsim (is_type 1: nsim * counts[i]) simulated
 * trajectories; weights[i] = mean(exp(w - sum(scales))) (:637-639, :959-961), posterior[i] = approx_loglik[i] (or 0 when
 * NULL) + log(weights[i]) (:1034-1035).  is_type 3 expands the draws by their counts first (expand(), :49-88):
 * weights / posterior then hold sum(counts) values, else n_stored.  Draw i uses the random stream i. */
int bssm_b200_is_correction_spdk_ung_parset(const bssm_ung_model* model, const bssm_parset* parset,
                                            uint32_t nsim, uint32_t seed, int is_type, const uint32_t* counts,
                                            const double* modes, const double* approx_loglik, uint32_t n_stored,
                                            double* weights, double* posterior);

/* replaces: postcorrection_nongaussian(..., model_type = 0) -> approx_mcmc::is_correction_psi<ssm_mng>
 *           (src/approx_mcmc.cpp:1038-1157) and is_correction_bsf<ssm_mng> (:1159-1273).
 *           modes is p x n x n_stored (mode_storage, src/approx_mcmc.cpp:29).  m <= 16, p <= 16. */
int bssm_b200_postcorrection_mng_parset(const bssm_mng_model* model, const bssm_parset* parset,
                                        int output_type, uint32_t nsim, uint32_t seed, uint32_t n_threads,
                                        int is_type, const uint32_t* counts, const double* modes,
                                        uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                        bssm_postcorr_out* out);
int bssm_b200_is_correction_bsf_mng_parset(const bssm_mng_model* model, const bssm_parset* parset,
                                           int output_type, uint32_t nsim, uint32_t seed, uint32_t n_threads,
                                           int is_type, const uint32_t* counts, const double* approx_loglik,
                                           uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                           bssm_postcorr_out* out);

/* replaces: the draw loop of approx_mcmc::is_correction_bsf (src/approx_mcmc.cpp:447-518), the
 *           correction phase of run_mcmc(..., sampling_method = "bsf") for svm / ar1_ng / bsm_ng
 *           (level only): one bootstrap filter per stored draw at that draw's theta;
 *           weight_i = exp(loglik_i - approx_loglik_i), posterior_i = log weight_i (caller adds the
 *           prior).  m = 1.  Arguments as bssm_b200_postcorrection_nongaussian, with
 *           approx_loglik (n_stored) in place of the modes. */
int bssm_b200_is_correction_bsf(const bssm_ung_model* model, const bssm_theta_map* map, int output_type,
                                uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                const uint32_t* counts, const double* theta, uint32_t n_par,
                                const double* approx_loglik, uint32_t n_stored, uint32_t draw_begin,
                                uint32_t draw_end, bssm_postcorr_out* out);

/* replaces: the draw loop of approx_mcmc::is_correction_bsf<ssm_nlg> (src/approx_mcmc.cpp:442-518), the
 *           correction phase of run_mcmc(..., sampling_method = "bsf") for ssm_nlg models
 *           (src/R_mcmc.cpp:790): one bootstrap filter per stored draw after update_model(theta_i);
 *           weight_i = exp(loglik_i - approx_loglik_i), posterior_i = log weight_i (the caller adds
 *           the prior).  Registry models (model->model_id), m = 1 or 2.  theta is n_par x n_stored;
 *           other arguments and outputs as bssm_b200_is_correction_bsf; draws shard the same way. */
int bssm_b200_is_correction_bsf_nlg(const bssm_nlg_model* model, int output_type, uint32_t nsim,
                                    uint32_t seed, uint32_t n_threads, int is_type, const uint32_t* counts,
                                    const double* theta, uint32_t n_par, const double* approx_loglik,
                                    uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                    bssm_postcorr_out* out);

/* replaces: the draw loop of approx_mcmc::is_correction_bsf<ssm_sde> (src/approx_mcmc.cpp:522-591),
 *           the correction phase of sde_is_mcmc (src/R_sde.cpp:221-282): one bootstrap filter per
 *           stored draw on the FINE mesh L_f at that draw's theta;
 *           weight_i = exp(loglik_i - approx_loglik_i), posterior_i = approx_loglik_i + log weight_i
 *           (the caller adds the prior).  theta is n_par x n_stored, approx_loglik has n_stored
 *           entries (the coarse-mesh estimates stored by the MCMC phase).  Other arguments and
 *           outputs as bssm_b200_postcorrection_nongaussian; draws shard the same way. */
int bssm_b200_is_correction_bsf_sde(const bssm_sde_model* model, uint32_t L_f, int output_type,
                                    uint32_t nsim, uint32_t seed, uint32_t n_threads, int is_type,
                                    const uint32_t* counts, const double* theta, uint32_t n_par,
                                    const double* approx_loglik, uint32_t n_stored,
                                    uint32_t draw_begin, uint32_t draw_end, bssm_postcorr_out* out);

/* replaces: suggest_n_nongaussian(model_, theta, candidates, replications, seed, model_type)
 *           src/R_postcorrection.cpp:14-104: for every candidate particle count, the standard
 *           deviation (arma::stddev, N - 1) of `replications` psi-APF log-likelihood estimates at
 *           theta.  One device batch of replicate filters per candidate; replicate j of candidate
 *           c uses stream c * replications + j.  Models: as bssm_b200_postcorrection_nongaussian. */
int bssm_b200_suggest_n_nongaussian(const bssm_ung_model* model, const bssm_theta_map* map,
                                    const double* theta, uint32_t n_par, const uint32_t* candidates,
                                    uint32_t n_candidates, uint32_t replications, uint32_t seed,
                                    double* sds);
/* the same for model_type 0 (ssm_mng, src/R_postcorrection.cpp:21-36).  ssm_ung / ssm_mng map theta onto
 * the model through an R closure (model.update_model(theta, update_fn)): the caller applies it and passes the
 * model AT theta -- for ssm_ung through bssm_b200_suggest_n_nongaussian with map->model_type = 1 (theta is
 * then not read), for ssm_mng here. */
int bssm_b200_suggest_n_mng(const bssm_mng_model* model, const uint32_t* candidates, uint32_t n_candidates,
                            uint32_t replications, uint32_t seed, double* sds);
/* replaces: suggest_n_nonlinear(...) src/R_postcorrection.cpp:107-147 (registry models) */
int bssm_b200_suggest_n_nonlinear(const bssm_nlg_model* model, const double* theta, uint32_t n_par,
                                  const uint32_t* candidates, uint32_t n_candidates,
                                  uint32_t replications, uint32_t seed, uint32_t max_iter,
                                  double conv_tol, double* sds);

/* replaces: postcorrection_nonlinear(y, Z, H, T, R, Zg, Tg, a1, P1, theta, ..., output_type, nsim,
 *           seed, n_threads, is_type, counts, theta, modes) src/R_postcorrection.cpp:252-317 ->
 *           is_correction_psi<ssm_nlg> (src/approx_mcmc.cpp:353-427); modes is m x n x n_stored.
 * Registry models with p = 1 (growth).  The per-draw Gaussian approximation
 * (approximate_for_is + Kalman smoother, O(n m^3) per draw) runs on the host with OpenMP
 * in this build; the particle filters of all draws run as one device batch. */
int bssm_b200_postcorrection_nonlinear(const bssm_nlg_model* model, int output_type, uint32_t nsim,
                                       uint32_t seed, uint32_t n_threads, int is_type,
                                       const uint32_t* counts, const double* theta, uint32_t n_par,
                                       const double* modes, uint32_t n_stored, uint32_t draw_begin,
                                       uint32_t draw_end, bssm_postcorr_out* out);

/* ---- resident jobs: the same filters with inputs kept in HBM ------------- */
/* A job owns device copies of the model, the pre-pass results and the output
 * buffers of `n_filters` independent replicate filters (filter f uses random
 * stream `filter = first_filter_id + f`, see pf_math.h).  run() only launches
 * kernels; nothing crosses PCIe.  This is what bench.py times for `value`. */

int bssm_b200_job_create_ung(const bssm_ung_model* model, int method, uint32_t nsim,
                             uint32_t seed, uint32_t n_filters, uint32_t first_filter_id,
                             int out_flags, int post, bssm_b200_job** job);
/* a resident batch of post-correction draws (is_type 2 / 3 only: equal N): the
 * draws [draw_begin, draw_end) of the expanded list, weights only
 * (output_type 3).  bench.py times run() of this job for the draws/s metric. */
int bssm_b200_job_create_postcorr(const bssm_ung_model* model, const bssm_theta_map* map,
                                  uint32_t nsim, uint32_t seed, int is_type, const uint32_t* counts,
                                  const double* theta, uint32_t n_par, const double* modes,
                                  uint32_t n_stored, uint32_t draw_begin, uint32_t draw_end,
                                  bssm_b200_job** job);
/* the same for ssm_nlg (growth): draws [draw_begin, draw_end), out_flags as in job_create_ung */
int bssm_b200_job_create_postcorr_nlg(const bssm_nlg_model* model, uint32_t nsim, uint32_t seed,
                                      int is_type, const uint32_t* counts, const double* theta,
                                      uint32_t n_par, const double* modes, uint32_t n_stored,
                                      uint32_t draw_begin, uint32_t draw_end, int out_flags,
                                      bssm_b200_job** job);
/* launches pre-pass + filter (+ the post-processing kernels) on the
 * job's stream and waits; ms = CUDA-event time on that stream */
/* The particle filter inside an MCMC loop (pm_mcmc / da_mcmc, src/mcmc.cpp:338-491: one filter
 * per iteration at a new theta): re-parameterise a resident job in place.  update_model(theta)
 * and the Newton sweeps of approximate() run on the host, only the mode (n doubles) and six
 * scalars are uploaded, the next job_run redoes the device pre-pass and all replicate filters.
 * `model` is the model the job was created from; m = 1 models (svm, ar1_ng, bsm_ng level). */
int bssm_b200_job_set_theta(bssm_b200_job* job, const bssm_ung_model* model, const bssm_theta_map* map,
                            const double* theta, uint32_t n_par);

/* The same for ssm_nlg (bootstrap / EKF particle filter / psi-APF jobs) and ssm_sde jobs: the filter that
 * nonlinear_pm_mcmc / sde_pm_mcmc call once per iteration at a new theta (src/R_mcmc.cpp:551-609 ->
 * mcmc::pm_mcmc src/mcmc.cpp:338-491; src/R_sde.cpp:114-166 -> the ssm_sde pm_mcmc src/mcmc.cpp:694-820,
 * `model.bsf_filter(nsim, L_f, ...)` at :710 and :759).  Bootstrap / EKF jobs take theta by value (no
 * transfer); psi-APF jobs redo approximate() on the host and refill the job's device buffers. */
int bssm_b200_job_set_theta_nlg(bssm_b200_job* job, const bssm_nlg_model* model, const double* theta,
                                uint32_t n_par);
int bssm_b200_job_set_theta_sde(bssm_b200_job* job, const double* theta, uint32_t n_par);
int bssm_b200_job_run(bssm_b200_job* job, float* ms);
/* per-kernel CUDA-event times of the last run: [prepass, filter, smoother] */
int bssm_b200_job_kernel_ms(const bssm_b200_job* job, float ms[3]);
/* how many filters of this job's geometry the device can run concurrently
 * (cudaOccupancyMaxActiveClusters for cluster launches): batch sizes that are a
 * multiple of it avoid a nearly empty last wave */
int bssm_b200_job_max_concurrent_filters(const bssm_b200_job* job);
/* diagnostics: SM cycles per phase of the filter kernel, accumulated over the
 * time loop by thread 0 of CTA `cta` during the last run (enabled by the
 * environment variable BSSM_B200_PHASES=1 at job creation; zeros otherwise):
 * [log-weights, max collective, exp + thresholds, scan collective,
 *  counts + normals + scatter + barrier, gather + propagate] */
int bssm_b200_job_phase_cycles(bssm_b200_job* job, uint32_t cta, uint64_t cycles[6]);
/* number of kernel launches one run() issues (for gpu_launches accounting) */
int bssm_b200_job_launches(const bssm_b200_job* job);
/* copies filter f's outputs to the host buffers in *out */
int bssm_b200_job_fetch(bssm_b200_job* job, uint32_t f, bssm_filter_out* out);
/* log-likelihoods of all filters, host copy */
int bssm_b200_job_logliks(bssm_b200_job* job, double* loglik);
/* the same into a caller-owned DEVICE buffer (e.g. a torch tensor handed to NCCL) */
int bssm_b200_job_logliks_dev(bssm_b200_job* job, double* loglik_dev);
void bssm_b200_job_destroy(bssm_b200_job* job);

#ifdef __cplusplus
}
#endif
#endif /* BSSM_B200_H */
