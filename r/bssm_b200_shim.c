/* bssm_b200_shim.c -- the `.Call` side of the drop-in boundary.
 *
 * NOT BUILT IN THIS REPOSITORY: it needs R's headers (Rinternals.h), which the
 * build image does not have; tests/test_abi.py only type-checks it with gcc
 * against declaration-only stand-ins (tests/mock_R/).  It is the glue a bssm
 * maintainer adds to reach libbssm_b200.so from R: the ssm_ung family, the
 * ssm_nlg / ssm_sde registry models (13 of the 13 entry points of SURVEY.md 8b;
 * gaussian_psi_smoother stays on bssm's Kalman path); everything it calls is the
 * plain C ABI of include/bssm_b200.h, which IS built and tested here (through
 * ctypes, tests/test_gpu_*.py).  Build inside an R installation with
 *   R CMD SHLIB bssm_b200_shim.c -I../include -L../bssm_b200/csrc -lbssm_b200
 *
 * Entry points mirror the reference's generated wrappers
 * (/root/reference/src/RcppExports.cpp:57-68 for bsf, :1287-1346 CallEntries):
 * same argument order as the R stubs in R/RcppExports.R:12-229, the model
 * passed as the same named list (src/model_ssm_ung.cpp:9-32).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "bssm_b200.h"

static SEXP list_get(SEXP list, const char* name) {
  SEXP names = getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  error("bssm_b200: model list has no element '%s'", name);
  return R_NilValue;
}

/* R list -> bssm_ung_model, zero copy (REAL() pointers; column-major as R stores them).
 * "time varying" = last dimension > 1, exactly src/model_ssm_ung.cpp:15-16. */
static void fill_ung(SEXP model, bssm_ung_model* M, double* xbeta_buf) {
  SEXP y = list_get(model, "y"), Z = list_get(model, "Z"), T = list_get(model, "T"), R = list_get(model, "R");
  SEXP a1 = list_get(model, "a1"), D = list_get(model, "D"), C = list_get(model, "C");
  SEXP xreg = list_get(model, "xreg"), beta = list_get(model, "beta");
  int n = (int)XLENGTH(y), m = (int)XLENGTH(a1);
  memset(M, 0, sizeof(*M));
  M->n = n; M->m = m;
  M->k = INTEGER(getAttrib(R, R_DimSymbol))[1];
  M->distribution = asInteger(list_get(model, "distribution"));
  M->Ztv = XLENGTH(Z) > m; M->Ttv = XLENGTH(T) > (R_xlen_t)m * m; M->Rtv = XLENGTH(R) > (R_xlen_t)m * M->k;
  M->Dtv = XLENGTH(D) > 1; M->Ctv = XLENGTH(C) > m;
  M->y = REAL(y); M->Z = REAL(Z); M->T = REAL(T); M->R = REAL(R); M->a1 = REAL(a1);
  M->P1 = REAL(list_get(model, "P1")); M->D = REAL(D); M->C = REAL(C);
  M->u = REAL(list_get(model, "u")); /* recycle u = 1 to length n on the R side (svm stores a scalar) */
  M->phi = asReal(list_get(model, "phi"));
  M->max_iter = asInteger(list_get(model, "max_iter"));
  M->conv_tol = asReal(list_get(model, "conv_tol"));
  M->local_approx = asLogical(list_get(model, "local_approx"));
  M->initial_mode = REAL(list_get(model, "initial_mode"));
  M->xbeta = NULL;
  if (XLENGTH(beta) > 0) { /* compute_xbeta, src/model_ssm_ung.h:83 */
    int q = (int)XLENGTH(beta);
    for (int t = 0; t < n; t++) {
      double s = 0;
      for (int j = 0; j < q; j++) s += REAL(xreg)[t + (R_xlen_t)n * j] * REAL(beta)[j];
      xbeta_buf[t] = s;
    }
    M->xbeta = xbeta_buf;
  }
}

static void check_rc(int rc) {
  if (rc != 0) error("bssm_b200: %s", bssm_b200_last_error()); /* C++ exception -> R error, as BEGIN_RCPP does */
}

/* Rcpp::List::containsElementNamed + operator[] */
static SEXP list_find(SEXP list, const char* name) {
  SEXP names = getAttrib(list, R_NamesSymbol);
  if (names == R_NilValue) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(list); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

/* R list -> bssm_mng_model (model_type 0), the fields of src/model_ssm_mng.cpp:7-30.  y, u and initial_mode are
 * n x p as R holds them (the reference transposes on entry; the library reads them as they are). */
static void fill_mng(SEXP model, bssm_mng_model* M) {
  SEXP y = list_get(model, "y"), Z = list_get(model, "Z"), T = list_get(model, "T"), R = list_get(model, "R");
  SEXP a1 = list_get(model, "a1"), D = list_get(model, "D"), C = list_get(model, "C");
  memset(M, 0, sizeof(*M));
  M->n = INTEGER(getAttrib(y, R_DimSymbol))[0]; M->p = INTEGER(getAttrib(y, R_DimSymbol))[1];
  M->m = (int)XLENGTH(a1); M->k = INTEGER(getAttrib(R, R_DimSymbol))[1];
  M->Ztv = XLENGTH(Z) > (R_xlen_t)M->p * M->m; M->Ttv = XLENGTH(T) > (R_xlen_t)M->m * M->m;
  M->Rtv = XLENGTH(R) > (R_xlen_t)M->m * M->k; M->Dtv = XLENGTH(D) > M->p; M->Ctv = XLENGTH(C) > M->m;
  /* R/post_correction.R:262-264 leaves doubles after pmatch(...) - 1: 1 poisson .. 5 gaussian */
  SEXP dist = PROTECT(coerceVector(list_get(model, "distribution"), INTSXP));
  int32_t* dcode = (int32_t*)R_alloc(M->p, sizeof(int32_t));
  for (int j = 0; j < M->p; j++) dcode[j] = INTEGER(dist)[j];
  UNPROTECT(1);
  M->distribution = dcode;
  M->y = REAL(y); M->u = REAL(list_get(model, "u")); M->Z = REAL(Z); M->T = REAL(T); M->R = REAL(R);
  M->a1 = REAL(a1); M->P1 = REAL(list_get(model, "P1")); M->D = REAL(D); M->C = REAL(C);
  M->phi = REAL(list_get(model, "phi"));
  M->max_iter = asInteger(list_get(model, "max_iter"));
  M->conv_tol = asReal(list_get(model, "conv_tol"));
  M->local_approx = asLogical(list_get(model, "local_approx"));
  M->initial_mode = REAL(list_get(model, "initial_mode"));
}

/* parset_ung / parset_mng constructors (src/parset_ng.cpp:5-82, :127-183): update_fn(theta_i) for every stored
 * draw, on R's main thread and before any device work, exactly where the reference evaluates it (before its
 * OpenMP region, src/approx_mcmc.cpp:698, :1051).  Components update_fn does not return stay NULL. */
static void build_parset(SEXP model, SEXP theta_, int n_par, int n_stored, int p_phi, bssm_parset* ps) {
  static const char* names[9] = {"phi", "Z", "T", "R", "a1", "P1", "D", "C", "beta"};
  const double** slots[9] = {&ps->phi, &ps->Z, &ps->T, &ps->R, &ps->a1, &ps->P1, &ps->D, &ps->C, &ps->beta};
  double* bufs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  R_xlen_t sizes[9];
  SEXP update_fn = list_get(model, "update_fn");
  memset(ps, 0, sizeof(*ps));
  for (int c = 0; c < 9; c++) sizes[c] = c == 0 ? p_phi : (list_find(model, names[c]) == R_NilValue ? 0 : XLENGTH(list_get(model, names[c])));
  for (int i = 0; i < n_stored; i++) {
    SEXP th = PROTECT(allocVector(REALSXP, n_par));
    memcpy(REAL(th), REAL(theta_) + (R_xlen_t)n_par * i, sizeof(double) * n_par);
    SEXP call = PROTECT(lang2(update_fn, th));
    SEXP res = PROTECT(eval(call, R_GlobalEnv));
    for (int c = 0; c < 9; c++) {
      SEXP e = list_find(res, names[c]);
      if (e == R_NilValue) continue;
      if (XLENGTH(e) != sizes[c]) error("bssm_b200: update_fn returned '%s' with a different size than the model's", names[c]);
      if (!bufs[c]) bufs[c] = (double*)R_alloc((size_t)sizes[c] * n_stored, sizeof(double));
      SEXP d = PROTECT(coerceVector(e, REALSXP));
      memcpy(bufs[c] + sizes[c] * i, REAL(d), sizeof(double) * sizes[c]);
      UNPROTECT(1);
    }
    UNPROTECT(3);
  }
  for (int c = 0; c < 9; c++) *slots[c] = bufs[c];
  if (ps->beta) {
    SEXP xreg = list_get(model, "xreg");
    ps->xreg = REAL(xreg);
    ps->n_xreg = (int)sizes[8];
  }
}

/* model_type 0 (ssm_mng) of bsf / bsf_smoother / psi_smoother / nongaussian_loglik
 * (src/R_bootstrap_filter.cpp:130-155, :366-386, src/R_psi.cpp:60-77, src/R_loglik.cpp:49-55).
 * which: 0 bsf (filter summaries), 1 bsf_smoother, 2 psi_smoother, 3 psi log-likelihood, 4 bsf log-likelihood */
static SEXP mng_call(SEXP model_, unsigned N, unsigned seed, int which) {
  bssm_mng_model G;
  fill_mng(model_, &G);
  const int n = G.n, m = G.m;
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  if (which >= 3) {
    check_rc(which == 3 ? bssm_b200_psi_filter_mng(&G, N, seed, &out) : bssm_b200_bsf_mng(&G, N, seed, &out));
    return ScalarReal(out.loglik);
  }
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  out.weights = REAL(w); out.alpha = REAL(alpha);
  SEXP res;
  if (which == 0) {
    SEXP at = PROTECT(allocMatrix(REALSXP, m, n + 1)), att = PROTECT(allocMatrix(REALSXP, m, n));
    SEXP Pt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1)), Ptt = PROTECT(alloc3DArray(REALSXP, m, m, n));
    out.at = REAL(at); out.att = REAL(att); out.Pt = REAL(Pt); out.Ptt = REAL(Ptt);
    check_rc(bssm_b200_bsf_mng(&G, N, seed, &out));
    const char* nm[] = {"at", "att", "Pt", "Ptt", "weights", "logLik", "alpha", ""};
    res = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(res, 0, at); SET_VECTOR_ELT(res, 1, att); SET_VECTOR_ELT(res, 2, Pt); SET_VECTOR_ELT(res, 3, Ptt);
    SET_VECTOR_ELT(res, 4, w); SET_VECTOR_ELT(res, 5, ScalarReal(out.loglik)); SET_VECTOR_ELT(res, 6, alpha);
    UNPROTECT(7);
  } else {
    SEXP ah = PROTECT(allocMatrix(REALSXP, m, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1));
    out.alphahat = REAL(ah); out.Vt = REAL(Vt);
    check_rc(which == 1 ? bssm_b200_bsf_smoother_mng(&G, N, seed, &out) : bssm_b200_psi_smoother_mng(&G, N, seed, &out));
    const char* nm[] = {"alphahat", "Vt", "weights", "logLik", "alpha", ""};
    res = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(res, 0, ah); SET_VECTOR_ELT(res, 1, Vt); SET_VECTOR_ELT(res, 2, w);
    SET_VECTOR_ELT(res, 3, ScalarReal(out.loglik)); SET_VECTOR_ELT(res, 4, alpha);
    UNPROTECT(5);
  }
  if (!R_FINITE(out.loglik)) warning("Particle filtering stopped prematurely due to nonfinite log-likelihood.");
  return res;
}

/* bsf(model_, nsim, seed, gaussian, model_type): src/R_bootstrap_filter.cpp:15-241 */
SEXP _bssmb200_bsf(SEXP model_, SEXP nsim_, SEXP seed_, SEXP gaussian_, SEXP model_type_) {
  if (asLogical(gaussian_)) error("bssm_b200: Gaussian models use bssm's own Kalman/BSF code path");
  if (asInteger(model_type_) == 0) return mng_call(model_, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), 0);
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y")), m = (int)XLENGTH(list_get(model_, "a1"));
  unsigned N = (unsigned)asInteger(nsim_);
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  SEXP at = PROTECT(allocMatrix(REALSXP, m, n + 1)), att = PROTECT(allocMatrix(REALSXP, m, n));
  SEXP Pt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1)), Ptt = PROTECT(alloc3DArray(REALSXP, m, m, n));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.at = REAL(at); out.att = REAL(att); out.Pt = REAL(Pt); out.Ptt = REAL(Ptt);
  out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf(&M, N, (unsigned)asInteger(seed_), &out));
  if (!R_FINITE(out.loglik)) /* src/R_bootstrap_filter.cpp:166-167 */
    warning("Particle filtering stopped prematurely due to nonfinite log-likelihood.");
  /* the reference returns at/att transposed (arma::inplace_trans): do t() on the R side */
  const char* nm[] = {"at", "att", "Pt", "Ptt", "weights", "logLik", "alpha", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, at); SET_VECTOR_ELT(res, 1, att); SET_VECTOR_ELT(res, 2, Pt);
  SET_VECTOR_ELT(res, 3, Ptt); SET_VECTOR_ELT(res, 4, w);
  SET_VECTOR_ELT(res, 5, ScalarReal(out.loglik)); SET_VECTOR_ELT(res, 6, alpha);
  UNPROTECT(7);
  return res;
}

/* psi_smoother(model_, nsim, seed, model_type): src/R_psi.cpp:52-171 */
SEXP _bssmb200_psi_smoother(SEXP model_, SEXP nsim_, SEXP seed_, SEXP model_type_) {
  if (asInteger(model_type_) == 0) return mng_call(model_, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), 2);
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y")), m = (int)XLENGTH(list_get(model_, "a1"));
  unsigned N = (unsigned)asInteger(nsim_);
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  SEXP ah = PROTECT(allocMatrix(REALSXP, m, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.alphahat = REAL(ah); out.Vt = REAL(Vt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_psi_smoother(&M, N, (unsigned)asInteger(seed_), &out));
  if (!R_FINITE(out.loglik))
    warning("Particle filtering stopped prematurely due to nonfinite log-likelihood.");
  const char* nm[] = {"alphahat", "Vt", "weights", "logLik", "alpha", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, ah); SET_VECTOR_ELT(res, 1, Vt); SET_VECTOR_ELT(res, 2, w);
  SET_VECTOR_ELT(res, 3, ScalarReal(out.loglik)); SET_VECTOR_ELT(res, 4, alpha);
  UNPROTECT(5);
  return res;
}

/* theta -> model mapping of the model classes whose update_model is C++ (R/model_type.R:1-16) */
static void fill_map(SEXP model, int model_type, bssm_theta_map* map) {
  memset(map, 0, sizeof(*map));
  map->model_type = model_type;
  if (model_type == 3) map->svm_type = asInteger(list_get(model, "svm_type"));
  if (model_type == 4) { map->mu_est = asLogical(list_get(model, "mu_est")); map->phi_est = asLogical(list_get(model, "phi_est")); }
  if (model_type == 2) { /* bsm_ng, src/model_bsm_ng.cpp:4-14 */
    SEXP fixed = list_get(model, "fixed");
    map->slope = asLogical(list_get(model, "slope")); map->seasonal = asLogical(list_get(model, "seasonal"));
    map->noise = asLogical(list_get(model, "noise")); map->phi_est = asLogical(list_get(model, "phi_est"));
    map->level_est = INTEGER(fixed)[0] == 0;
    map->slope_est = map->slope && INTEGER(fixed)[1] == 0;
    map->seasonal_est = map->seasonal && INTEGER(fixed)[2] == 0;
  }
  if (model_type == 2 || model_type == 4) { /* beta is the tail of theta: update_model -> compute_xbeta */
    SEXP xreg = list_get(model, "xreg");
    map->n_xreg = XLENGTH(xreg) > 0 ? INTEGER(getAttrib(xreg, R_DimSymbol))[1] : 0;
    map->xreg = map->n_xreg > 0 ? REAL(xreg) : NULL;
  }
}

/* bsf_smoother(model_, nsim, seed, gaussian, model_type): src/R_bootstrap_filter.cpp:245-457 */
SEXP _bssmb200_bsf_smoother(SEXP model_, SEXP nsim_, SEXP seed_, SEXP gaussian_, SEXP model_type_) {
  if (asLogical(gaussian_)) error("bssm_b200: Gaussian models use bssm's own Kalman/BSF code path");
  if (asInteger(model_type_) == 0) return mng_call(model_, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), 1);
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y")), m = (int)XLENGTH(list_get(model_, "a1"));
  unsigned N = (unsigned)asInteger(nsim_);
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  SEXP ah = PROTECT(allocMatrix(REALSXP, m, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.alphahat = REAL(ah); out.Vt = REAL(Vt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf_smoother(&M, N, (unsigned)asInteger(seed_), &out));
  if (!R_FINITE(out.loglik))
    warning("Particle filtering stopped prematurely due to nonfinite log-likelihood.");
  const char* nm[] = {"alphahat", "Vt", "weights", "logLik", "alpha", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, ah); SET_VECTOR_ELT(res, 1, Vt); SET_VECTOR_ELT(res, 2, w);
  SET_VECTOR_ELT(res, 3, ScalarReal(out.loglik)); SET_VECTOR_ELT(res, 4, alpha);
  UNPROTECT(5);
  return res;
}

/* nongaussian_loglik(model_, nsim, sampling_method, seed, model_type): src/R_loglik.cpp:41-87.
 * sampling_method 1 = psi-APF, 2 = bootstrap filter, 3 = SPDK (ssm_ung family); nothing but the scalar crosses PCIe. */
SEXP _bssmb200_nongaussian_loglik(SEXP model_, SEXP nsim_, SEXP sampling_method_, SEXP seed_, SEXP model_type_) {
  if (asInteger(model_type_) == 0) {
    int sm = asInteger(sampling_method_);
    if (sm != 1 && sm != 2) error("bssm_b200: sampling_method must be 1 (psi) or 2 (bsf); SPDK stays on bssm's CPU path");
    return mng_call(model_, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), sm == 1 ? 3 : 4);
  }
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y"));
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  int method = asInteger(sampling_method_);
  if (method == 1) check_rc(bssm_b200_psi_filter(&M, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), &out));
  else if (method == 2) check_rc(bssm_b200_bsf(&M, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_), &out));
  else if (method == 3) { /* SPDK, src/model_ssm_ung.cpp:199-211 */
    double ll[2];
    check_rc(bssm_b200_importance_sample_ung(&M, (unsigned)asInteger(nsim_), 1, (unsigned)asInteger(seed_), NULL, NULL, ll));
    return ScalarReal(ll[0]);
  } else error("bssm_b200: sampling_method must be 1 (psi), 2 (bsf) or 3 (spdk)");
  return ScalarReal(out.loglik);
}

/* importance_sample_ng(model_, nsim, use_antithetic, seed, model_type): src/R_importance_sample.cpp:9-75.
 * list(alpha = m x (n+1) x nsim, weights = nsim); the R caller does the aperm (R/importance_sample.R:68). */
SEXP _bssmb200_importance_sample_ng(SEXP model_, SEXP nsim_, SEXP use_antithetic_, SEXP seed_, SEXP model_type_) {
  if (asInteger(model_type_) == 0) error("bssm_b200: importance_sample of ssm_mng models stays on bssm's CPU path");
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y")), m = (int)XLENGTH(list_get(model_, "a1"));
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  const unsigned N = (unsigned)asInteger(nsim_);
  SEXP alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N)), w = PROTECT(allocVector(REALSXP, N));
  /* the reference's wrapper drops use_antithetic (simulate_states(nsim) pairs the draws, :29): so do we */
  (void)use_antithetic_;
  check_rc(bssm_b200_importance_sample_ung(&M, N, 1, (unsigned)asInteger(seed_), REAL(alpha), REAL(w), NULL));
  const char* nm[] = {"alpha", "weights", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, alpha); SET_VECTOR_ELT(res, 1, w);
  UNPROTECT(3);
  return res;
}

/* suggest_n_nongaussian(model_, theta, candidates, replications, seed, model_type): src/R_postcorrection.cpp:14-104 */
SEXP _bssmb200_suggest_n_nongaussian(SEXP model_, SEXP theta_, SEXP candidates_, SEXP replications_, SEXP seed_,
                                     SEXP model_type_) {
  const int model_type = asInteger(model_type_);
  int nc = (int)XLENGTH(candidates_);
  uint32_t* cand = (uint32_t*)R_alloc(nc, sizeof(uint32_t));
  for (int i = 0; i < nc; i++) cand[i] = (uint32_t)REAL(candidates_)[i]; /* arma::vec candidates */
  SEXP sds = PROTECT(allocVector(REALSXP, nc));
  const unsigned reps = (unsigned)asInteger(replications_), seed = (unsigned)asInteger(seed_);
  bssm_parset ps;
  if (model_type == 0) { /* ssm_mng: model.update_model(theta, update_fn), src/R_postcorrection.cpp:21-24 */
    bssm_mng_model G;
    fill_mng(model_, &G);
    build_parset(model_, theta_, (int)XLENGTH(theta_), 1, G.p, &ps); /* one "draw": the slices are the arrays */
    if (ps.phi) G.phi = ps.phi;
    if (ps.Z) G.Z = ps.Z;
    if (ps.T) G.T = ps.T;
    if (ps.R) G.R = ps.R;
    if (ps.a1) G.a1 = ps.a1;
    if (ps.P1) G.P1 = ps.P1;
    if (ps.D) G.D = ps.D;
    if (ps.C) G.C = ps.C;
    check_rc(bssm_b200_suggest_n_mng(&G, cand, (unsigned)nc, reps, seed, REAL(sds)));
    UNPROTECT(1);
    return sds;
  }
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y"));
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  bssm_theta_map map;
  if (model_type == 1) { /* ssm_ung with update_fn (:37-40): apply it here, the library gets the model at theta */
    memset(&map, 0, sizeof(map));
    map.model_type = 1;
    build_parset(model_, theta_, (int)XLENGTH(theta_), 1, 1, &ps);
    if (ps.phi) M.phi = ps.phi[0];
    if (ps.Z) M.Z = ps.Z;
    if (ps.T) M.T = ps.T;
    if (ps.R) M.R = ps.R;
    if (ps.a1) M.a1 = ps.a1;
    if (ps.P1) M.P1 = ps.P1;
    if (ps.D) M.D = ps.D;
    if (ps.C) M.C = ps.C;
    if (ps.beta) { /* compute_xbeta */
      for (int t = 0; t < n; t++) {
        double sum = 0;
        for (int j = 0; j < ps.n_xreg; j++) sum += ps.xreg[t + (R_xlen_t)n * j] * ps.beta[j];
        xb[t] = sum;
      }
      M.xbeta = xb;
    }
  } else {
    fill_map(model_, model_type, &map);
  }
  check_rc(bssm_b200_suggest_n_nongaussian(&M, &map, REAL(theta_), (unsigned)XLENGTH(theta_), cand, (unsigned)nc,
                                           reps, seed, REAL(sds)));
  UNPROTECT(1);
  return sds;
}

/* postcorrection_nongaussian(model_, model_type, output_type, nsim, seed, n_threads, is_type,
 *                            counts, theta, modes): src/R_postcorrection.cpp:150-248.
 * Returns what the reference returns per output_type (:226-245): alpha (n+1) x m x n_draws for 1,
 * alphahat / Vt for 2, weights and posterior always. */
SEXP _bssmb200_postcorrection_nongaussian(SEXP model_, SEXP model_type_, SEXP output_type_, SEXP nsim_,
                                          SEXP seed_, SEXP n_threads_, SEXP is_type_, SEXP counts_,
                                          SEXP theta_, SEXP modes_) {
  const int model_type = asInteger(model_type_);
  bssm_ung_model M;
  bssm_mng_model G;
  bssm_theta_map map;
  bssm_parset ps;
  int n, m = (int)XLENGTH(list_get(model_, "a1"));
  int n_stored = (int)XLENGTH(counts_), n_par = INTEGER(getAttrib(theta_, R_DimSymbol))[0];
  if (model_type == 0) { /* ssm_mng: y is n x p, update_fn in R (src/R_postcorrection.cpp:186-194) */
    fill_mng(model_, &G);
    n = G.n;
    build_parset(model_, theta_, n_par, n_stored, G.p, &ps);
  } else {
    n = (int)XLENGTH(list_get(model_, "y"));
    double* xb = (double*)R_alloc(n, sizeof(double));
    fill_ung(model_, &M, xb);
    if (model_type == 1) build_parset(model_, theta_, n_par, n_stored, 1, &ps); /* ssm_ung, :195-202 */
    else fill_map(model_, model_type, &map);
  }
  int is_type = asInteger(is_type_), output_type = asInteger(output_type_);
  R_xlen_t nd = 0;
  for (int i = 0; i < n_stored; i++) nd += is_type == 3 ? INTEGER(counts_)[i] : 1;
  SEXP w = PROTECT(allocVector(REALSXP, nd)), post = PROTECT(allocVector(REALSXP, nd));
  SEXP alpha = PROTECT(output_type == 1 ? alloc3DArray(REALSXP, n + 1, m, (int)nd) : allocVector(REALSXP, 0));
  SEXP ah = PROTECT(output_type == 2 ? allocMatrix(REALSXP, n + 1, m) : allocVector(REALSXP, 0));
  SEXP Vt = PROTECT(output_type == 2 ? alloc3DArray(REALSXP, m, m, n + 1) : allocVector(REALSXP, 0));
  bssm_postcorr_out out;
  memset(&out, 0, sizeof(out));
  out.weights = REAL(w); out.posterior = REAL(post);
  if (output_type == 1) out.alpha = REAL(alpha);
  if (output_type == 2) { out.alphahat = REAL(ah); out.Vt = REAL(Vt); }
  const unsigned nsim = (unsigned)asInteger(nsim_), seed = (unsigned)asInteger(seed_), thr = (unsigned)asInteger(n_threads_);
  const uint32_t* counts = (const uint32_t*)INTEGER(counts_);
  if (model_type == 0)
    check_rc(bssm_b200_postcorrection_mng_parset(&G, &ps, output_type, nsim, seed, thr, is_type, counts, REAL(modes_),
                                                 (unsigned)n_stored, 0, 0, &out));
  else if (model_type == 1)
    check_rc(bssm_b200_postcorrection_ung_parset(&M, &ps, output_type, nsim, seed, thr, is_type, counts, REAL(modes_),
                                                 (unsigned)n_stored, 0, 0, &out));
  else
    check_rc(bssm_b200_postcorrection_nongaussian(&M, &map, output_type, nsim, seed, thr, is_type, counts, REAL(theta_),
                                                  (unsigned)n_par, REAL(modes_), (unsigned)n_stored, 0, 0, &out));
  const char* nm[] = {"alpha", "alphahat", "Vt", "weights", "posterior", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, alpha); SET_VECTOR_ELT(res, 1, ah); SET_VECTOR_ELT(res, 2, Vt);
  SET_VECTOR_ELT(res, 3, w); SET_VECTOR_ELT(res, 4, post);
  UNPROTECT(6);
  return res;
}


/* ---------------------------------------------------------------- ssm_nlg / ssm_sde
 * The reference passes these model classes as compiled function pointers (externalptr arguments Z, H,
 * T, R, Zg, Tg, a1, P1, log_prior_pdf; drift / diffusion / ddiffusion / log_obs_density for ssm_sde).
 * Host function pointers cannot run in a kernel: the sibling R file (r/bssm_b200.R) puts the REGISTRY
 * ID of the shipped example model (BSSM_NLG_* / BSSM_SDE_*, include/bssm_b200.h) into the first
 * pointer argument and passes the others through untouched, so that the argument lists stay those of
 * R/RcppExports.R. */
static void fill_nlg(SEXP y, SEXP id_, SEXP theta, SEXP known_params, SEXP n_states, SEXP n_etas, bssm_nlg_model* M) {
  memset(M, 0, sizeof(*M));
  if (TYPEOF(id_) == STRSXP) {
    /* Z_fn, H_fn, T_fn, R_fn, a1_fn, P1_fn as CUDA source (one string) in place of the Z pointer: bootstrap
     * filter only, compiled at run time (bssm_b200_nlg_register caches by source text) */
    int32_t id = -1;
    check_rc(bssm_b200_nlg_register(CHAR(STRING_ELT(id_, 0)), asInteger(n_states), asInteger(n_etas), &id));
    M->model_id = id;
  } else if (TYPEOF(id_) == INTSXP || TYPEOF(id_) == REALSXP) {
    M->model_id = asInteger(id_);
  } else {
    error("bssm_b200: ssm_nlg needs a registry model id or the model's CUDA source in place of the Z pointer");
  }
  M->p = INTEGER(getAttrib(y, R_DimSymbol))[0]; /* R passes t(model$y): p x n */
  M->n = INTEGER(getAttrib(y, R_DimSymbol))[1];
  M->m = asInteger(n_states); M->k = asInteger(n_etas);
  M->y = REAL(y);
  M->theta = REAL(theta); M->n_theta = (int)XLENGTH(theta);
  M->known_params = REAL(known_params); M->n_known = (int)XLENGTH(known_params);
}
static SEXP filter_result(SEXP at, SEXP att, SEXP Pt, SEXP Ptt, SEXP w, double loglik, SEXP alpha) {
  const char* nm[] = {"at", "att", "Pt", "Ptt", "weights", "logLik", "alpha", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, at); SET_VECTOR_ELT(res, 1, att); SET_VECTOR_ELT(res, 2, Pt); SET_VECTOR_ELT(res, 3, Ptt);
  SET_VECTOR_ELT(res, 4, w); SET_VECTOR_ELT(res, 5, ScalarReal(loglik)); SET_VECTOR_ELT(res, 6, alpha);
  UNPROTECT(1);
  return res;
}
static SEXP smoother_result(SEXP ah, SEXP Vt, SEXP w, double loglik, SEXP alpha) {
  const char* nm[] = {"alphahat", "Vt", "weights", "logLik", "alpha", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, ah); SET_VECTOR_ELT(res, 1, Vt); SET_VECTOR_ELT(res, 2, w);
  SET_VECTOR_ELT(res, 3, ScalarReal(loglik)); SET_VECTOR_ELT(res, 4, alpha);
  UNPROTECT(1);
  return res;
}
static void warn_stop(double loglik) {
  if (!R_FINITE(loglik)) warning("Particle filtering stopped prematurely due to nonfinite log-likelihood.");
}

/* bsf_nlg(y, Z, H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf, known_params, known_tv_params, n_states,
 *         n_etas, time_varying, nsim, seed): src/R_bootstrap_filter.cpp:461-507 */
SEXP _bssmb200_bsf_nlg(SEXP y, SEXP Z, SEXP H, SEXP T, SEXP R, SEXP Zg, SEXP Tg, SEXP a1, SEXP P1, SEXP theta,
                       SEXP log_prior_pdf, SEXP known_params, SEXP known_tv_params, SEXP n_states, SEXP n_etas,
                       SEXP time_varying, SEXP nsim_, SEXP seed_) {
  bssm_nlg_model M;
  fill_nlg(y, Z, theta, known_params, n_states, n_etas, &M);
  const int n = M.n, m = M.m;
  unsigned N = (unsigned)asInteger(nsim_);
  SEXP at = PROTECT(allocMatrix(REALSXP, m, n + 1)), att = PROTECT(allocMatrix(REALSXP, m, n));
  SEXP Pt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1)), Ptt = PROTECT(alloc3DArray(REALSXP, m, m, n));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.at = REAL(at); out.att = REAL(att); out.Pt = REAL(Pt); out.Ptt = REAL(Ptt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf_nlg(&M, N, (unsigned)asInteger(seed_), &out));
  warn_stop(out.loglik);
  SEXP res = filter_result(at, att, Pt, Ptt, w, out.loglik, alpha);
  UNPROTECT(6);
  return res;
}

/* bsf_smoother_nlg(same 18 arguments): src/R_bootstrap_filter.cpp:509-551 */
SEXP _bssmb200_bsf_smoother_nlg(SEXP y, SEXP Z, SEXP H, SEXP T, SEXP R, SEXP Zg, SEXP Tg, SEXP a1, SEXP P1, SEXP theta,
                                SEXP log_prior_pdf, SEXP known_params, SEXP known_tv_params, SEXP n_states,
                                SEXP n_etas, SEXP time_varying, SEXP nsim_, SEXP seed_) {
  bssm_nlg_model M;
  fill_nlg(y, Z, theta, known_params, n_states, n_etas, &M);
  const int n = M.n, m = M.m;
  unsigned N = (unsigned)asInteger(nsim_);
  SEXP ah = PROTECT(allocMatrix(REALSXP, m, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.alphahat = REAL(ah); out.Vt = REAL(Vt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf_smoother_nlg(&M, N, (unsigned)asInteger(seed_), &out));
  warn_stop(out.loglik);
  SEXP res = smoother_result(ah, Vt, w, out.loglik, alpha);
  UNPROTECT(4);
  return res;
}

/* psi_smoother_nlg(the 16 model arguments, nsim, seed, max_iter, conv_tol, iekf_iter): src/R_psi.cpp:173-221 */
SEXP _bssmb200_psi_smoother_nlg(SEXP y, SEXP Z, SEXP H, SEXP T, SEXP R, SEXP Zg, SEXP Tg, SEXP a1, SEXP P1, SEXP theta,
                                SEXP log_prior_pdf, SEXP known_params, SEXP known_tv_params, SEXP n_states,
                                SEXP n_etas, SEXP time_varying, SEXP nsim_, SEXP seed_, SEXP max_iter_,
                                SEXP conv_tol_, SEXP iekf_iter_) {
  bssm_nlg_model M;
  fill_nlg(y, Z, theta, known_params, n_states, n_etas, &M);
  const int n = M.n, m = M.m;
  unsigned N = (unsigned)asInteger(nsim_);
  SEXP ah = PROTECT(allocMatrix(REALSXP, m, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, m, m, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, m, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.alphahat = REAL(ah); out.Vt = REAL(Vt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_psi_smoother_nlg(&M, N, (unsigned)asInteger(seed_), (unsigned)asInteger(max_iter_),
                                      asReal(conv_tol_), (unsigned)asInteger(iekf_iter_), &out));
  warn_stop(out.loglik);
  SEXP res = smoother_result(ah, Vt, w, out.loglik, alpha);
  UNPROTECT(4);
  return res;
}

/* postcorrection_nonlinear(the 16 model arguments, output_type, nsim, seed, n_threads, is_type, counts,
 *                          theta, modes): src/R_postcorrection.cpp:252-317 */
SEXP _bssmb200_postcorrection_nonlinear(SEXP y, SEXP Z, SEXP H, SEXP T, SEXP R, SEXP Zg, SEXP Tg, SEXP a1, SEXP P1,
                                        SEXP theta_init, SEXP log_prior_pdf, SEXP known_params, SEXP known_tv_params,
                                        SEXP n_states, SEXP n_etas, SEXP time_varying, SEXP output_type_, SEXP nsim_,
                                        SEXP seed_, SEXP n_threads_, SEXP is_type_, SEXP counts_, SEXP theta_,
                                        SEXP modes_) {
  bssm_nlg_model M;
  fill_nlg(y, Z, theta_init, known_params, n_states, n_etas, &M);
  const int n = M.n, m = M.m;
  int n_stored = (int)XLENGTH(counts_), n_par = INTEGER(getAttrib(theta_, R_DimSymbol))[0];
  int is_type = asInteger(is_type_), output_type = asInteger(output_type_);
  R_xlen_t nd = 0;
  for (int i = 0; i < n_stored; i++) nd += is_type == 3 ? INTEGER(counts_)[i] : 1;
  SEXP w = PROTECT(allocVector(REALSXP, nd)), post = PROTECT(allocVector(REALSXP, nd));
  SEXP alpha = PROTECT(output_type == 1 ? alloc3DArray(REALSXP, n + 1, m, (int)nd) : allocVector(REALSXP, 0));
  SEXP ah = PROTECT(output_type == 2 ? allocMatrix(REALSXP, n + 1, m) : allocVector(REALSXP, 0));
  SEXP Vt = PROTECT(output_type == 2 ? alloc3DArray(REALSXP, m, m, n + 1) : allocVector(REALSXP, 0));
  bssm_postcorr_out out;
  memset(&out, 0, sizeof(out));
  out.weights = REAL(w); out.posterior = REAL(post);
  if (output_type == 1) out.alpha = REAL(alpha);
  if (output_type == 2) { out.alphahat = REAL(ah); out.Vt = REAL(Vt); }
  check_rc(bssm_b200_postcorrection_nonlinear(&M, output_type, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_),
                                              (unsigned)asInteger(n_threads_), is_type, (const uint32_t*)INTEGER(counts_),
                                              REAL(theta_), (unsigned)n_par, REAL(modes_), (unsigned)n_stored, 0, 0, &out));
  const char* nm[] = {"alpha", "alphahat", "Vt", "weights", "posterior", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, alpha); SET_VECTOR_ELT(res, 1, ah); SET_VECTOR_ELT(res, 2, Vt);
  SET_VECTOR_ELT(res, 3, w); SET_VECTOR_ELT(res, 4, post);
  UNPROTECT(6);
  return res;
}

static void fill_sde(SEXP y, SEXP x0, SEXP positive, SEXP id_, SEXP theta, bssm_sde_model* M) {
  memset(M, 0, sizeof(*M));
  if (TYPEOF(id_) == STRSXP) {
    /* the four model functions as CUDA source (one string) in place of the drift pointer: compiled at run
     * time into the filter kernels (bssm_b200_sde_register caches by source text) */
    int32_t id = -1;
    check_rc(bssm_b200_sde_register(CHAR(STRING_ELT(id_, 0)), &id));
    M->model_id = id;
  } else if (TYPEOF(id_) == INTSXP || TYPEOF(id_) == REALSXP) {
    M->model_id = asInteger(id_);
  } else {
    error("bssm_b200: ssm_sde needs a registry model id or the model's CUDA source in place of the drift pointer");
  }
  M->n = (int)XLENGTH(y); M->y = REAL(y);
  M->theta = REAL(theta); M->n_theta = (int)XLENGTH(theta);
  M->x0 = asReal(x0); M->positive = asLogical(positive);
}

/* bsf_sde(y, x0, positive, drift, diffusion, ddiffusion, log_prior_pdf, log_obs_density, theta, nsim, L, seed):
 * src/R_sde.cpp:35-72.  att / Ptt have n + 1 columns for SDE models (the last one zero), as in the reference. */
SEXP _bssmb200_bsf_sde(SEXP y, SEXP x0, SEXP positive, SEXP drift, SEXP diffusion, SEXP ddiffusion, SEXP log_prior_pdf,
                       SEXP log_obs_density, SEXP theta, SEXP nsim_, SEXP L_, SEXP seed_) {
  bssm_sde_model M;
  fill_sde(y, x0, positive, drift, theta, &M);
  const int n = M.n;
  unsigned N = (unsigned)asInteger(nsim_);
  SEXP at = PROTECT(allocMatrix(REALSXP, 1, n + 1)), att = PROTECT(allocMatrix(REALSXP, 1, n + 1));
  SEXP Pt = PROTECT(alloc3DArray(REALSXP, 1, 1, n + 1)), Ptt = PROTECT(alloc3DArray(REALSXP, 1, 1, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, 1, n + 1, N));
  memset(REAL(att), 0, sizeof(double) * (n + 1)); memset(REAL(Ptt), 0, sizeof(double) * (n + 1));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.at = REAL(at); out.att = REAL(att); out.Pt = REAL(Pt); out.Ptt = REAL(Ptt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf_sde(&M, N, (unsigned)asInteger(L_), (unsigned)asInteger(seed_), &out));
  warn_stop(out.loglik);
  SEXP res = filter_result(at, att, Pt, Ptt, w, out.loglik, alpha);
  UNPROTECT(6);
  return res;
}

/* bsf_smoother_sde(same 12 arguments): src/R_sde.cpp:75-111 */
SEXP _bssmb200_bsf_smoother_sde(SEXP y, SEXP x0, SEXP positive, SEXP drift, SEXP diffusion, SEXP ddiffusion,
                                SEXP log_prior_pdf, SEXP log_obs_density, SEXP theta, SEXP nsim_, SEXP L_, SEXP seed_) {
  bssm_sde_model M;
  fill_sde(y, x0, positive, drift, theta, &M);
  const int n = M.n;
  unsigned N = (unsigned)asInteger(nsim_);
  SEXP ah = PROTECT(allocMatrix(REALSXP, 1, n + 1)), Vt = PROTECT(alloc3DArray(REALSXP, 1, 1, n + 1));
  SEXP w = PROTECT(allocMatrix(REALSXP, N, n + 1)), alpha = PROTECT(alloc3DArray(REALSXP, 1, n + 1, N));
  bssm_filter_out out;
  memset(&out, 0, sizeof(out));
  out.alphahat = REAL(ah); out.Vt = REAL(Vt); out.weights = REAL(w); out.alpha = REAL(alpha);
  check_rc(bssm_b200_bsf_smoother_sde(&M, N, (unsigned)asInteger(L_), (unsigned)asInteger(seed_), &out));
  warn_stop(out.loglik);
  SEXP res = smoother_result(ah, Vt, w, out.loglik, alpha);
  UNPROTECT(4);
  return res;
}

/* loglik_sde(same 12 arguments): src/R_sde.cpp:10-32 -- nothing but the scalar crosses PCIe */
SEXP _bssmb200_loglik_sde(SEXP y, SEXP x0, SEXP positive, SEXP drift, SEXP diffusion, SEXP ddiffusion, SEXP log_prior_pdf,
                          SEXP log_obs_density, SEXP theta, SEXP nsim_, SEXP L_, SEXP seed_) {
  bssm_sde_model M;
  fill_sde(y, x0, positive, drift, theta, &M);
  double loglik = 0.0;
  check_rc(bssm_b200_loglik_sde(&M, (unsigned)asInteger(nsim_), (unsigned)asInteger(L_), (unsigned)asInteger(seed_), &loglik));
  return ScalarReal(loglik);
}

/* stratified_sample(p, r, N): src/stratified_sample.cpp:9-28 (unit hook, R/RcppExports.R:228) */
SEXP _bssmb200_stratified_sample(SEXP p, SEXP r, SEXP N_) {
  unsigned N = (unsigned)asInteger(N_);
  SEXP xp = PROTECT(allocVector(INTSXP, N));
  check_rc(bssm_b200_stratified_sample(REAL(p), (unsigned)XLENGTH(p), REAL(r), N, (uint32_t*)INTEGER(xp)));
  UNPROTECT(1);
  return xp;
}

static const R_CallMethodDef CallEntries[] = {
  {"_bssmb200_bsf", (DL_FUNC)&_bssmb200_bsf, 5},
  {"_bssmb200_psi_smoother", (DL_FUNC)&_bssmb200_psi_smoother, 4},
  {"_bssmb200_bsf_smoother", (DL_FUNC)&_bssmb200_bsf_smoother, 5},
  {"_bssmb200_nongaussian_loglik", (DL_FUNC)&_bssmb200_nongaussian_loglik, 5},
  {"_bssmb200_importance_sample_ng", (DL_FUNC)&_bssmb200_importance_sample_ng, 5},
  {"_bssmb200_suggest_n_nongaussian", (DL_FUNC)&_bssmb200_suggest_n_nongaussian, 6},
  {"_bssmb200_postcorrection_nongaussian", (DL_FUNC)&_bssmb200_postcorrection_nongaussian, 10},
  {"_bssmb200_bsf_nlg", (DL_FUNC)&_bssmb200_bsf_nlg, 18},
  {"_bssmb200_bsf_smoother_nlg", (DL_FUNC)&_bssmb200_bsf_smoother_nlg, 18},
  {"_bssmb200_psi_smoother_nlg", (DL_FUNC)&_bssmb200_psi_smoother_nlg, 21},
  {"_bssmb200_postcorrection_nonlinear", (DL_FUNC)&_bssmb200_postcorrection_nonlinear, 24},
  {"_bssmb200_bsf_sde", (DL_FUNC)&_bssmb200_bsf_sde, 12},
  {"_bssmb200_bsf_smoother_sde", (DL_FUNC)&_bssmb200_bsf_smoother_sde, 12},
  {"_bssmb200_loglik_sde", (DL_FUNC)&_bssmb200_loglik_sde, 12},
  {"_bssmb200_stratified_sample", (DL_FUNC)&_bssmb200_stratified_sample, 3},
  {NULL, NULL, 0}
};

void R_init_bssmb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

void R_unload_bssmb200(DllInfo* dll) { bssm_b200_shutdown(); } /* reference: R/zzz.R:1-3 */
