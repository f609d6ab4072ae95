/* bssm_b200_shim.c -- the `.Call` side of the drop-in boundary.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: it needs R's headers (Rinternals.h),
 * which the build image does not have.  It is the ~150-line glue a bssm
 * maintainer adds to reach libbssm_b200.so from R; everything it calls is the
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

/* bsf(model_, nsim, seed, gaussian, model_type): src/R_bootstrap_filter.cpp:15-241 */
SEXP _bssmb200_bsf(SEXP model_, SEXP nsim_, SEXP seed_, SEXP gaussian_, SEXP model_type_) {
  if (asLogical(gaussian_)) error("bssm_b200: Gaussian models use bssm's own Kalman/BSF code path");
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

/* postcorrection_nongaussian(model_, model_type, output_type, nsim, seed, n_threads, is_type,
 *                            counts, theta, modes): src/R_postcorrection.cpp:150-248 (output_type 3 shown) */
SEXP _bssmb200_postcorrection_nongaussian(SEXP model_, SEXP model_type_, SEXP output_type_, SEXP nsim_,
                                          SEXP seed_, SEXP n_threads_, SEXP is_type_, SEXP counts_,
                                          SEXP theta_, SEXP modes_) {
  bssm_ung_model M;
  int n = (int)XLENGTH(list_get(model_, "y"));
  double* xb = (double*)R_alloc(n, sizeof(double));
  fill_ung(model_, &M, xb);
  bssm_theta_map map;
  memset(&map, 0, sizeof(map));
  map.model_type = asInteger(model_type_);
  if (map.model_type == 3) map.svm_type = asInteger(list_get(model_, "svm_type"));
  if (map.model_type == 4) { map.mu_est = asLogical(list_get(model_, "mu_est")); map.phi_est = asLogical(list_get(model_, "phi_est")); }
  int n_stored = (int)XLENGTH(counts_), n_par = INTEGER(getAttrib(theta_, R_DimSymbol))[0];
  int is_type = asInteger(is_type_);
  R_xlen_t nd = 0;
  for (int i = 0; i < n_stored; i++) nd += is_type == 3 ? INTEGER(counts_)[i] : 1;
  SEXP w = PROTECT(allocVector(REALSXP, nd)), post = PROTECT(allocVector(REALSXP, nd));
  bssm_postcorr_out out;
  memset(&out, 0, sizeof(out));
  out.weights = REAL(w); out.posterior = REAL(post);
  check_rc(bssm_b200_postcorrection_nongaussian(&M, &map, 3, (unsigned)asInteger(nsim_), (unsigned)asInteger(seed_),
                                                (unsigned)asInteger(n_threads_), is_type,
                                                (const uint32_t*)INTEGER(counts_), REAL(theta_), (unsigned)n_par,
                                                REAL(modes_), (unsigned)n_stored, 0, 0, &out));
  const char* nm[] = {"weights", "posterior", ""};
  SEXP res = PROTECT(mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(res, 0, w); SET_VECTOR_ELT(res, 1, post);
  UNPROTECT(3);
  return res;
}

static const R_CallMethodDef CallEntries[] = {
  {"_bssmb200_bsf", (DL_FUNC)&_bssmb200_bsf, 5},
  {"_bssmb200_psi_smoother", (DL_FUNC)&_bssmb200_psi_smoother, 4},
  {"_bssmb200_postcorrection_nongaussian", (DL_FUNC)&_bssmb200_postcorrection_nongaussian, 10},
  {NULL, NULL, 0}
};

void R_init_bssmb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

void R_unload_bssmb200(DllInfo* dll) { bssm_b200_shutdown(); } /* reference: R/zzz.R:1-3 */
