# bssm_b200.R -- R side of the drop-in: masks the reference's `.Call` stubs
# (R/RcppExports.R:12-229) for the particle-filter path, so bootstrap_filter(),
# particle_smoother() and post_correct() of an installed bssm run on the GPU.
# Needs r/bssm_b200_shim.c built with R CMD SHLIB (not possible in this repo's
# image: no R).  Everything else in bssm stays untouched.
dyn.load("bssmb200.so")

bsf <- function(model_, nsim, seed, gaussian, model_type) {
  if (gaussian) return(bssm:::bsf(model_, nsim, seed, gaussian, model_type))
  model_$u <- rep_len(model_$u, length(model_$y))          # svm stores u = 1
  out <- .Call("_bssmb200_bsf", model_, nsim, seed, gaussian, model_type)
  out$at <- t(out$at); out$att <- t(out$att)               # arma::inplace_trans
  out
}

psi_smoother <- function(model_, nsim, seed, model_type) {
  model_$u <- rep_len(model_$u, length(model_$y))
  out <- .Call("_bssmb200_psi_smoother", model_, nsim, seed, model_type)
  out$alphahat <- t(out$alphahat)
  out
}

postcorrection_nongaussian <- function(model_, model_type, output_type, nsim, seed, n_threads,
                                       is_type, counts, theta, modes) {
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_postcorrection_nongaussian", model_, model_type, output_type, nsim, seed,
        n_threads, is_type, as.integer(counts), theta, modes)
}

bsf_smoother <- function(model_, nsim, seed, gaussian, model_type) {
  if (gaussian) return(bssm:::bsf_smoother(model_, nsim, seed, gaussian, model_type))
  model_$u <- rep_len(model_$u, length(model_$y))
  out <- .Call("_bssmb200_bsf_smoother", model_, nsim, seed, gaussian, model_type)
  out$alphahat <- t(out$alphahat)
  out
}

nongaussian_loglik <- function(model_, nsim, sampling_method, seed, model_type) {
  # the pure approximation (nsim = 0), SPDK with a single draw and SPDK of ssm_mng stay on the CPU path
  if (nsim == 0 || (sampling_method == 3 && (nsim < 2 || model_type == 0)))
    return(bssm:::nongaussian_loglik(model_, nsim, sampling_method, seed, model_type))
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_nongaussian_loglik", model_, nsim, sampling_method, seed, model_type)
}

importance_sample_ng <- function(model_, nsim, use_antithetic, seed, model_type) {
  if (model_type == 0 || nsim < 2)
    return(bssm:::importance_sample_ng(model_, nsim, use_antithetic, seed, model_type))
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_importance_sample_ng", model_, nsim, use_antithetic, seed, model_type)
}

suggest_n_nongaussian <- function(model_, theta, candidates, replications, seed, model_type) {
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_suggest_n_nongaussian", model_, theta, as.numeric(candidates), replications, seed, model_type)
}


# ---- ssm_nlg / ssm_sde: the shipped example models are device policies selected by id
# (include/bssm_b200.h BSSM_NLG_* / BSSM_SDE_*).  A model built with bssm's cpp_example_model("nlg_growth")
# etc. carries compiled function pointers that cannot run in a kernel; tag it once with
#   model$b200_id <- b200_model_id("nlg_growth")
# and the masked stubs below put the id in place of the first pointer argument.
b200_model_id <- function(name) {
  switch(name, nlg_growth = 0L, nlg_ar_exp = 1L, nlg_sin_exp = 2L, nlg_linear_gaussian = 3L,
         sde_gbm = 0L, sde_poisson_OU = 1L, stop("no device policy for this example model"))
}
# An ssm_sde model of the user's own: the drift pointer carries the CUDA source of the four functions
#   attr(model$drift, "b200_source") <- "__device__ double drift(double x, const double* theta) {...} ..."
# (same bodies as the Rcpp::sourceCpp file of vignettes/sde_model.Rmd, scalar signatures of
# include/bssm_b200.h bssm_b200_sde_register); the shim compiles it on first use.
.b200_id <- function(ptr) {
  src <- attr(ptr, "b200_source")
  if (!is.null(src)) return(as.character(src))
  id <- attr(ptr, "b200_id")
  if (is.null(id)) stop("bssm_b200: tag the model's Z_fn / drift pointer with attr(., 'b200_id') <- b200_model_id(name)")
  as.integer(id)
}

bsf_nlg <- function(y, Z, H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf, known_params, known_tv_params,
                    n_states, n_etas, time_varying, nsim, seed) {
  out <- .Call("_bssmb200_bsf_nlg", y, .b200_id(Z), H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf, known_params,
               known_tv_params, n_states, n_etas, time_varying, nsim, seed)
  out$at <- t(out$at); out$att <- t(out$att)
  out
}
bsf_smoother_nlg <- function(y, Z, H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf, known_params, known_tv_params,
                             n_states, n_etas, time_varying, nsim, seed) {
  out <- .Call("_bssmb200_bsf_smoother_nlg", y, .b200_id(Z), H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf,
               known_params, known_tv_params, n_states, n_etas, time_varying, nsim, seed)
  out$alphahat <- t(out$alphahat)
  out
}
psi_smoother_nlg <- function(y, Z, H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf, known_params, known_tv_params,
                             n_states, n_etas, time_varying, nsim, seed, max_iter, conv_tol, iekf_iter) {
  out <- .Call("_bssmb200_psi_smoother_nlg", y, .b200_id(Z), H, T, R, Zg, Tg, a1, P1, theta, log_prior_pdf,
               known_params, known_tv_params, n_states, n_etas, time_varying, nsim, seed, max_iter, conv_tol, iekf_iter)
  out$alphahat <- t(out$alphahat)
  out
}
postcorrection_nonlinear <- function(y, Z, H, T, R, Zg, Tg, a1, P1, theta_init, log_prior_pdf, known_params,
                                     known_tv_params, n_states, n_etas, time_varying, output_type, nsim, seed,
                                     n_threads, is_type, counts, theta, modes) {
  .Call("_bssmb200_postcorrection_nonlinear", y, .b200_id(Z), H, T, R, Zg, Tg, a1, P1, theta_init, log_prior_pdf,
        known_params, known_tv_params, n_states, n_etas, time_varying, output_type, nsim, seed, n_threads, is_type,
        as.integer(counts), theta, modes)
}
bsf_sde <- function(y, x0, positive, drift, diffusion, ddiffusion, log_prior_pdf, log_obs_density, theta, nsim, L, seed) {
  out <- .Call("_bssmb200_bsf_sde", y, x0, positive, .b200_id(drift), diffusion, ddiffusion, log_prior_pdf,
               log_obs_density, theta, nsim, L, seed)
  out$at <- t(out$at); out$att <- t(out$att)
  out
}
bsf_smoother_sde <- function(y, x0, positive, drift, diffusion, ddiffusion, log_prior_pdf, log_obs_density, theta,
                             nsim, L, seed) {
  out <- .Call("_bssmb200_bsf_smoother_sde", y, x0, positive, .b200_id(drift), diffusion, ddiffusion, log_prior_pdf,
               log_obs_density, theta, nsim, L, seed)
  out$alphahat <- t(out$alphahat)
  out
}
loglik_sde <- function(y, x0, positive, drift, diffusion, ddiffusion, log_prior_pdf, log_obs_density, theta, nsim, L, seed) {
  .Call("_bssmb200_loglik_sde", y, x0, positive, .b200_id(drift), diffusion, ddiffusion, log_prior_pdf,
        log_obs_density, theta, nsim, L, seed)
}
stratified_sample <- function(p, r, N) .Call("_bssmb200_stratified_sample", as.numeric(p), as.numeric(r), as.integer(N))

# check_intmax caps particles at 1e5 (R/check_arguments.R:407-418); the GPU path lifts it
for (f in c("bsf", "bsf_smoother", "psi_smoother", "nongaussian_loglik", "importance_sample_ng", "suggest_n_nongaussian",
            "postcorrection_nongaussian", "bsf_nlg", "bsf_smoother_nlg", "psi_smoother_nlg",
            "postcorrection_nonlinear", "bsf_sde", "bsf_smoother_sde", "loglik_sde", "stratified_sample")) {
  fn <- get(f)
  environment(fn) <- asNamespace("bssm")
  utils::assignInNamespace(f, fn, "bssm")
}
