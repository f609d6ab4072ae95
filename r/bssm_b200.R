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
  if (sampling_method > 2 || nsim == 0)   # SPDK / pure approximation stay on the CPU path
    return(bssm:::nongaussian_loglik(model_, nsim, sampling_method, seed, model_type))
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_nongaussian_loglik", model_, nsim, sampling_method, seed, model_type)
}

suggest_n_nongaussian <- function(model_, theta, candidates, replications, seed, model_type) {
  model_$u <- rep_len(model_$u, length(model_$y))
  .Call("_bssmb200_suggest_n_nongaussian", model_, theta, as.numeric(candidates), replications, seed, model_type)
}

# check_intmax caps particles at 1e5 (R/check_arguments.R:407-418); the GPU path lifts it
for (f in c("bsf", "bsf_smoother", "psi_smoother", "nongaussian_loglik", "suggest_n_nongaussian",
            "postcorrection_nongaussian")) {
  fn <- get(f)
  environment(fn) <- asNamespace("bssm")
  utils::assignInNamespace(f, fn, "bssm")
}
