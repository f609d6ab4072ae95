"""Host-side mirror of the reference's R surface for the particle-filter path.

  bootstrap_filter   R/bootstrap_filter.R:85-120   -> bssm_b200_bsf
  particle_smoother  R/particle_smoother.R:140-186 -> bssm_b200_psi_smoother /
                                                      bssm_b200_bsf_smoother
  logLik             R/loglik.R (particle part)    -> filter-only jobs

Same argument names and checks as the R functions (particles, seed, method);
results come back as dicts with the R list's names and array shapes
(alpha already permuted to (n+1) x m x N as R/bootstrap_filter.R:118 does).
Everything is computed by the CUDA library; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FilterOut, check, dptr, lib, uptr
from .models import Model

METHOD_PSI, METHOD_BSF = 1, 2
POST_NONE, POST_FILTER_SUMMARY, POST_SMOOTHER = 0, 1, 2
OUT_ALPHA, OUT_WEIGHTS, OUT_INDICES = 1, 2, 4


def _check_particles(model, particles, seed):
    # R/bootstrap_filter.R:88-105 (check_intmax caps particles at 1e5 in the
    # reference, R/check_arguments.R:407-418; lifted here: N up to 2^20 is the point)
    particles = int(particles)
    if particles < 2:
        raise ValueError("Argument 'particles' should be an integer >= 2.")
    if particles > 2 ** 20:  # exactness of the fixed-point weight sums (pf_math.h) and one resident group of CTAs
        raise ValueError("Argument 'particles' is too large for one filter (at most 2^20).")
    seed = int(seed)
    if not 0 <= seed <= 2 ** 31 - 1:
        raise ValueError("Argument 'seed' should be an integer in [0, 2^31-1].")
    return particles, seed


class _Buffers:
    def __init__(self, model, N, alpha=True, weights=True, indices=False, filt=False, smooth=False):
        n, m = model.n, model.m
        self.out = FilterOut()
        self.alpha = np.zeros((m, n + 1, N), order="F") if alpha else None
        self.weights = np.zeros((N, n + 1), order="F") if weights else None
        self.indices = np.zeros((N, n), dtype=np.uint32, order="F") if indices else None
        self.at = np.zeros((m, n + 1), order="F") if filt else None
        self.att = np.zeros((m, n), order="F") if filt else None
        self.Pt = np.zeros((m, m, n + 1), order="F") if filt else None
        self.Ptt = np.zeros((m, m, n), order="F") if filt else None
        self.alphahat = np.zeros((m, n + 1), order="F") if smooth else None
        self.Vt = np.zeros((m, m, n + 1), order="F") if smooth else None
        o = self.out
        o.alpha, o.weights, o.indices = dptr(self.alpha), dptr(self.weights), uptr(self.indices)
        o.at, o.att, o.Pt, o.Ptt = dptr(self.at), dptr(self.att), dptr(self.Pt), dptr(self.Ptt)
        o.alphahat, o.Vt = dptr(self.alphahat), dptr(self.Vt)


def _warn_stopped(out):
    if not np.isfinite(out.loglik):
        import warnings
        # src/R_bootstrap_filter.cpp:166-167
        warnings.warn("Particle filtering stopped prematurely due to nonfinite log-likelihood.")


def bootstrap_filter(model, particles, seed=1, return_indices=False, L=None):
    """bootstrap_filter: .nongaussian R/bootstrap_filter.R:85-120, .ssm_nlg :127-159,
    .ssm_sde :165-202 (L = discretisation level, 2^L Milstein sub-steps)."""
    N, seed = _check_particles(model, particles, seed)
    b = _Buffers(model, N, indices=return_indices, filt=True)
    ms = model.c_struct()
    if model.cls == "ssm_nlg":
        check(lib().bssm_b200_bsf_nlg(C.byref(ms), N, seed, C.byref(b.out)))
    elif model.cls == "ssm_mng":
        check(lib().bssm_b200_bsf_mng(C.byref(ms), N, seed, C.byref(b.out)))
    elif model.cls == "ssm_sde":
        if L is None or int(L) < 1:
            raise ValueError("Discretization level L must be larger than 0.")  # R/bootstrap_filter.R:169
        # the reference returns att / Ptt with n+1 columns for SDE models (last one zero)
        n, m = model.n, model.m
        b.att = np.zeros((m, n + 1), order="F")
        b.Ptt = np.zeros((m, m, n + 1), order="F")
        b.out.att, b.out.Ptt = dptr(b.att), dptr(b.Ptt)
        check(lib().bssm_b200_bsf_sde(C.byref(ms), N, int(L), seed, C.byref(b.out)))
    else:
        check(lib().bssm_b200_bsf(C.byref(ms), N, seed, C.byref(b.out)))
    _warn_stopped(b.out)
    res = {"at": b.at.T.copy(), "att": b.att.T.copy(), "Pt": b.Pt, "Ptt": b.Ptt,
           "weights": b.weights, "logLik": b.out.loglik,
           "alpha": np.transpose(b.alpha, (1, 0, 2)),  # aperm(alpha, c(2, 1, 3))
           "device_ms": b.out.device_ms, "stopped_at": b.out.stopped_at}
    if return_indices:
        res["indices"] = b.indices
    return res


def ekpf_filter(model, particles, seed=1, return_indices=False):
    """ekpf_filter, R/ekpf_filter.R (-> ekpf, src/R_ekpf.cpp:7-53): EKF-based particle filter of an
    ssm_nlg model; same outputs as bootstrap_filter."""
    if model.cls != "ssm_nlg":
        raise ValueError("ekpf_filter is defined for ssm_nlg models")
    N, seed = _check_particles(model, particles, seed)
    b = _Buffers(model, N, indices=return_indices, filt=True)
    ms = model.c_struct()
    check(lib().bssm_b200_ekpf_nlg(C.byref(ms), N, seed, C.byref(b.out)))
    _warn_stopped(b.out)
    res = {"at": b.at.T.copy(), "att": b.att.T.copy(), "Pt": b.Pt, "Ptt": b.Ptt,
           "weights": b.weights, "logLik": b.out.loglik, "alpha": np.transpose(b.alpha, (1, 0, 2)),
           "device_ms": b.out.device_ms, "stopped_at": b.out.stopped_at}
    if return_indices:
        res["indices"] = b.indices
    return res


def particle_smoother(model, particles, method="psi", seed=1, return_indices=False, L=None):
    """particle_smoother: .lineargaussian R/particle_smoother.R:85-135, .nongaussian :140-186,
    .ssm_nlg :192-245, .ssm_sde :251-291.

    Returns the `.Call`-level result: alphahat and Vt have n + 1 time points (the extra one is the
    prediction for time n + 1); the R wrappers drop that last row (e.g. :179-181)."""
    if method not in ("psi", "bsf", "ekf"):
        raise ValueError("method must be 'psi', 'bsf' or 'ekf'")
    if method == "ekf" and model.cls != "ssm_nlg":
        raise ValueError("method = 'ekf' is defined for ssm_nlg models")  # R/particle_smoother.R:192-245
    dist = getattr(model, "distribution", None)
    if isinstance(dist, str) and dist == "gaussian" and method == "psi":
        # particle_smoother.lineargaussian(method = "psi") -> gaussian_psi_smoother, R/particle_smoother.R:111-128
        return gaussian_psi_smoother(model, particles, seed)
    N, seed = _check_particles(model, particles, seed)
    b = _Buffers(model, N, indices=return_indices, smooth=True)
    ms = model.c_struct()
    if model.cls == "ssm_nlg" and method == "psi":
        # particle_smoother.ssm_nlg -> psi_smoother_nlg(..., max_iter, conv_tol, iekf_iter), R/particle_smoother.R:221-228
        check(lib().bssm_b200_psi_smoother_nlg(C.byref(ms), N, seed, int(getattr(model, "max_iter", 100)),
                                               float(getattr(model, "conv_tol", 1e-8)),
                                               int(getattr(model, "iekf_iter", 0)), C.byref(b.out)))
    elif model.cls == "ssm_mng":
        # psi_smoother(model_type = 0) src/R_psi.cpp:57-78 / bsf_smoother src/R_bootstrap_filter.cpp:360-385
        fn = lib().bssm_b200_psi_smoother_mng if method == "psi" else lib().bssm_b200_bsf_smoother_mng
        check(fn(C.byref(ms), N, seed, C.byref(b.out)))
    elif model.cls == "ssm_nlg" and method == "ekf":
        check(lib().bssm_b200_ekpf_smoother_nlg(C.byref(ms), N, seed, C.byref(b.out)))
    elif model.cls == "ssm_nlg":
        check(lib().bssm_b200_bsf_smoother_nlg(C.byref(ms), N, seed, C.byref(b.out)))
    elif model.cls == "ssm_sde":
        if L is None or int(L) < 1:
            raise ValueError("Discretization level L must be larger than 0.")
        check(lib().bssm_b200_bsf_smoother_sde(C.byref(ms), N, int(L), seed, C.byref(b.out)))
    else:
        fn = lib().bssm_b200_psi_smoother if method == "psi" else lib().bssm_b200_bsf_smoother
        check(fn(C.byref(ms), N, seed, C.byref(b.out)))
    _warn_stopped(b.out)
    res = {"alphahat": b.alphahat.T.copy(), "Vt": b.Vt, "weights": b.weights,
           "logLik": b.out.loglik, "alpha": np.transpose(b.alpha, (1, 0, 2)),
           "device_ms": b.out.device_ms, "stopped_at": b.out.stopped_at,
           "approx_logLik": b.out.approx_loglik}
    if return_indices:
        res["indices"] = b.indices
    return res


class FilterJob:
    """Resident batch of replicate filters (bssm_b200_job_*): device-side
    inputs and outputs, run() launches kernels only."""

    def __init__(self, model, particles, method="psi", seed=1, n_filters=1, first_filter=0,
                 out_flags=0, post=POST_NONE, L=None):
        N, seed = _check_particles(model, particles, seed)
        self.model, self.N, self.n_filters = model, N, int(n_filters)
        self._ms = model.c_struct()
        self._h = C.c_void_p()
        meth = METHOD_PSI if method == "psi" else METHOD_BSF
        if model.cls == "ssm_mng":
            create = lib().bssm_b200_job_create_mng_psi if method == "psi" else lib().bssm_b200_job_create_mng
            check(create(C.byref(self._ms), N, seed, self.n_filters, int(first_filter),
                                                 int(out_flags), int(post), C.byref(self._h)))
        elif model.cls == "ssm_nlg" and method == "psi":
            check(lib().bssm_b200_job_create_nlg_psi(C.byref(self._ms), N, seed, int(getattr(model, "max_iter", 100)),
                                                     float(getattr(model, "conv_tol", 1e-8)), self.n_filters,
                                                     int(first_filter), int(out_flags), int(post), C.byref(self._h)))
        elif model.cls == "ssm_nlg":
            create = lib().bssm_b200_job_create_nlg_ekpf if method == "ekf" else lib().bssm_b200_job_create_nlg
            check(create(C.byref(self._ms), N, seed, self.n_filters, int(first_filter),
                                                 int(out_flags), int(post), C.byref(self._h)))
        elif model.cls == "ssm_sde":
            check(lib().bssm_b200_job_create_sde(C.byref(self._ms), N, int(L), seed, self.n_filters,
                                                 int(first_filter), int(out_flags), int(post), C.byref(self._h)))
        else:
            check(lib().bssm_b200_job_create_ung(C.byref(self._ms), meth, N, seed, self.n_filters,
                                                 int(first_filter), int(out_flags), int(post),
                                                 C.byref(self._h)))

    def run(self):
        ms = C.c_float()
        check(lib().bssm_b200_job_run(self._h, C.byref(ms)))
        return ms.value

    def set_theta(self, theta):
        """Re-parameterise the resident job (the filter inside an MCMC loop, src/mcmc.cpp:338-491):
        update_model(theta) + approximate() on the host, then run() again."""
        theta = _lib.f64(np.ravel(theta))
        if self.model.cls == "ssm_nlg":    # bootstrap / EKF particle filter jobs: theta travels by value
            check(lib().bssm_b200_job_set_theta_nlg(self._h, C.byref(self._ms), dptr(theta), theta.size))
            return
        if self.model.cls == "ssm_sde":
            check(lib().bssm_b200_job_set_theta_sde(self._h, dptr(theta), theta.size))
            return
        tm = theta_map(self.model)
        check(lib().bssm_b200_job_set_theta(self._h, C.byref(self._ms), C.byref(tm), dptr(theta), theta.size))

    def kernel_ms(self):
        a = (C.c_float * 3)()
        check(lib().bssm_b200_job_kernel_ms(self._h, C.byref(a)))
        return list(a)

    def launches(self):
        return lib().bssm_b200_job_launches(self._h)

    def phase_cycles(self, cta=0):
        a = (C.c_uint64 * 6)()
        check(lib().bssm_b200_job_phase_cycles(self._h, int(cta), C.byref(a)))
        return list(a)

    def max_concurrent_filters(self):
        return lib().bssm_b200_job_max_concurrent_filters(self._h)

    def logliks(self):
        out = np.zeros(self.n_filters)
        check(lib().bssm_b200_job_logliks(self._h, dptr(out)))
        return out

    def logliks_to_device(self, dev_ptr):
        check(lib().bssm_b200_job_logliks_dev(self._h, C.c_void_p(dev_ptr)))

    def fetch(self, f=0, alpha=False, weights=False, indices=False, filt=False, smooth=False):
        b = _Buffers(self.model, self.N, alpha=alpha, weights=weights, indices=indices, filt=filt, smooth=smooth)
        check(lib().bssm_b200_job_fetch(self._h, int(f), C.byref(b.out)))
        return b

    def close(self):
        if self._h:
            lib().bssm_b200_job_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def importance_sample(model, nsim, use_antithetic=True, max_iter=None, conv_tol=None, seed=1, return_loglik=False):
    """importance_sample.nongaussian, R/importance_sample.R:41-70 -> importance_sample_ng
    (src/R_importance_sample.cpp:9-75): `nsim` draws from the smoothing distribution of the approximating
    Gaussian model (simulation smoother, one antithetic variable) and their importance weights.  Returns
    alpha (n+1) x m x nsim and weights (nsim); ssm_ung family, nsim >= 2.  The reference never forwards
    `use_antithetic` to simulate_states (its draws are always paired, :29); here FALSE draws nsim independent
    trajectories."""
    if not isinstance(model, Model) or model.distribution == "gaussian":
        raise ValueError("importance_sample is implemented for non-Gaussian ssm_ung models (svm, bsm_ng, ar1_ng, ssm_ung)")
    if not isinstance(use_antithetic, (bool, np.bool_)):
        raise ValueError("Argument 'use_antithetic' should be TRUE or FALSE. ")
    nsim, seed = _check_particles(model, nsim, seed)
    keep = (model.max_iter, model.conv_tol)
    try:
        if max_iter is not None:
            model.max_iter = int(max_iter)
        if conv_tol is not None:
            model.conv_tol = float(conv_tol)
        ms = model.c_struct()
        alpha = np.zeros((model.m, model.n + 1, nsim), order="F")
        w = np.zeros(nsim)
        ll = np.zeros(2)
        check(lib().bssm_b200_importance_sample_ung(C.byref(ms), nsim, int(bool(use_antithetic)), seed, dptr(alpha),
                                                    dptr(w), dptr(ll)))
    finally:
        model.max_iter, model.conv_tol = keep
    out = {"alpha": np.transpose(alpha, (1, 0, 2)), "weights": w}
    if return_loglik:
        out["logLik"], out["approx_logLik"] = float(ll[0]), float(ll[1])
    return out


def logLik(model, particles, method="psi", seed=1, L=None):
    """logLik (particle part), R/loglik.R:60-161: log-likelihood only.  method "spdk" = non-sequential
    importance sampling from the Gaussian approximation (src/model_ssm_ung.cpp:199-211)."""
    if method == "spdk":
        if not isinstance(model, Model) or model.distribution == "gaussian":
            raise ValueError("method 'spdk': non-Gaussian ssm_ung models")
        N, seed = _check_particles(model, particles, seed)
        ms = model.c_struct()
        ll = np.zeros(2)
        check(lib().bssm_b200_importance_sample_ung(C.byref(ms), N, 1, seed, None, None, dptr(ll)))
        return float(ll[0])
    job = FilterJob(model, particles, method=method, seed=seed, L=L)
    try:
        job.run()
        return float(job.logliks()[0])
    finally:
        job.close()


def stratified_sample(p, r):
    """Unit hook of src/stratified_sample.cpp:9-28 (fixed-point cumsum)."""
    p = _lib.f64(np.ravel(p))
    r = _lib.f64(np.ravel(r))
    xp = np.zeros(r.size, dtype=np.uint32)
    check(lib().bssm_b200_stratified_sample(dptr(p), p.size, dptr(r), r.size, uptr(xp)))
    return xp


# ----------------------------------------------------------- post-correction
def theta_map(model):
    """which C++ update_model applies (R/model_type.R:1-16)"""
    from ._lib import ThetaMap
    tm = ThetaMap()
    if model.cls == "svm":
        tm.model_type, tm.svm_type = 3, int(model.extra.get("svm_type", 0))
    elif model.cls == "ar1_ng":
        tm.model_type = 4
        tm.mu_est = int(model.extra.get("mu_est", 1))
        tm.phi_est = int(model.extra.get("phi_est", 0))
    elif model.cls == "bsm_ng":
        tm.model_type = 2
        for key in ("slope", "seasonal", "noise", "level_est", "slope_est", "seasonal_est", "phi_est"):
            setattr(tm, key, int(bool(model.extra.get(key, False))))
    else:
        raise ValueError("post_correct is implemented for svm, ar1_ng, bsm_ng and ssm_nlg models")
    if model.cls in ("ar1_ng", "bsm_ng") and model.xreg is not None and model.xreg.size:
        # beta is the tail of theta: update_model recomputes xbeta (src/model_ar1_ng.cpp:28-31, model_bsm_ng.cpp:46-49)
        xr = _lib.f64(np.asfortranarray(model.xreg.reshape(model.n, -1)))
        tm.n_xreg, tm.xreg = xr.shape[1], dptr(xr)
        tm._keep = xr
    return tm


def suggest_N(model, theta, candidates=tuple(range(10, 101, 10)), replications=100, seed=1):
    """suggest_N, R/post_correction.R:85-134: sd of the psi-APF log-likelihood estimate for each
    candidate particle count; N = the first candidate with sd < 1."""
    replications = int(replications)
    if not 2 <= replications <= 1000:
        raise ValueError("replications must be in 2..1000")  # check_intmax(..., max = 1000)
    cand = np.ascontiguousarray(candidates, dtype=np.int64)
    if cand.size < 1 or np.any(cand < 1):
        raise ValueError("Argument 'candidates' should be vector of positive integers. ")
    theta = _lib.f64(np.ravel(theta))
    cand32 = np.ascontiguousarray(cand, dtype=np.uint32)
    sds = np.zeros(cand.size)
    if getattr(model, "update_fn", None) is not None and model.cls in ("ssm_ung", "ssm_mng"):
        # model.update_model(theta, update_fn), src/R_postcorrection.cpp:21-52: apply the closure here
        import copy
        mi = copy.copy(model)
        for name, val in model.update_fn(theta).items():
            if name == "phi" and model.cls == "ssm_ung":
                mi.phi = float(val)
            else:
                own = getattr(model, name)
                setattr(mi, name, _lib.f64(np.reshape(np.asarray(val, dtype=np.float64), np.shape(own), order="F")))
        ms = mi.c_struct()
        if model.cls == "ssm_mng":
            check(lib().bssm_b200_suggest_n_mng(C.byref(ms), uptr(cand32), cand32.size, replications, int(seed), dptr(sds)))
        else:
            from ._lib import ThetaMap
            tm = ThetaMap()
            tm.model_type = 1
            check(lib().bssm_b200_suggest_n_nongaussian(C.byref(ms), C.byref(tm), dptr(theta), theta.size, uptr(cand32),
                                                        cand32.size, replications, int(seed), dptr(sds)))
        below = np.nonzero(sds < 1)[0]
        return {"N": int(cand[below[0]]) if below.size else None, "results": {"N": cand.copy(), "sd": sds}}
    if theta.size != np.size(model.theta):
        raise ValueError("Length of 'theta' does not match length of 'model$theta'.")
    ms = model.c_struct()
    if model.cls == "ssm_nlg":
        check(lib().bssm_b200_suggest_n_nonlinear(C.byref(ms), dptr(theta), theta.size, uptr(cand32), cand32.size,
                                                  replications, int(seed), int(getattr(model, "max_iter", 100)),
                                                  float(getattr(model, "conv_tol", 1e-8)), dptr(sds)))
    else:
        tm = theta_map(model)
        check(lib().bssm_b200_suggest_n_nongaussian(C.byref(ms), C.byref(tm), dptr(theta), theta.size, uptr(cand32),
                                                    cand32.size, replications, int(seed), dptr(sds)))
    below = np.nonzero(sds < 1)[0]
    return {"N": int(cand[below[0]]) if below.size else None,
            "results": {"N": cand.copy(), "sd": sds}}


def _postcorr_inputs(model, mcmc_output):
    theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)  # t(theta): n_par x n_stored
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    rows = model.n * (model.m if model.cls == "ssm_nlg" else 1)  # nlg stores m x n modes per draw
    modes = np.asfortranarray(np.asarray(mcmc_output["modes"], dtype=np.float64).reshape(rows, -1, order="F"))
    if theta.shape[1] != counts.size or modes.shape[1] != counts.size:
        raise ValueError("theta, counts and modes disagree on the number of stored draws")
    return theta, counts, modes


PARSET_FIELDS = ("phi", "Z", "T", "R", "a1", "P1", "D", "C", "beta")


def build_parset(model, theta):
    """parset_ung / parset_mng constructors (src/parset_ng.cpp:5-82, :127-183): evaluates
    model.update_fn(theta_i) for every stored draw and stacks the components it returns.

    theta: n_stored x n_par.  update_fn returns a dict with any of phi, Z, T, R, a1, P1, D, C (and beta
    for ssm_ung), each shaped like the model's own array.  Returns (Parset struct, arrays kept alive)."""
    from ._lib import Parset
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    first = model.update_fn(theta[0])
    unknown = set(first) - set(PARSET_FIELDS)
    if unknown:
        raise ValueError("update_fn returned unknown components: %s" % sorted(unknown))
    if "beta" in first and model.cls != "ssm_ung":
        raise ValueError("only ssm_ung models have a regression part")
    ps, keep = Parset(), {}
    for name in first:
        own = {"a1": model.a1, "P1": model.P1, "phi": np.atleast_1d(model.phi), "Z": model.Z, "T": model.T,
               "R": model.R, "D": model.D, "C": model.C, "beta": getattr(model, "beta", None)}[name]
        size = int(np.size(own))
        stack = np.empty((theta.shape[0], size))
        for i in range(theta.shape[0]):
            comp = first if i == 0 else model.update_fn(theta[i])
            v = np.ravel(np.asarray(comp[name], dtype=np.float64), order="F")
            if v.size != size:
                raise ValueError("update_fn: %s has %d elements, the model's has %d" % (name, v.size, size))
            stack[i] = v
        keep[name] = _lib.f64(np.ravel(stack))
        setattr(ps, name, dptr(keep[name]))
    if "beta" in first:
        keep["xreg"] = _lib.f64(np.ravel(np.asarray(model.xreg, dtype=np.float64).reshape(model.n, -1), order="F"))
        ps.xreg, ps.n_xreg = dptr(keep["xreg"]), keep["xreg"].size // model.n
    ps._keep = keep
    return ps


def n_draws(mcmc_output, is_type):
    counts = np.asarray(mcmc_output["counts"])
    return int(counts.sum()) if is_type == 3 else int(counts.size)


def _is_correction_spdk(model, mcmc_output, N, it, seed, ot, draw_range):
    """approx_mcmc::is_correction_spdk (src/approx_mcmc.cpp:606-678): the correction phase of
    run_mcmc(..., sampling_method = "spdk") for ssm_ung models with an update_fn (parameter sets), theta output"""
    if getattr(model, "update_fn", None) is None or model.cls != "ssm_ung":
        raise ValueError("sampling_method 'spdk': ssm_ung models with an update_fn (theta -> dict of model components)")
    if ot != 3 or draw_range is not None:
        raise ValueError("sampling_method 'spdk': output_type 'theta' (3) and all draws in one call")
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    modes = np.asfortranarray(np.asarray(mcmc_output["modes"], dtype=np.float64).reshape(model.n, -1, order="F"))
    if modes.shape[1] != counts.size or np.shape(mcmc_output["theta"])[0] != counts.size:
        raise ValueError("theta, counts and modes disagree on the number of stored draws")
    ps = build_parset(model, np.asarray(mcmc_output["theta"], dtype=np.float64))
    al = mcmc_output.get("approx_loglik")
    al = None if al is None else _lib.f64(np.ravel(al))
    nd = n_draws(mcmc_output, it)
    w, post = np.zeros(nd), np.zeros(nd)
    ms = model.c_struct()
    check(lib().bssm_b200_is_correction_spdk_ung_parset(C.byref(ms), C.byref(ps), N, seed, it, uptr(counts), dptr(modes),
                                                        dptr(al), counts.size, dptr(w), dptr(post)))
    res = dict(mcmc_output)
    base = np.asarray(mcmc_output.get("posterior", np.zeros(counts.size)), dtype=np.float64)
    if it == 3:
        base = np.repeat(base, counts)
    res["weights"], res["posterior"], res["mcmc_type"] = w, base + post, "is%d" % it
    return res


def post_correct(model, mcmc_output, particles, threads=1, is_type="is2", seed=1, draw_range=None,
                 return_loglik=False, L=None, sampling_method="psi"):
    """post_correct, R/post_correction.R:245-308 (nongaussian branch).

    mcmc_output: dict with theta (n_stored x n_par), counts (n_stored), modes
    (n x n_stored), output_type (1 full / 2 summary / 3 theta) and posterior.
    `threads` is the number of GPUs of this process to use (the reference's OpenMP threads become one
    host thread per device inside the C call; results do not depend on it).
    draw_range = (begin, end) computes only that shard of the (expanded) draws.

    sampling_method = "bsf": the correction phase of run_mcmc(..., sampling_method = "bsf")
    (is_correction_bsf, src/approx_mcmc.cpp:447-518) for m = 1 ung-family models and the ssm_nlg
    registry models; mcmc_output carries approx_loglik (n_stored) and the weights are
    exp(loglik_bsf - approx_loglik).

    ssm_sde models: the correction phase of run_mcmc.ssm_sde (is_correction_bsf on the fine mesh
    L, src/approx_mcmc.cpp:522-591); mcmc_output then carries approx_loglik (n_stored) instead of modes.
    """
    from ._lib import PostcorrOut
    N, seed = _check_particles(model, particles, seed)
    it = {"is1": 1, "is2": 2, "is3": 3}[is_type.lower()]
    ot = int(mcmc_output.get("output_type", 3))
    if sampling_method == "spdk":
        return _is_correction_spdk(model, mcmc_output, N, it, seed, ot, draw_range)
    if sampling_method not in ("psi", "bsf"):
        raise ValueError("sampling_method must be 'psi', 'bsf' or 'spdk'")
    bsf = sampling_method == "bsf" and model.cls != "ssm_sde"
    use_parset = getattr(model, "update_fn", None) is not None and model.cls in ("ssm_ung", "ssm_mng")
    if model.cls == "ssm_mng" and not use_parset:
        raise ValueError("post_correct for ssm_mng needs model.update_fn (theta -> dict of model components)")
    if use_parset and not bsf:
        theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)
        counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
        rows = model.n * (model.p if model.cls == "ssm_mng" else 1)  # mode_storage: p x n per draw
        modes = np.asfortranarray(np.asarray(mcmc_output["modes"], dtype=np.float64).reshape(rows, -1, order="F"))
        if theta.shape[1] != counts.size or modes.shape[1] != counts.size:
            raise ValueError("theta, counts and modes disagree on the number of stored draws")
    elif model.cls == "ssm_sde" or bsf:
        if model.cls == "ssm_sde" and (L is None or int(L) < 1):
            raise ValueError("Discretization level L must be larger than 0.")
        theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)
        counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
        approx = _lib.f64(np.ravel(mcmc_output["approx_loglik"]))
        if theta.shape[1] != counts.size or approx.size != counts.size:
            raise ValueError("theta, counts and approx_loglik disagree on the number of stored draws")
        modes = None
    else:
        theta, counts, modes = _postcorr_inputs(model, mcmc_output)
    total = n_draws(mcmc_output, it)
    b, e = (0, total) if draw_range is None else draw_range
    nd = e - b
    n, m = model.n, model.m
    out = PostcorrOut()
    w, post = np.zeros(nd), np.zeros(nd)
    alpha = np.zeros((n + 1, m, nd), order="F") if ot == 1 else None
    ah = np.zeros((n + 1, m), order="F") if ot == 2 else None
    Vt = np.zeros((m, m, n + 1), order="F") if ot == 2 else None
    ll = np.zeros(nd) if return_loglik else None
    out.weights, out.posterior, out.alpha = dptr(w), dptr(post), dptr(alpha)
    out.alphahat, out.Vt, out.loglik = dptr(ah), dptr(Vt), dptr(ll)
    ms = model.c_struct()
    if use_parset:
        # postcorrection_nongaussian(model_type = 0 | 1): update_fn is evaluated here for every stored draw,
        # as the parset constructors do before the reference's OpenMP region (src/approx_mcmc.cpp:698, :1051)
        ps = build_parset(model, np.asarray(mcmc_output["theta"], dtype=np.float64))
        kind = "mng" if model.cls == "ssm_mng" else "ung"
        fn = getattr(lib(), ("bssm_b200_is_correction_bsf_%s_parset" if bsf else "bssm_b200_postcorrection_%s_parset") % kind)
        check(fn(C.byref(ms), C.byref(ps), ot, N, seed, int(threads), it, uptr(counts),
                 dptr(approx) if bsf else dptr(modes), counts.size, b, e, C.byref(out)))
    elif bsf and model.cls == "ssm_nlg":
        check(lib().bssm_b200_is_correction_bsf_nlg(
            C.byref(ms), ot, N, seed, int(threads), it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(approx), counts.size, b, e, C.byref(out)))
    elif bsf:
        tm = theta_map(model)
        check(lib().bssm_b200_is_correction_bsf(
            C.byref(ms), C.byref(tm), ot, N, seed, int(threads), it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(approx), counts.size, b, e, C.byref(out)))
    elif model.cls == "ssm_sde":
        check(lib().bssm_b200_is_correction_bsf_sde(
            C.byref(ms), int(L), ot, N, seed, int(threads), it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(approx), counts.size, b, e, C.byref(out)))
    elif model.cls == "ssm_nlg":  # post_correct -> postcorrection_nonlinear, R/post_correction.R:272-283
        check(lib().bssm_b200_postcorrection_nonlinear(
            C.byref(ms), ot, N, seed, int(threads), it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(modes), counts.size, b, e, C.byref(out)))
    else:
        tm = theta_map(model)
        check(lib().bssm_b200_postcorrection_nongaussian(
            C.byref(ms), C.byref(tm), ot, N, seed, int(threads), it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(modes), counts.size, b, e, C.byref(out)))
    res = dict(mcmc_output)
    res["weights"] = w
    base = np.asarray(mcmc_output.get("posterior", np.zeros(counts.size)), dtype=np.float64)
    if it == 3:
        base = np.repeat(base, counts)
    res["posterior"] = base[b:e] + post
    if ot == 1:
        res["alpha"] = alpha
    elif ot == 2:
        res["alphahat"], res["Vt"], res["sum_w"] = ah, Vt, out.sum_w
    if return_loglik:
        res["loglik"] = ll
    res["mcmc_type"] = "is%d" % it
    res["device_ms"], res["launches"] = out.device_ms, out.launches
    return res


class PostcorrJob(FilterJob):
    """Resident post-correction shard (bssm_b200_job_create_postcorr): weights only."""

    def __init__(self, model, mcmc_output, particles, is_type="is2", seed=1, draw_range=None, out_flags=0):
        N, seed = _check_particles(model, particles, seed)
        it = {"is2": 2, "is3": 3}[is_type.lower()]
        theta, counts, modes = _postcorr_inputs(model, mcmc_output)
        total = n_draws(mcmc_output, it)
        b, e = (0, total) if draw_range is None else draw_range
        self.model, self.N, self.n_filters = model, N, e - b
        self._h = C.c_void_p()
        self._ms = model.c_struct()
        if model.cls == "ssm_nlg":
            check(lib().bssm_b200_job_create_postcorr_nlg(
                C.byref(self._ms), N, seed, it, uptr(counts), dptr(theta), theta.shape[0], dptr(modes),
                counts.size, b, e, int(out_flags), C.byref(self._h)))
            return
        self._tm = theta_map(model)
        check(lib().bssm_b200_job_create_postcorr(
            C.byref(self._ms), C.byref(self._tm), N, seed, it, uptr(counts), dptr(theta), theta.shape[0],
            dptr(modes), counts.size, b, e, C.byref(self._h)))


def gaussian_approx_mode(model):
    """mode of the Gaussian approximation (ssm_ung::approximate), from model.initial_mode"""
    ms = model.c_struct()
    mode = np.zeros(model.n)
    check(lib().bssm_b200_approximate_mode(C.byref(ms), dptr(mode)))
    return mode


def synthetic_svm_mcmc_output(model, n_stored, output_type=3, jitter=0.05, gen_seed=7):
    """Synthetic stand-in for run_mcmc output on an svm model (BASELINE config 4):
    theta_i = theta * exp(jitter * N(0,1)) with rho clipped to (-.999, .999),
    counts ~ 1 + Poisson(3), modes_i = converged mode at theta_i (Newton sweeps
    warm-started from the mode at theta)."""
    from . import models
    rng = np.random.default_rng(gen_seed)
    th0 = np.asarray(model.theta)
    theta = th0[None, :] * np.exp(jitter * rng.normal(size=(n_stored, th0.size)))
    theta[:, 0] = np.clip(theta[:, 0], -0.999, 0.999)
    counts = (1 + rng.poisson(3, size=n_stored)).astype(np.uint32)
    base = gaussian_approx_mode(model)
    modes = np.zeros((model.n, n_stored), order="F")
    for i in range(n_stored):
        mi = models.svm(model.y, theta[i, 0], theta[i, 1], sigma=theta[i, 2])
        mi.initial_mode = base + 2 * np.log(th0[2] / theta[i, 2])  # the signal shifts with log sigma
        modes[:, i] = gaussian_approx_mode(mi)
    return {"theta": theta, "counts": counts, "modes": modes, "output_type": output_type,
            "posterior": np.zeros(n_stored)}


def gaussian_psi_smoother(model, particles, seed=1):
    """particle_smoother.lineargaussian(method = "psi") -> gaussian_psi_smoother, R/particle_smoother.R:90-110:
    `particles` independent draws from the smoothing distribution; returns (n+1) x m x N plus their summary."""
    N, seed = _check_particles(model, particles, seed)
    alpha = np.zeros((model.m, model.n + 1, N), order="F")
    ms = model.c_struct()
    check(lib().bssm_b200_gaussian_psi_smoother(C.byref(ms), N, seed, dptr(alpha)))
    return {"alpha": np.transpose(alpha, (1, 0, 2)), "alphahat": alpha.mean(axis=2).T,
            "Vt": np.stack([np.atleast_2d(np.cov(alpha[:, t, :], bias=True)) for t in range(model.n + 1)], axis=2)}
