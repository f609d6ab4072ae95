"""Model constructors mirroring the reference's R constructors.

Each returns a plain ``Model`` whose fields are the entries of the R list the
reference's C++ constructor reads (/root/reference/src/model_ssm_ung.cpp:9-32),
with the same names, shapes (column-major) and defaults:

  svm      R/models.R:1183-1238
  bsm_ng   R/models.R:926-1114
  ar1_ng   R/models.R:1285-1370
  ssm_ung  general constructor (matrices supplied by the caller)

Priors are MCMC-side and out of scope here: where the R constructors take a
prior object the Python ones take the initial value.
"""
import ctypes as C

import numpy as np

from ._lib import UngModel, dptr, f64

DISTRIBUTIONS = {"svm": 0, "poisson": 1, "binomial": 2, "negative binomial": 3, "gamma": 4, "gaussian": 5}


def init_mode(y, u, distribution):
    """R/init_mode.R:3-31"""
    y = np.array(y, dtype=np.float64)
    u = np.broadcast_to(np.asarray(u, dtype=np.float64), y.shape)
    nan = np.isnan(y)
    if distribution == "poisson":
        v = y / u
        v[nan | (v < 0.1)] = 0.1
        return np.log(v)
    if distribution == "binomial":
        p = (np.where(nan, 0.5, y) + 0.5) / (u + 1)
        return np.log(p / (1 - p))
    if distribution == "gamma":
        v = y / u
        v[nan | (v < 1)] = 1
        return np.log(v)
    if distribution == "negative binomial":
        v = y / u
        v[nan | (v < 1 / 6)] = 1 / 6
        return np.log(v)
    if distribution == "gaussian":
        return y
    raise ValueError("distribution must be poisson, binomial, gamma, gaussian or negative binomial")


class Model:
    """An ssm_ung-family model: the R list, as numpy arrays."""

    def __init__(self, cls, y, Z, T, R, a1, P1, D, C_, u, phi, distribution, initial_mode,
                 xreg=None, beta=None, theta=None, max_iter=100, conv_tol=1e-8, local_approx=True,
                 extra=None, H=None):
        self.cls = cls
        self.y = f64(np.ravel(y))
        self.n = self.y.size
        self.a1 = f64(np.ravel(a1))
        self.m = self.a1.size
        m, n = self.m, self.n
        self.Z = f64(np.reshape(Z, (m, -1), order="F"))
        self.T = f64(np.reshape(T, (m, m, -1), order="F"))
        R = np.asarray(R, dtype=np.float64)
        self.k = R.shape[1] if R.ndim >= 2 else 1
        self.R = f64(np.reshape(R, (m, self.k, -1), order="F"))
        self.P1 = f64(np.reshape(P1, (m, m), order="F"))
        self.D = f64(np.ravel(D))
        self.C = f64(np.reshape(C_, (m, -1), order="F"))
        self.u = f64(np.broadcast_to(np.asarray(u, dtype=np.float64), (n,)))
        self.phi = float(phi)
        self.distribution = distribution
        self.initial_mode = f64(np.ravel(initial_mode))
        self.xreg = None if xreg is None else f64(xreg)
        self.beta = None if beta is None else f64(np.ravel(beta))
        self.theta = np.array([] if theta is None else theta, dtype=np.float64)
        self.max_iter = int(max_iter)
        self.conv_tol = float(conv_tol)
        self.local_approx = bool(local_approx)
        self.extra = dict(extra or {})
        self.H = None if H is None else f64(np.ravel(H))  # linear-Gaussian models only
        if self.H is not None and self.H.size not in (1, n):
            raise ValueError("H must have length 1 or n")
        for name, arr, last in (("Z", self.Z, 1), ("T", self.T, 2), ("R", self.R, 2), ("C", self.C, 1)):
            if arr.shape[last] not in (1, n):
                raise ValueError("%s: last dimension must be 1 or n" % name)
        if self.D.size not in (1, n):
            raise ValueError("D must have length 1 or n")

    @property
    def xbeta(self):
        if self.xreg is None or self.xreg.size == 0:
            return None
        # compute_xbeta, src/model_ssm_ung.h:83 -- column by column in a fixed order (no BLAS, no FMA), the
        # same sums update_model(theta) forms inside the library when beta is part of theta
        xr = self.xreg.reshape(self.n, -1)
        xb = np.zeros(self.n)
        for j in range(xr.shape[1]):
            xb = xb + xr[:, j] * self.beta[j]
        return f64(xb)

    def c_struct(self):
        """bssm_ung_model view of this model; keeps the arrays alive on the struct."""
        s = UngModel()
        s.n, s.m, s.k = self.n, self.m, self.k
        s.distribution = DISTRIBUTIONS[self.distribution]
        s.Ztv = int(self.Z.shape[1] > 1)
        s.Ttv = int(self.T.shape[2] > 1)
        s.Rtv = int(self.R.shape[2] > 1)
        s.Dtv = int(self.D.size > 1)
        s.Ctv = int(self.C.shape[1] > 1)
        xb = self.xbeta
        keep = [self.y, self.Z, self.T, self.R, self.a1, self.P1, self.D, self.C, xb, self.u, self.initial_mode]
        s.y, s.Z, s.T, s.R = dptr(self.y), dptr(self.Z), dptr(self.T), dptr(self.R)
        s.a1, s.P1, s.D, s.C = dptr(self.a1), dptr(self.P1), dptr(self.D), dptr(self.C)
        s.xbeta = dptr(xb) if xb is not None else None
        s.u = dptr(self.u)
        s.phi = self.phi
        s.max_iter, s.local_approx, s.conv_tol = self.max_iter, int(self.local_approx), self.conv_tol
        s.initial_mode = dptr(self.initial_mode)
        s.H = dptr(self.H) if self.H is not None else None
        s.Htv = int(self.H is not None and self.H.size > 1)
        keep.append(self.H)
        s._keep = keep
        return s


def ssm_ung(y, Z, T, R, a1, P1, distribution, phi=1.0, u=1.0, D=0.0, C=None, initial_mode=None,
            xreg=None, beta=None, **kw):
    y = np.ravel(np.asarray(y, dtype=np.float64))
    a1 = np.ravel(np.asarray(a1, dtype=np.float64))
    if C is None:
        C = np.zeros((a1.size, 1))
    if initial_mode is None:
        initial_mode = init_mode(y, u, distribution)
    update_fn = kw.pop("update_fn", None)
    model = Model("ssm_ung", y, Z, T, R, a1, P1, np.ravel(D), C, u, phi, distribution, initial_mode,
                  xreg=xreg, beta=beta, **kw)
    # R/models.R ssm_ung(update_fn = ...): theta -> list of the model components that depend on it
    model.update_fn = update_fn
    return model


def svm(y, rho, sd_ar, sigma=None, mu=None):
    """Stochastic volatility model, R/models.R:1183-1238."""
    if (sigma is None) == (mu is None):
        raise ValueError("Define either sigma or mu, but not both.")
    y = np.ravel(np.asarray(y, dtype=np.float64))
    if not -1 < rho < 1:
        raise ValueError("rho must be in (-1, 1)")
    if sd_ar <= 0:
        raise ValueError("sd_ar must be positive")
    base = np.log(np.maximum(1e-4, y ** 2))
    if mu is not None:
        svm_type, initial_mode, a1, phi = 1, base, mu, 1.0
        Cm = [[mu * (1 - rho)]]
        theta = [rho, sd_ar, mu]
    else:
        svm_type, initial_mode, a1, phi = 0, base - 2 * np.log(sigma), 0.0, sigma
        Cm = [[0.0]]
        theta = [rho, sd_ar, sigma]
    P1 = [[sd_ar ** 2 / (1 - rho ** 2)]]
    return Model("svm", y, [[1.0]], [[[rho]]], np.reshape([sd_ar], (1, 1, 1)), [a1], P1, [0.0], Cm, 1.0, phi,
                 "svm", initial_mode, theta=theta, extra={"svm_type": svm_type})


def ar1_ng(y, rho, sigma, mu, distribution, phi=1.0, u=1.0, xreg=None, beta=None):
    """Non-Gaussian model with AR(1) latent process, R/models.R:1285-1370."""
    y = np.ravel(np.asarray(y, dtype=np.float64))
    if distribution not in ("poisson", "binomial", "negative binomial", "gamma"):
        raise ValueError("bad distribution")
    if distribution in ("poisson", "binomial"):
        phi = 1.0
    return Model("ar1_ng", y, [[1.0]], [[[rho]]], np.reshape([sigma], (1, 1, 1)), [mu],
                 [[sigma ** 2 / (1 - rho ** 2)]], [0.0], [[mu * (1 - rho)]], u, phi, distribution,
                 init_mode(y, u, distribution), xreg=xreg, beta=beta,
                 theta=[rho, np.log(sigma), mu] + ([] if beta is None else list(np.ravel(beta))),
                 extra={"mu_est": 1, "phi_est": 0})


def bsm_ng(y, sd_level, sd_slope=None, sd_seasonal=None, sd_noise=None, distribution="poisson",
           phi=1.0, u=1.0, period=1, a1=None, P1=None, xreg=None, beta=None):
    """Non-Gaussian basic structural model, R/models.R:926-1114."""
    y = np.ravel(np.asarray(y, dtype=np.float64))
    if distribution not in ("poisson", "binomial", "negative binomial", "gamma"):
        raise ValueError("bad distribution")
    slope = sd_slope is not None
    seasonal = sd_seasonal is not None
    noise = sd_noise is not None
    if not seasonal:
        period = 1
    npar_R = 1 + int(slope) + int(seasonal) + int(noise)
    m = 1 + int(slope) + int(seasonal) * (period - 1) + int(noise)
    a1 = np.zeros(m) if a1 is None else np.ravel(np.asarray(a1, dtype=np.float64))
    P1 = np.diag(np.full(m, 100.0)) if P1 is None else np.array(P1, dtype=np.float64)
    Z = np.zeros((m, 1))
    Z[0, 0] = 1
    s0 = 1 + int(slope)  # 0-based row of the first seasonal state
    if seasonal:
        Z[s0, 0] = 1
    T = np.zeros((m, m))
    T[0, 0] = 1
    if slope:
        T[0:2, 1] = 1
    if seasonal:
        T[s0, s0:m] = -1  # R: T[(2 + slope), (2 + slope):m] <- -1
        for j in range(period - 2):
            T[s0 + 1 + j, s0 + j] = 1
    R = np.zeros((m, max(1, npar_R)))
    R[0, 0] = sd_level
    if slope:
        R[1, 1] = sd_slope
    if seasonal:
        R[s0, s0] = sd_seasonal
    if noise:
        P1[m - 1, m - 1] = sd_noise ** 2
        Z[m - 1, 0] = 1
        R[m - 1, max(1, R.shape[1] - 1) - 1] = sd_noise
    if distribution in ("poisson", "binomial"):
        phi = 1.0
    # theta as R/models.R:1092-1100 orders it: log sds of the estimated components, log phi, beta
    phi_est = distribution in ("negative binomial", "gamma")
    with np.errstate(divide="ignore"):  # a fixed sd of 0 (constant component) has log sd = -inf
        theta = [np.log(sd_level)] + ([np.log(sd_slope)] if slope else []) + ([np.log(sd_seasonal)] if seasonal else []) \
            + ([np.log(sd_noise)] if noise else []) + ([np.log(phi)] if phi_est else []) \
            + ([] if beta is None else list(np.ravel(beta)))
    extra = {"slope": slope, "seasonal": seasonal, "noise": noise, "level_est": True, "slope_est": slope,
             "seasonal_est": seasonal, "phi_est": phi_est}
    return Model("bsm_ng", y, Z, T.reshape(m, m, 1), R.reshape(m, R.shape[1], 1), a1, P1, [0.0],
                 np.zeros((m, 1)), u, phi, distribution, init_mode(y, u, distribution),
                 xreg=xreg, beta=beta, theta=theta, extra=extra)


class NlgModel:
    """ssm_nlg (R/models.R:1545-1570) with the user functions taken from the device
    registry instead of compiled function pointers: model="growth" is the
    nlg_growth example of R/cpp_example_models.R:336-511, model="ar_exp" the nlg_ar_exp
    example of :194-335 (theta = mu, rho, log sigma, log H; no known parameters), model="sin_exp"
    the nlg_sin_exp example of :635-757 (theta = log H, log R), model="linear_gaussian" the
    nlg_linear_gaussian example of :512-634 (theta = log H; a random walk observed with noise)."""
    REGISTRY = {"growth": (0, 2, 2, 1, 3, 6), "ar_exp": (1, 1, 1, 1, 4, 0),
                "sin_exp": (2, 1, 1, 1, 2, 0),
                "linear_gaussian": (3, 1, 1, 1, 1, 0)}  # id, m, k, p, n_theta, n_known

    def __init__(self, y, model, theta, known_params=(), iekf_iter=0, n_states=None, n_etas=None):
        self.cls = "ssm_nlg"
        if isinstance(model, (int, np.integer)) and int(model) >= 1000:
            # an id from nlg_register: user functors (bootstrap filter), dimensions as given there
            self.model_id, self.m, self.k, self.p, nth, nkn = int(model), int(n_states), int(n_etas), 1, 1, 0
        elif model in self.REGISTRY:
            self.model_id, self.m, self.k, self.p, nth, nkn = self.REGISTRY[model]
        else:
            raise ValueError("ssm_nlg: model must be one of %s or an id returned by nlg_register" % sorted(self.REGISTRY))
        self.y = f64(np.ravel(y))
        self.n = self.y.size
        self.theta = f64(np.ravel(theta))
        self.known_params = f64(np.ravel(known_params))
        if self.theta.size < nth or self.known_params.size < nkn:
            raise ValueError("ssm_nlg %s: needs %d parameters and %d known parameters" % (model, nth, nkn))
        self.max_iter, self.conv_tol = 100, 1e-8  # R/models.R:1567
        self.iekf_iter = int(iekf_iter)            # particle_smoother(..., iekf_iter), R/particle_smoother.R:192-228
        if self.known_params.size == 0:
            self.known_params = f64(np.zeros(1))   # never read; keeps the pointer valid

    def c_struct(self):
        from ._lib import NlgModel as S
        s = S()
        s.model_id, s.n, s.p, s.m, s.k = self.model_id, self.n, self.p, self.m, self.k
        s.y, s.theta, s.n_theta = dptr(self.y), dptr(self.theta), self.theta.size
        s.known_params, s.n_known = dptr(self.known_params), self.known_params.size
        s._keep = [self.y, self.theta, self.known_params]
        return s


class SdeModel:
    """ssm_sde (R/models.R:1637-1650); model="gbm" is the sde_gbm example of
    R/cpp_example_models.R:109-193 (theta = mu, sigma_x, sigma_y), model="poisson_ou" the
    sde_poisson_ou example of :24-108 (theta = log rho, nu, log sigma)."""
    REGISTRY = {"gbm": (0, 3), "poisson_ou": (1, 3)}

    def __init__(self, y, model, theta, x0, positive):
        self.cls = "ssm_sde"
        if isinstance(model, (int, np.integer)) and int(model) >= 1000:
            self.model_id, nth = int(model), 1  # an id from sde_register: user functors, any theta
        elif model in self.REGISTRY:
            self.model_id, nth = self.REGISTRY[model]
        else:
            raise ValueError("ssm_sde: model must be one of %s or an id returned by sde_register" % sorted(self.REGISTRY))
        self.m = 1
        self.y = f64(np.ravel(y))
        self.n = self.y.size
        self.theta = f64(np.ravel(theta))
        self.x0, self.positive = float(x0), bool(positive)
        if self.theta.size < nth:
            raise ValueError("ssm_sde %s: needs %d parameters" % (model, nth))

    def c_struct(self):
        from ._lib import SdeModel as S
        s = S()
        s.model_id, s.n = self.model_id, self.n
        s.y, s.theta, s.n_theta = dptr(self.y), dptr(self.theta), self.theta.size
        s.x0, s.positive = self.x0, int(self.positive)
        s._keep = [self.y, self.theta]
        return s


def nlg_register(source, n_states, n_etas):
    """Registers an ssm_nlg model given by Z_fn, H_fn, T_fn, R_fn, a1_fn, P1_fn as CUDA C++ source (the
    counterpart of the function pointers of R/models.R ssm_nlg; signatures in include/bssm_b200.h
    bssm_b200_nlg_register).  Bootstrap filter / smoother, the filter inside an MCMC loop and the
    bootstrap correction phase run on it; returns the id to pass as ssm_nlg(model=..., n_states, n_etas)."""
    import ctypes as C
    from ._lib import check, lib
    mid = C.c_int32(-1)
    check(lib().bssm_b200_nlg_register(source.encode(), int(n_states), int(n_etas), C.byref(mid)))
    return int(mid.value)


def ssm_nlg(y, model, theta, known_params=(), iekf_iter=0, n_states=None, n_etas=None):
    return NlgModel(y, model, theta, known_params, iekf_iter=iekf_iter, n_states=n_states, n_etas=n_etas)


def sde_register(source):
    """Registers an ssm_sde model given by its four functions as CUDA C++ source (the counterpart of the
    drift / diffusion / ddiffusion / obs_density pointers of R/models.R ssm_sde, compiled there with
    Rcpp::sourceCpp): `__device__ double drift(double x, const double* theta)`, `diffusion`, `ddiffusion`
    with the same signature and `__device__ double log_obs_density(double y, double x, const double* theta)`.
    Compiled at run time into the engine's filter kernels; returns the id to pass as ssm_sde(model=...)."""
    import ctypes as C
    from ._lib import check, lib
    mid = C.c_int32(-1)
    check(lib().bssm_b200_sde_register(source.encode(), C.byref(mid)))
    return int(mid.value)


def ssm_sde(y, model, theta, x0, positive):
    return SdeModel(y, model, theta, x0, positive)


class MngModel:
    """ssm_mng (R/models.R:538-600): p observation series with their own distributions,
    one linear-Gaussian state.  y, u: n x p; Z: p x m x (1|n); D: p x (1|n)."""
    CODES = {"poisson": 1, "binomial": 2, "negative binomial": 3, "gamma": 4, "gaussian": 5}

    def __init__(self, y, Z, T, R, a1, P1, distribution, phi=1.0, u=None, D=None, C=None,
                 max_iter=100, conv_tol=1e-8, local_approx=True):
        self.cls = "ssm_mng"
        self.max_iter, self.conv_tol, self.local_approx = int(max_iter), float(conv_tol), bool(local_approx)
        self.y = f64(np.atleast_2d(np.asarray(y, dtype=np.float64)))
        self.n, self.p = self.y.shape
        self.a1 = f64(np.ravel(a1))
        self.m = self.a1.size
        n, p, m = self.n, self.p, self.m
        self.Z = f64(np.reshape(Z, (p, m, -1), order="F"))
        self.T = f64(np.reshape(T, (m, m, -1), order="F"))
        R = np.asarray(R, dtype=np.float64)
        self.k = R.shape[1]
        self.R = f64(np.reshape(R, (m, self.k, -1), order="F"))
        self.P1 = f64(np.reshape(P1, (m, m), order="F"))
        self.D = f64(np.zeros((p, 1)) if D is None else np.reshape(D, (p, -1), order="F"))
        self.C = f64(np.zeros((m, 1)) if C is None else np.reshape(C, (m, -1), order="F"))
        self.u = f64(np.ones((n, p)) if u is None else np.broadcast_to(np.asarray(u, dtype=np.float64), (n, p)))
        if isinstance(distribution, str):
            distribution = [distribution] * p
        self.distribution = np.array([self.CODES[d] for d in distribution], dtype=np.int32)
        self.phi = f64(np.broadcast_to(np.asarray(phi, dtype=np.float64), (p,)))
        # R/models.R:576-579: column-wise init_mode
        names = {v: k for k, v in self.CODES.items()}
        self.initial_mode = f64(np.column_stack(
            [init_mode(self.y[:, j], self.u[:, j], names[int(self.distribution[j])]) for j in range(p)]))

    def c_struct(self):
        from ._lib import MngModel as S
        s = S()
        s.n, s.m, s.k, s.p = self.n, self.m, self.k, self.p
        s.Ztv, s.Ttv, s.Rtv = int(self.Z.shape[2] > 1), int(self.T.shape[2] > 1), int(self.R.shape[2] > 1)
        s.Dtv, s.Ctv = int(self.D.shape[1] > 1), int(self.C.shape[1] > 1)
        s.distribution = self.distribution.ctypes.data_as(C.POINTER(C.c_int32))
        s.y, s.u, s.Z, s.T, s.R = dptr(self.y), dptr(self.u), dptr(self.Z), dptr(self.T), dptr(self.R)
        s.a1, s.P1, s.D, s.C, s.phi = dptr(self.a1), dptr(self.P1), dptr(self.D), dptr(self.C), dptr(self.phi)
        s.max_iter, s.local_approx, s.conv_tol = self.max_iter, int(self.local_approx), self.conv_tol
        s.initial_mode = dptr(self.initial_mode)
        s._keep = [self.y, self.u, self.Z, self.T, self.R, self.a1, self.P1, self.D, self.C, self.phi, self.distribution,
                   self.initial_mode]
        return s


def ssm_mng(y, Z, T, R, a1, P1, distribution, phi=1.0, u=None, D=None, C=None, **kw):
    update_fn = kw.pop("update_fn", None)
    model = MngModel(y, Z, T, R, a1, P1, distribution, phi=phi, u=u, D=D, C=C, **kw)
    model.update_fn = update_fn  # R/models.R ssm_mng(update_fn = ...)
    return model


def ssm_ulg(y, Z, H, T, R, a1, P1, D=0.0, C=None):
    """General univariate linear-Gaussian model, R/models.R ssm_ulg (src/model_ssm_ulg.cpp:9-37)."""
    y = np.ravel(np.asarray(y, dtype=np.float64))
    a1 = np.ravel(np.asarray(a1, dtype=np.float64))
    if C is None:
        C = np.zeros((a1.size, 1))
    return Model("ssm_ulg", y, Z, T, R, a1, P1, np.ravel(D), C, 1.0, 1.0, "gaussian", np.zeros(y.size), H=H)


def bsm_lg(y, sd_y, sd_level, sd_slope=None, a1=None, P1=None):
    """Basic structural Gaussian model (level + optional slope), R/models.R bsm_lg."""
    y = np.ravel(np.asarray(y, dtype=np.float64))
    slope = sd_slope is not None
    m = 1 + int(slope)
    a1 = np.zeros(m) if a1 is None else np.ravel(np.asarray(a1, dtype=np.float64))
    P1 = np.diag(np.full(m, 100.0)) if P1 is None else np.array(P1, dtype=np.float64)
    Z = np.zeros((m, 1)); Z[0, 0] = 1
    T = np.zeros((m, m)); T[0, 0] = 1
    R = np.zeros((m, m)); R[0, 0] = sd_level
    if slope:
        T[0:2, 1] = 1
        R[1, 1] = sd_slope
    mdl = ssm_ulg(y, Z, [sd_y], T.reshape(m, m, 1), R.reshape(m, m, 1), a1, P1)
    mdl.cls = "bsm_lg"
    return mdl


def ar1_lg(y, rho, sigma, mu, sd_y):
    """Univariate Gaussian model with AR(1) latent process, R/models.R:1385-1470."""
    y = np.ravel(np.asarray(y, dtype=np.float64))
    mdl = ssm_ulg(y, [[1.0]], [sd_y], np.reshape([rho], (1, 1, 1)), np.reshape([sigma], (1, 1, 1)), [mu],
                  [[sigma ** 2 / (1 - rho ** 2)]], C=[[mu * (1 - rho)]])
    mdl.cls = "ar1_lg"
    return mdl
