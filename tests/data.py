"""Seeded synthetic inputs for the parity tests and the bench."""
import json
import os

import numpy as np

import bssm_b200

HERE = os.path.dirname(os.path.abspath(__file__))


def poisson_series():
    return np.array(json.load(open(os.path.join(HERE, "golden", "poisson_series.json")))["y"], dtype=np.float64)


def exchange():
    return np.array(json.load(open(os.path.join(HERE, "golden", "exchange.json")))["y"], dtype=np.float64)


def svm_series(n, rho=0.98, sd_ar=0.2, sigma=1.0, gen_seed=2024):
    """alpha_1 ~ N(0, sd^2/(1-rho^2)), alpha_{t+1} = rho alpha_t + sd eta, y = sigma e^{alpha/2} eps"""
    rng = np.random.default_rng(gen_seed)
    a = np.empty(n)
    a[0] = rng.normal() * sd_ar / np.sqrt(1 - rho ** 2)
    for t in range(1, n):
        a[t] = rho * a[t - 1] + sd_ar * rng.normal()
    return sigma * np.exp(a / 2) * rng.normal(size=n)


def svm_model(n, rho=0.98, sd_ar=0.2, sigma=1.0, gen_seed=2024, missing=()):
    y = svm_series(n, rho, sd_ar, sigma, gen_seed)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.svm(y, rho, sd_ar, sigma=sigma)


def poisson_bsm_model(n=100, missing=()):
    """the documented example of R/bootstrap_filter.R:78-82: bsm_ng on poisson_series"""
    y = poisson_series()[:n].copy()
    for t in missing:
        y[t] = np.nan
    return bssm_b200.bsm_ng(y, sd_level=0.1, sd_slope=0.01, P1=np.diag([1.0, 2.0]), distribution="poisson")


def svm_mcmc_output(model, n_stored, output_type=3, jitter=0.05, gen_seed=7):
    """Synthetic stand-in for run_mcmc(...) output on an svm model:
    theta draws = the model's theta * exp(jitter * N(0,1)) (rho clipped to (-.999, .999)),
    counts ~ 1 + Poisson(3), modes = the converged mode of the Gaussian approximation at each
    theta, computed by the ORACLE's approximate() (test infrastructure)."""
    import copy

    import orc
    rng = np.random.default_rng(gen_seed)
    th0 = np.asarray(model.theta)
    theta = th0[None, :] * np.exp(jitter * rng.normal(size=(n_stored, th0.size)))
    theta[:, 0] = np.clip(theta[:, 0], -0.999, 0.999)
    counts = (1 + rng.poisson(3, size=n_stored)).astype(np.uint32)
    modes = np.zeros((model.n, n_stored), order="F")
    for i in range(n_stored):
        mi = bssm_b200.svm(model.y, theta[i, 0], theta[i, 1], sigma=theta[i, 2])
        modes[:, i] = orc.psi_prepass(mi).mode
    return {"theta": theta, "counts": counts, "modes": modes, "output_type": output_type,
            "posterior": np.zeros(n_stored)}


def bsm_mcmc_output(model, n_stored, output_type=3, jitter=0.1, gen_seed=13):
    """synthetic run_mcmc stand-in for a bsm_ng model: theta_i = theta + jitter * N(0,1) on the log
    scale (regression coefficients: additive), modes = converged mode at each theta (oracle)."""
    import copy

    import orc
    rng = np.random.default_rng(gen_seed)
    th0 = np.asarray(model.theta)
    theta = th0[None, :] + jitter * rng.normal(size=(n_stored, th0.size))
    counts = (1 + rng.poisson(3, size=n_stored)).astype(np.uint32)
    modes = np.zeros((model.n, n_stored), order="F")
    ex = model.extra
    for i in range(n_stored):
        mi = copy.copy(model)
        mi.R = model.R.copy(); mi.P1 = model.P1.copy()
        j = 0
        mi.R[0, 0, 0] = np.exp(theta[i, j]); j += 1
        if ex["slope"]:
            mi.R[1, 1, 0] = np.exp(theta[i, j]); j += 1
        if ex["seasonal"]:
            r = 1 + int(ex["slope"])
            mi.R[r, r, 0] = np.exp(theta[i, j]); j += 1
        if ex["noise"]:
            c = 1 + int(ex["slope"]) + int(ex["seasonal"])
            mi.R[model.m - 1, c, 0] = np.exp(theta[i, j]); mi.P1[-1, -1] = mi.R[model.m - 1, c, 0] ** 2; j += 1
        if ex["phi_est"]:
            mi.phi = float(np.exp(theta[i, j])); j += 1
        if model.xreg is not None:
            mi.beta = theta[i, j:].copy()
        modes[:, i] = orc.approx_mode_m(mi)
    return {"theta": theta, "counts": counts, "modes": modes, "output_type": output_type,
            "posterior": np.zeros(n_stored)}


def growth_model(n=200, gen_seed=11, missing=()):
    """BASELINE config 3: the logistic growth model of vignettes/psi_pf.Rmd:133-158 (dT=.1, K=500,
    p1=50, r'1=-1.5, H=1, R1=.05, R2=1, a1=(-1.5, 50), P1=diag(1, 100)) in the packaged
    (log-theta) parameterisation of R/cpp_example_models.R:336-511."""
    rng = np.random.default_rng(gen_seed)
    dT, K, H, R1, R2 = 0.1, 500.0, 1.0, 0.05, 1.0
    r_, p_ = np.empty(n), np.empty(n)
    r_[0], p_[0] = -1.5, 50.0
    for t in range(1, n):
        r = np.exp(r_[t - 1]) / (1 + np.exp(r_[t - 1]))
        p_[t] = K * p_[t - 1] * np.exp(r * dT) / (K + p_[t - 1] * (np.exp(r * dT) - 1)) + R2 * rng.normal()
        r_[t] = r_[t - 1] + R1 * rng.normal()
    y = p_ + H * rng.normal(size=n)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_nlg(y, "growth", theta=np.log([H, R1, R2]),
                             known_params=[dT, K, -1.5, 50.0, 1.0, 100.0])


def ar_exp_model(n=80, gen_seed=21, missing=(), iekf_iter=0):
    """nlg_ar_exp (R/cpp_example_models.R:194-335): AR(1) state observed through its exponential"""
    rng = np.random.default_rng(gen_seed)
    mu, rho, sigma, H = 0.5, 0.9, 0.25, 0.4
    a = np.empty(n)
    a[0] = mu + sigma / np.sqrt(1 - rho ** 2) * rng.normal()
    for t in range(1, n):
        a[t] = mu * (1 - rho) + rho * a[t - 1] + sigma * rng.normal()
    y = np.exp(a) + H * rng.normal(size=n)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_nlg(y, "ar_exp", theta=[0.4, 0.85, np.log(0.3), np.log(0.5)], iekf_iter=iekf_iter)


def sin_exp_model(n=60, gen_seed=33, missing=()):
    """nlg_sin_exp (R/cpp_example_models.R:635-757): alpha_{t+1} = sin(alpha_t) + R eta, y = exp(alpha) + H eps"""
    rng = np.random.default_rng(gen_seed)
    H, R = 0.3, 0.5
    a = np.empty(n)
    a[0] = rng.normal()
    for t in range(1, n):
        a[t] = np.sin(a[t - 1]) + R * rng.normal()
    y = np.exp(a) + H * rng.normal(size=n)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_nlg(y, "sin_exp", theta=[np.log(0.35), np.log(0.45)])


def r_sin_exp_data(n=10, na=()):
    """test_ekpf.R:7-14 / :47-55: set.seed(1); y[1] <- rnorm(1, exp(x[1]), 0.1); then alternately
    x[i+1] <- rnorm(1, sin(x[i]), 0.1), y[i+1] <- rnorm(1, exp(x[i+1]), 0.1)"""
    from r_rng import RStream
    rs = RStream(1)
    x = np.zeros(n)
    y = np.zeros(n)
    y[0] = rs.rnorm(1, np.exp(x[0]), 0.1)[0]
    for i in range(n - 1):
        x[i + 1] = rs.rnorm(1, np.sin(x[i]), 0.1)[0]
        y[i + 1] = rs.rnorm(1, np.exp(x[i + 1]), 0.1)[0]
    for t in na:
        y[t] = np.nan
    return y


def gbm_model(n=50, gen_seed=42, missing=()):
    """BASELINE config 5: GBM as tests/testthat/test_sde.R:7-25 (mu=.05, sigma_x=.3, sigma_y=1, x0=1),
    filter theta = (.08, .4, 1.5), positive = TRUE."""
    rng = np.random.default_rng(gen_seed)
    mu, sx, sy, x = 0.05, 0.3, 1.0, 1.0
    dt = 1.0 / 64
    xs = np.empty(n)
    for k in range(n):
        for _ in range(64):
            x = x * np.exp((mu - 0.5 * sx ** 2) * dt + sx * np.sqrt(dt) * rng.normal())
        xs[k] = x
    y = np.log(xs) + sy * rng.normal(size=n)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_sde(y, "gbm", theta=[0.08, 0.4, 1.5], x0=1.0, positive=True)


def poisson_ou_model(n=50, gen_seed=8, missing=()):
    """Ornstein-Uhlenbeck state with Poisson observations, as vignettes/sde_model.Rmd simulates it
    (rho = 0.5, nu = 2, sigma = 0.3, x0 = nu, unit spacing, Euler on a fine mesh)."""
    rng = np.random.default_rng(gen_seed)
    rho, nu, sigma = 0.5, 2.0, 0.3
    x, dt = nu, 1.0 / 64
    y = np.empty(n)
    for k in range(n):
        for _ in range(64):
            x += rho * (nu - x) * dt + sigma * np.sqrt(dt) * rng.normal()
        y[k] = rng.poisson(np.exp(x))
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_sde(y, "poisson_ou", theta=[np.log(0.45), 1.9, np.log(0.35)], x0=2.0, positive=False)


def mng_model(n=60, gen_seed=5, missing=True):
    """a 3-variate ssm_mng (Poisson, negative binomial, Gaussian) driven by a 2-dimensional state,
    in the spirit of the 4-variate model of tests/testthat/test_basics.R:100-160"""
    rng = np.random.default_rng(gen_seed)
    a = np.zeros((n, 2))
    for t in range(1, n):
        a[t] = [a[t - 1, 0] + a[t - 1, 1] + 0.1 * rng.normal(), 0.95 * a[t - 1, 1] + 0.02 * rng.normal()]
    Z = np.array([[1.0, 0.0], [0.8, 0.5], [1.0, 1.0]])
    sig = a @ Z.T + np.array([1.0, 0.5, 0.0])
    y = np.column_stack([rng.poisson(np.exp(sig[:, 0])), rng.negative_binomial(3, 3 / (3 + np.exp(sig[:, 1]))),
                         sig[:, 2] + 0.5 * rng.normal(size=n)]).astype(float)
    if missing:
        y[3, 0] = np.nan
        y[7, :] = np.nan
        y[n - 1, 2] = np.nan
    T = np.array([[1.0, 1.0], [0.0, 0.95]])
    R = np.diag([0.1, 0.02])
    return bssm_b200.ssm_mng(y, Z, T, R, a1=[0.0, 0.0], P1=np.diag([1.0, 0.1]),
                             distribution=["poisson", "negative binomial", "gaussian"], phi=[1.0, 3.0, 0.5],
                             D=[1.0, 0.5, 0.0])


def growth_mcmc_output(model, n_stored, output_type=3, jitter=0.03, gen_seed=9):
    """synthetic run_mcmc stand-in for the growth model: theta_i = theta + jitter * N(0,1) (theta is on
    the log scale), counts ~ 1 + Poisson(2), modes_i = the oracle's converged mode at theta_i (m x n)"""
    import orc
    rng = np.random.default_rng(gen_seed)
    th0 = np.asarray(model.theta)
    theta = th0[None, :] + jitter * rng.normal(size=(n_stored, th0.size))
    counts = (1 + rng.poisson(2, size=n_stored)).astype(np.uint32)
    modes = np.zeros((model.m * model.n, n_stored), order="F")
    for i in range(n_stored):
        name = {0: "growth", 1: "ar_exp", 2: "sin_exp"}[model.model_id]
        if name == "ar_exp":
            theta[i, 1] = np.clip(theta[i, 1], -0.98, 0.98)
        mi = bssm_b200.ssm_nlg(model.y, name, theta=theta[i], known_params=model.known_params)
        modes[:, i] = orc.nlg_psi_filter(mi, 4, seed=1).mode.ravel(order="F")
    return {"theta": theta, "counts": counts, "modes": modes, "output_type": output_type,
            "posterior": np.zeros(n_stored)}


def lg_model(n=80, gen_seed=21, missing=()):
    """local linear trend Gaussian model (bsm_lg: level + slope), as in
    tests/testthat/test_particle_smoother.R:17-31"""
    rng = np.random.default_rng(gen_seed)
    lvl, slp, y = 0.0, 0.1, np.empty(n)
    for t in range(n):
        y[t] = lvl + rng.normal()
        lvl, slp = lvl + slp + 0.3 * rng.normal(), slp + 0.05 * rng.normal()
    for t in missing:
        y[t] = np.nan
    return bssm_b200.bsm_lg(y, sd_y=1.0, sd_level=0.3, sd_slope=0.05, P1=np.diag([4.0, 1.0]))


def linear_gaussian_model(n=60, gen_seed=17, missing=(), log_H=-0.3):
    """nlg_linear_gaussian (R/cpp_example_models.R:512-634): alpha_{t+1} = alpha_t + eta, y_t = alpha_t + e^{theta0} eps,
    a1 = 0, P1 = 1"""
    rng = np.random.default_rng(gen_seed)
    a = np.cumsum(np.concatenate([[rng.normal()], rng.normal(size=n - 1)]))
    y = a + np.exp(log_H) * rng.normal(size=n)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.ssm_nlg(y, "linear_gaussian", theta=[log_H])
