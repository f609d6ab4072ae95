"""The reference's OWN known-answer tests for pieces of the hot path, reproduced without R.

helske/bssm's testthat files pin a few numerical results whose inputs come from R's random number
stream (`set.seed()` + `rnorm`).  tests/r_rng.py restates that stream (Mersenne-Twister with R's
seeding, normals by inversion), so these golden vectors can be regenerated here and held against the
oracle (CPU, this file) and the product (tests/test_gpu_golden.py):

  * tests/testthat/test_approx.R:38-53   Gaussian approximation of an svm model: pseudo-observations,
    their standard deviations and the smoothed mode after two Newton sweeps, 15 digits
  * tests/testthat/test_basics.R:176-201 dmvnorm log-density (the density behind the ssm_nlg weights)
  * tests/testthat/test_sde.R:7-25,55-58 the GBM data set and its log-likelihood "about -17"
"""
import ctypes as C

import numpy as np
import pytest

import bssm_b200
import data
import orc
from r_rng import RStream

# ---- golden values copied from the reference's test files (cited above)
SVM_APPROX_Y = [-1.47548927809174, -11.2190916117862, 0.263154138901814, -121.519769682058, -36.0386937004332]
SVM_APPROX_H = [2.01061310553144, 4.84658294043645, 0.712674409714633, 15.6217737012134, 8.54936618861792]
SVM_SMOOTHED = [-0.0999179077423753, -0.101594935319188, -0.0985572218431492, -0.103275329248674,
                -0.103028083292436]
DMVNORM_LOGP = -14.0607446337904


def svm_kat_model():
    """test_approx.R:39-44: svm(rnorm(5), sigma = uniform(1, 0, 10), rho = uniform(0.95, 0, 1),
    sd_ar = uniform(0.1, 0, 1)); gaussian_approx(model, max_iter = 2, conv_tol = 1e-8)"""
    y = RStream(123).rnorm(5)
    m = bssm_b200.svm(y, rho=0.95, sd_ar=0.1, sigma=1.0)
    m.max_iter, m.conv_tol = 2, 1e-8
    return m


def gbm_kat_model():
    """test_sde.R:7-25"""
    rs = RStream(42)
    n, dt, mu, sigma_x, sigma_y, x = 50, 1.0, 0.05, 0.3, 1.0, 1.0
    y = np.empty(n)
    for k in range(n):
        x = x * np.exp((mu - 0.5 * sigma_x ** 2) * dt + np.sqrt(dt) * rs.rnorm(1, sd=sigma_x)[0])
        y[k] = rs.rnorm(1, np.log(x), sigma_y)[0]
    return bssm_b200.ssm_sde(y, "gbm", theta=[0.08, 0.4, 1.5], x0=1.0, positive=True)


def test_r_stream_restatement():
    """values every R user has seen: set.seed(123); rnorm(5) etc."""
    np.testing.assert_allclose(RStream(123).rnorm(5), [-0.560475646552213, -0.23017748948328, 1.55870831414912,
                                                       0.070508391424576, 0.129287735160946], rtol=1e-13)
    np.testing.assert_allclose(RStream(1).rnorm(3), [-0.626453810742332, 0.183643324222082, -0.835628612410047],
                               rtol=1e-13)
    np.testing.assert_allclose(RStream(42).runif(3), [0.914806043496355, 0.937075413297862, 0.286139534786344],
                               rtol=1e-13)


def test_svm_gaussian_approximation_matches_reference_golden_vector():
    """pins svm model construction (initial mode, P1), laplace_iter, the Kalman fast smoother and the
    fact that approximate() leaves the pseudo-observations linearised at the last-but-one mode"""
    m = svm_kat_model()
    r = orc.psi_prepass(m)
    np.testing.assert_allclose(r.ytilde, SVM_APPROX_Y, rtol=2e-14)
    np.testing.assert_allclose(r.Htilde, SVM_APPROX_H, rtol=2e-14)
    np.testing.assert_allclose(r.mode, SVM_SMOOTHED, rtol=2e-14)
    # the general-m restatement (the product's host path for m > 1) gives the same numbers at m = 1
    r2 = orc.psi_filter_m(m, 8, seed=1)
    np.testing.assert_allclose(orc.approx_mode_m(m), SVM_SMOOTHED, rtol=2e-14)
    assert r2.approx_loglik == pytest.approx(r.approx_loglik, rel=1e-13)


def test_dmvnorm_matches_reference_golden_value():
    f = orc.L().orc_dmvnorm_logpdf
    f.restype = C.c_double
    f.argtypes = [orc.c_double_p, orc.c_double_p, orc.c_double_p, C.c_int]
    g = orc.L().orc_dnorm1
    g.restype = C.c_double
    g.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    # test_basics.R:178-181: dmvnorm(1, 3, 2, lwr = TRUE) == dnorm(1, 3, 2, log = TRUE)
    for sd in (2.0, 4.0):
        ref = -0.5 * np.log(2 * np.pi) - np.log(sd) - 0.5 * ((1 - 3) / sd) ** 2
        assert g(1.0, 3.0, sd, 1) == pytest.approx(ref, rel=1e-14)
        assert g(1.0, 3.0, sd * sd, 0) == pytest.approx(ref, rel=1e-14)
    # test_basics.R:183-186: set.seed(1); a <- crossprod(matrix(rnorm(9), 3, 3))
    A = RStream(1).rnorm(9).reshape(3, 3, order="F")
    a = np.asfortranarray(A.T @ A)
    x = np.array([1.0, 2.0, 3.0])
    mean = -0.1 * np.array([3.0, 2.0, 1.0])
    assert f(orc.dptr(x), orc.dptr(mean), orc.dptr(a), 3) == pytest.approx(DMVNORM_LOGP, abs=1e-12)


def gbm_exact_loglik(m):
    """log x of a geometric Brownian motion is a Gaussian random walk with drift, and the observations
    are N(log x, sigma_y^2): the exact log-likelihood is a scalar Kalman filter on z = log x."""
    mu, sx, sy = m.theta
    drift, q, h = mu - 0.5 * sx * sx, sx * sx, sy * sy
    a, p, ll = np.log(m.x0), 0.0, 0.0
    for yk in m.y:
        a, p = a + drift, p + q
        if np.isfinite(yk):
            f = p + h
            v = yk - a
            ll += -0.5 * (np.log(2 * np.pi * f) + v * v / f)
            k = p / f
            a, p = a + k * v, p * (1 - k)
    return ll


def test_gbm_loglik_reference_assertion_and_exact_value():
    """test_sde.R:55-58 asserts `expect_equal(ll, -17, tol = 1)` for logLik(model, 10000, L = 3): with
    all.equal's relative difference |ll + 17| / |ll| < 1 that only says ll < -8.5.  The same assertion is
    made here on the regenerated data set, and the estimate is ALSO held against the exact value: the
    log of a GBM is a Gaussian random walk, so a Kalman filter gives the likelihood in closed form
    (the Milstein filter approaches it as L grows)."""
    m = gbm_kat_model()
    exact = gbm_exact_loglik(m)
    ll3 = np.mean([orc.sde_bsf_filter(m, 10000, 3, seed=s).loglik for s in (1, 2)])
    assert abs(ll3 + 17.0) / abs(ll3) < 1.0          # the reference's own assertion
    ll6 = np.mean([orc.sde_bsf_filter(m, 10000, 6, seed=s).loglik for s in (1, 2)])
    assert abs(ll6 - exact) < 0.15                   # discretisation + Monte Carlo error
    assert abs(ll3 - exact) < 0.5


def test_sin_exp_ekf_loglik_matches_reference_known_answers():
    """test_ekpf.R:64-69: for the nlg_sin_exp model with y[2:3] <- NA and theta = (log .1, log .1),
    `ekf(model)$logLik` is 2.65163101109689, with iekf_iter = 1 it is 2.61650080342709, and one
    iteration of the iterated EKF already equals two.  The EKF is the start of the psi-APF
    approximation and the proposal of the EKF particle filter (include/bssm_b200/nlg_gauss_approx.h)."""
    y = data.r_sin_exp_data(10, na=(1, 2))
    model = bssm_b200.ssm_nlg(y, "sin_exp", theta=[np.log(0.1), np.log(0.1)])
    ll0, _ = orc.nlg_ekf_loglik(model, 0)
    ll1, att1 = orc.nlg_ekf_loglik(model, 1)
    ll2, att2 = orc.nlg_ekf_loglik(model, 2)
    assert ll0 == pytest.approx(2.65163101109689, rel=1.5e-8)   # expect_equal's default tolerance
    assert ll1 == pytest.approx(2.61650080342709, rel=1.5e-8)
    assert ll2 == pytest.approx(ll1, rel=1.5e-8)
    np.testing.assert_allclose(att2, att1, rtol=1.5e-8)


def _glm_irls(y, X, family, u=None, phi=None, offset=None, iters=100):
    """maximum likelihood of a canonical/log-link GLM by Newton-Raphson on the score (numpy, independent
    of the engine and of the oracle): what R's glm() / glm.nb() converge to"""
    beta = np.zeros(X.shape[1])
    off = np.zeros(len(y)) if offset is None else offset
    for _ in range(iters):
        eta = X @ beta + off
        if family == "poisson":
            mu = np.exp(eta); score = y - mu; w = mu
        elif family == "binomial":
            pr = 1.0 / (1.0 + np.exp(-eta)); score = y - u * pr; w = u * pr * (1 - pr)
        else:  # negative binomial, log link, fixed size phi
            mu = np.exp(eta); score = (y - mu) * phi / (phi + mu); w = mu * phi / (phi + mu)
        step = np.linalg.solve(X.T @ (w[:, None] * X), X.T @ score)
        beta = beta + step
        if np.max(np.abs(step)) < 1e-13:
            break
    return beta


def test_poisson_glm_is_the_fixed_point_of_the_approximation():
    """test_approx.R:60-108 (`results for poisson GLM are equal to glm function`): with constant states
    (R = 0) and a diffuse prior the mode of the Gaussian approximation is the GLM's linear predictor.
    The engine holds m <= 3, so the check uses the reference's own one-state formulations of that test:
    model_poisson4 (ssm_ung with the regression part as D, :86-88) and model_poisson2 (bsm_ng with
    xreg and beta, :75-77), which the reference asserts give identical approximations (:97-107)."""
    counts = np.array([18, 17, 15, 20, 10, 20, 25, 13, 12], dtype=float)
    outcome = np.tile([0, 1, 2], 3)
    treatment = np.repeat([0, 1, 2], 3)
    X = np.column_stack([np.ones(9), outcome == 1, outcome == 2, treatment == 1, treatment == 2]).astype(float)
    coef = _glm_irls(counts, X, "poisson")
    # published glm fit of this classic data set (Dobson 1990; ?glm example): intercept 3.045, outcome2 -0.4543
    assert coef[0] == pytest.approx(3.044522, abs=2e-6) and coef[1] == pytest.approx(-0.4542553, abs=2e-6)
    m4 = bssm_b200.ssm_ung(counts, 1.0, 1.0, 0.0, [0.0], [[1e7]], "poisson", D=X[:, 1:] @ coef[1:])
    m2 = bssm_b200.bsm_ng(counts, sd_level=0.0, xreg=X[:, 1:], beta=coef[1:], P1=[[1e7]], distribution="poisson")
    p4, p2 = orc.psi_prepass(m4), orc.psi_prepass(m2)
    np.testing.assert_allclose(p4.alphahat[:9], coef[0], rtol=0, atol=1e-5)   # smoother(model)$alphahat[1, ]
    np.testing.assert_allclose(p4.mode, X @ coef, rtol=0, atol=1e-5)
    np.testing.assert_allclose(p2.ytilde, p4.ytilde, rtol=1.5e-8)             # expect_equivalent(...$y, ...$y)
    np.testing.assert_allclose(p2.Htilde, p4.Htilde, rtol=1.5e-8)             # ...$H


def test_binomial_and_negative_binomial_glm_fixed_points():
    """test_approx.R:110-143: binomial GLM (u = 20 trials, ?glm's budworm data) and a negative binomial
    GLM with exposure u, phi fixed: the smoothed state of the approximating model is the GLM intercept.
    One-state formulations as above (the quine data of :129-142 is MASS's, not the reference's: a seeded
    negative binomial sample with the same structure stands in)."""
    ldose = np.tile(np.arange(6.0), 2)
    numdead = np.array([1, 4, 9, 13, 18, 20, 0, 2, 6, 10, 12, 16], dtype=float)
    male = np.repeat([1.0, 0.0], 6)
    X = np.column_stack([np.ones(12), male, ldose, male * ldose])
    coef = _glm_irls(numdead, X, "binomial", u=20.0)
    assert coef[0] == pytest.approx(-2.9935, abs=1e-4)  # ?glm: sexF intercept of the budworm fit
    mb = bssm_b200.ssm_ung(numdead, 1.0, 1.0, 0.0, [0.0], [[1e5]], "binomial", u=np.full(12, 20.0),
                           D=X[:, 1:] @ coef[1:])
    np.testing.assert_allclose(orc.psi_prepass(mb).alphahat[:12], coef[0], rtol=0, atol=1e-4)

    rng = np.random.default_rng(123)
    n, phi = 146, 1.3
    offs = rng.integers(10, 21, size=n).astype(float) + rng.integers(1, 60, size=n)
    lam = offs * np.exp(-1.2)
    y = rng.negative_binomial(phi, phi / (phi + lam)).astype(float)
    b0 = _glm_irls(y, np.ones((n, 1)), "nb", phi=phi, offset=np.log(offs))[0]
    mnb = bssm_b200.bsm_ng(y, sd_level=0.0, u=offs, P1=[[1e7]], phi=phi, distribution="negative binomial")
    mnb.conv_tol = 1e-12
    np.testing.assert_allclose(orc.psi_prepass(mnb).alphahat[0], b0, rtol=1.5e-8)


def test_two_iid_series_give_the_same_approximation_as_two_univariate_models():
    """test_approx.R:145-160: ssm_mng with two independent negative binomial series (phi = 2, Z = T = I,
    R = 0.5 I) against two ssm_ung models, conv_tol = 1e-12, tolerance 1e-6"""
    rng = np.random.default_rng(1)
    y = rng.binomial(10, 1.0 / (1.0 + np.exp(-0.5 * rng.normal(size=20)))).astype(float).reshape(10, 2, order="F")
    mm = bssm_b200.ssm_mng(y, np.eye(2).reshape(2, 2, 1), np.eye(2).reshape(2, 2, 1),
                            (0.5 * np.eye(2)).reshape(2, 2, 1), [0.0, 0.0], np.eye(2),
                            ["negative binomial"] * 2, phi=[2.0, 2.0], conv_tol=1e-12)
    _, ah, mode, _ = orc.mng_approx(mm)
    for j in range(2):
        mu = bssm_b200.ssm_ung(y[:, j], 1.0, 1.0, 0.5, [0.0], [[1.0]], "negative binomial", phi=2.0, conv_tol=1e-12)
        pj = orc.psi_prepass(mu)
        np.testing.assert_allclose(mode[j], pj.mode, rtol=0, atol=1e-6)
        np.testing.assert_allclose(ah[j], pj.alphahat, rtol=0, atol=1e-6)


def test_sin_exp_approximation_converges_in_two_sweeps():
    """test_approx.R:181-199: nlg_sin_exp, n = 30, y[2:5] <- NA: gaussian_approx(model) equals
    gaussian_approx(model, max_iter = 2)"""
    y = data.r_sin_exp_data(30, na=(1, 2, 3, 4))
    model = bssm_b200.ssm_nlg(y, "sin_exp", theta=[np.log(0.1), np.log(0.1)])
    full = orc.nlg_psi_filter(model, 8, 1)
    two = orc.nlg_psi_filter(model, 8, 1, max_iter=2)
    np.testing.assert_allclose(two.mode, full.mode, rtol=1.5e-8, atol=1e-12)
    assert two.approx_loglik == pytest.approx(full.approx_loglik, rel=1.5e-8)


def test_gbm_smoothed_states_match_the_published_posterior_means():
    """test_sde.R:30-34: posterior means theta = (0.053, 0.253, 1.058), x_1 = 1.254, x_50 = 2.960 (from the
    paper of the example), tolerance 0.1: the bootstrap particle smoother at the published theta, on the
    regenerated data set, gives those two state means within the reference's tolerance"""
    m = gbm_kat_model()
    mm = bssm_b200.ssm_sde(m.y, "gbm", theta=[0.053, 0.253, 1.058], x0=1.0, positive=True)
    got = []
    for s in (1, 2, 3):
        r = orc.sde_bsf_filter(mm, 5000, 4, seed=s)
        a = r.alpha.copy()
        out = orc.filter_smoother(a, r.indices)
        a = a if out is None else out
        got.append(np.ravel(a.mean(axis=-1))[[0, 49]])
    got = np.mean(got, axis=0)
    target = np.array([1.254, 2.960])
    assert np.mean(np.abs(got - target)) / np.mean(np.abs(target)) < 0.1
