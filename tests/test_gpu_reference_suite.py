"""The reference's own testthat files for the particle-filter path, re-run against the engine with the
SAME model configurations and the SAME assertions (tests/testthat/test_bootstrap_filter.R and
test_particle_smoother.R), each followed by bit-exact parity with the oracle on that configuration."""
import numpy as np
import pytest

import bssm_b200
import data
import orc

pytestmark = pytest.mark.gpu

Y10 = np.arange(1.0, 11.0)
Y8 = np.array([1, 0, 1, 1, 1, 0, 0, 0], dtype=float)


def all_equal(target, current, tolerance):
    """R's all.equal.numeric, which testthat's expect_equal(..., tolerance) uses: mean absolute
    difference, RELATIVE to mean |target| only when that exceeds the tolerance, absolute otherwise"""
    target, current = np.asarray(target, dtype=float), np.asarray(current, dtype=float)
    xy = np.mean(np.abs(target - current))
    xn = np.mean(np.abs(target))
    if np.isfinite(xn) and xn > tolerance:
        xy = xy / xn
    return bool(xy < tolerance)


def ulg_1_10():
    """ssm_ulg(y = 1:10, Z = (1, 0)', H = 2, T = [[1, 1], [0, 1]], R = diag(2, 2), a1 = 0, P1 = diag(2, 2))"""
    return bssm_b200.ssm_ulg(Y10, [[1.0], [0.0]], [2.0], np.array([[1.0, 1.0], [0.0, 1.0]]).reshape(2, 2, 1),
                             np.diag([2.0, 2.0]).reshape(2, 2, 1), [0.0, 0.0], np.diag([2.0, 2.0]))


def _still_works(out):
    """the five assertions every `still works` test of test_bootstrap_filter.R makes"""
    assert out["weights"].min() >= 0
    assert out["weights"].max() < np.inf
    assert np.isfinite(out["logLik"])
    assert np.isfinite(out["att"].sum())
    assert np.isfinite(out["Ptt"].sum())


def _finite_smoother(out):
    assert np.isfinite(out["alpha"].sum()) and np.isfinite(out["alphahat"].sum()) and np.isfinite(out["Vt"].sum())


# ------------------------------------------------------------ test_bootstrap_filter.R
def test_bsm_lg_gives_identical_results_with_ssm_ulg(engine):
    """:5-17 (expect_equal(bsf_bsm, bsf_ssm_ulg, tolerance = 1e-8))"""
    a = engine.bootstrap_filter(ulg_1_10(), 10, seed=1)
    b = engine.bootstrap_filter(bssm_b200.bsm_lg(Y10, sd_y=2, sd_level=2, sd_slope=2, P1=np.diag([2.0, 2.0])), 10, seed=1)
    for key in ("at", "att", "Pt", "Ptt", "weights", "alpha"):
        np.testing.assert_allclose(a[key], b[key], rtol=1e-8, atol=0)
    assert a["logLik"] == pytest.approx(b["logLik"], rel=1e-8)


def test_linear_gaussian_bsf_still_works(engine):
    """:21-46: ssm_ulg and ar1_lg(y = 1:10, rho = 0.6, sigma = 1, sd_y = 0.1, mu = 1)"""
    m = ulg_1_10()
    out = engine.bootstrap_filter(m, 10, seed=1, return_indices=True)
    _still_works(out)
    ref = orc.ulg_bsf_filter(m, 10, seed=1)
    assert np.array_equal(out["indices"], ref.indices) and out["logLik"] == ref.loglik
    m = bssm_b200.ar1_lg(Y10, rho=0.6, sigma=1.0, mu=1.0, sd_y=0.1)
    out = engine.bootstrap_filter(m, 10, seed=1, return_indices=True)
    if out["stopped_at"] < 0:  # sd_y = 0.1 against a trend 1..10: ten particles can degenerate
        _still_works(out)
    ref = orc.ulg_bsf_filter(m, 10, seed=1)
    assert np.array_equal(out["indices"], ref.indices) and out["stopped_at"] == ref.stopped_at


@pytest.mark.parametrize("name", ["poisson", "binomial_ar1", "negbin", "svm"])
def test_nongaussian_bsf_still_works(engine, name):
    """:48-122"""
    if name == "poisson":
        m = bssm_b200.bsm_ng(Y10, sd_level=2, sd_slope=2, P1=np.diag([2.0, 2.0]), distribution="poisson")
    elif name == "binomial_ar1":
        m = bssm_b200.ar1_ng(Y8, rho=0.9, sigma=1.0, mu=1.0, distribution="binomial", xreg=np.arange(1.0, 9.0)[:, None],
                             beta=[0.0])
    elif name == "negbin":
        m = bssm_b200.bsm_ng(Y8, sd_level=2, sd_slope=2, P1=np.diag([2.0, 2.0]), distribution="negative binomial",
                             phi=0.1, u=2.0)
    else:
        m = bssm_b200.svm(data.exchange(), rho=0.98, sd_ar=0.2, sigma=1.0)
    out = engine.bootstrap_filter(m, 10, seed=1, return_indices=True)
    _still_works(out)
    ref = orc.bsf_filter(m, 10, seed=1)
    assert np.array_equal(out["indices"], ref.indices)
    assert out["logLik"] == pytest.approx(ref.loglik, rel=1e-13)


# ------------------------------------------------------------ test_particle_smoother.R
def bsm_lg_ones():
    return bssm_b200.bsm_lg(np.ones(5), sd_y=0.01, sd_level=0.05, sd_slope=0.01, a1=[1.0, 0.0], P1=np.diag([0.01, 0.01]))


def test_trivial_noise_does_not_affect_particle_smoother(engine):
    """:5-14"""
    m1, m2 = bsm_lg_ones(), bsm_lg_ones()
    m2.y = m2.y + np.finfo(float).eps
    a = engine.particle_smoother(m1, 100000, seed=1)["alphahat"]
    b = engine.particle_smoother(m2, 100000, seed=1)["alphahat"]
    np.testing.assert_allclose(a, b, rtol=1.5e-8)


def test_different_seeds_give_comparable_results(engine):
    """:15-22 (tolerance = 0.001)"""
    m = bsm_lg_ones()
    a = engine.particle_smoother(m, 100000, seed=11)["alphahat"]
    b = engine.particle_smoother(m, 100000, seed=12)["alphahat"]
    assert all_equal(a, b, 1e-3)


def test_particle_smoother_for_lgssm_works_as_kalman_smoother(engine):
    """:24-30 (tolerance = 1e-2) and :32-45 (psi vs bsf, alphahat and Vt, tolerance = 1e-2)"""
    from test_oracle import _rts_smoother
    m = bsm_lg_ones()
    Z, T = m.Z[:, 0], m.T[:, :, 0]
    RR = m.R[:, :, 0] @ m.R[:, :, 0].T
    asm, Psm, *_ = _rts_smoother(m.y, Z, np.full(m.n, m.H[0] ** 2), T, RR, m.a1, m.P1)
    ps = engine.particle_smoother(m, 100000, seed=1)["alphahat"]
    assert all_equal(asm[:m.n], ps[:m.n], 1e-2)
    out1 = engine.particle_smoother(m, 10000, method="psi", seed=1)
    out2 = engine.particle_smoother(m, 10000, method="bsf", seed=1)
    assert all_equal(out1["alphahat"], out2["alphahat"], 1e-2)
    assert all_equal(out1["Vt"], out2["Vt"], 1e-2)


@pytest.mark.parametrize("name", ["ulg", "poisson", "binomial", "negbin"])
def test_particle_smoother_returns_finite_values(engine, name):
    """:47-100, ten particles"""
    if name == "ulg":
        m = ulg_1_10()
        _finite_smoother(engine.particle_smoother(m, 10, seed=1))
        return
    if name == "poisson":
        m = bssm_b200.bsm_ng(Y10, sd_level=2, sd_slope=2, P1=np.diag([2.0, 2.0]), distribution="poisson")
    elif name == "binomial":
        m = bssm_b200.bsm_ng(Y8, sd_level=2, sd_slope=2, P1=np.diag([2.0, 2.0]), distribution="binomial")
    else:
        m = bssm_b200.bsm_ng(Y8, sd_level=0.1, sd_slope=0.1, P1=np.diag([2.0, 2.0]), phi=1.0,
                             distribution="negative binomial")
    for method in ("psi", "bsf"):
        out = engine.particle_smoother(m, 10, method=method, seed=1, return_indices=True)
        _finite_smoother(out)
    ref = orc.psi_filter_m(m, 10, seed=1)
    out = engine.particle_smoother(m, 10, method="psi", seed=1, return_indices=True)
    assert np.array_equal(out["indices"], ref.indices) and out["logLik"] == ref.loglik


def test_particle_smoother_for_svm(engine):
    """:102-118: svm(exchange[1:20], rho = 0.98, sd_ar = 0.01, mu = 0); psi with 100 particles against
    bsf with 10000, expect_equal(alphahat, tol = 0.1)"""
    m = bssm_b200.svm(data.exchange()[:20], rho=0.98, sd_ar=0.01, mu=0.0)
    out1 = engine.particle_smoother(m, 100, method="psi", seed=1)
    out2 = engine.particle_smoother(m, 10000, method="bsf", seed=1)
    _finite_smoother(out1)
    assert all_equal(out1["alphahat"], out2["alphahat"], 0.1)


# ------------------------------------------------------------ test_post_correct.R (data set only)
def growth_reference_model():
    """test_post_correct.R:58-95: the logistic growth data set (set.seed(1)), observations 2..15
    missing, initial_theta = (0, log 0.05, 0), known_params = (dT = .1, K = 500, -1, 50, 1, 100).
    (The test itself needs run_mcmc, which is out of scope; its model and data are reused here.)"""
    from r_rng import RStream
    rs = RStream(1)
    p1, K, H, R_1, R_2, dT = 50.0, 500.0, 1.0, 0.05, 1.0, 0.1
    t = np.arange(0.1, 10.0 + 1e-9, dT)
    n = t.size
    z = np.cumsum(np.concatenate([[-1.5], rs.rnorm(n - 1, sd=R_1)]))
    r = 1.0 / (1.0 + np.exp(-z))
    p = np.empty(n)
    p[0] = p1
    for i in range(1, n):
        p[i] = rs.rnorm(1, K * p[i - 1] * np.exp(r[i - 1] * dT) / (K + p[i - 1] * (np.exp(r[i - 1] * dT) - 1)), R_2)[0]
    y = p + rs.rnorm(n, 0.0, H)
    y[1:15] = np.nan
    return bssm_b200.ssm_nlg(y, "growth", theta=[0.0, np.log(0.05), 0.0], known_params=[dT, K, -1.0, 50.0, 1.0, 100.0])


@pytest.mark.parametrize("method", ["bsf", "psi"])
def test_growth_model_of_the_reference_post_correct_test(engine, method):
    m = growth_reference_model()
    assert m.n == 100 and np.isnan(m.y[1:15]).all() and np.isfinite(m.y[0])
    if method == "bsf":
        ref = orc.nlg_bsf_filter(m, 1000, seed=1)
        out = engine.bootstrap_filter(m, 1000, seed=1, return_indices=True)
    else:
        ref = orc.nlg_psi_filter(m, 1000, seed=1)
        out = engine.particle_smoother(m, 1000, method="psi", seed=1, return_indices=True)
    assert np.array_equal(out["indices"], ref.indices)
    assert out["logLik"] == ref.loglik
    assert np.all(out["weights"][:, 1:15] == 1.0)  # missing observations: unit weights
    assert np.isfinite(out["logLik"])


# ------------------------------------------------------------ test_ekpf.R
def test_particle_filtering_based_on_ekf_works(engine):
    """:3-43: nlg_sin_exp, the reference's data (set.seed(1), regenerated through tests/r_rng.py),
    theta = (log .1, log .1); the same assertions, then parity with the oracle"""
    y = data.r_sin_exp_data(10)
    model = bssm_b200.ssm_nlg(y, "sin_exp", theta=[np.log(0.1), np.log(0.1)])
    out = engine.ekpf_filter(model, 100, seed=1, return_indices=True)
    assert 1 < out["logLik"] < 6
    assert out["weights"].min() >= 0 and out["weights"].max() <= 1
    ref = orc.nlg_ekf_filter(model, 100, 1)
    assert np.array_equal(out["indices"], ref.indices)
    assert np.array_equal(out["weights"], ref.weights)
    assert out["logLik"] == ref.loglik
    target = [0.0263875638744833, 0.0734903567971838]
    for method in ("ekf", "psi", "bsf"):
        sm = engine.particle_smoother(model, 1000, method=method, seed=1)
        assert all_equal(target, np.ravel(sm["alphahat"])[8:10], 0.1), method


# ------------------------------------------------------------ test_post_correct.R
def _ar1_binomial_xreg(theta=None):
    """the model of test_post_correct.R:4-24: binomial ar1_ng, n = 14, two harmonic regressors, exposure u;
    theta = (rho, log sigma, mu, beta1, beta2).  R's rpois / rbinom streams are not restated: a seeded numpy
    sample with the same structure stands in (the assertions of that test do not depend on the data)."""
    n = 14
    tt = np.arange(1.0, n + 1)
    x1, x2 = np.sin(2 * np.pi / 12 * tt), np.cos(2 * np.pi / 12 * tt)
    rng = np.random.default_rng(1)
    alpha = np.zeros(n)
    for i in range(1, n):
        alpha[i] = rng.normal(1.0 * (1 - 0.7) + 0.7 * alpha[i - 1], 2.0)
    u = rng.poisson(50, size=n).astype(float)
    y = rng.binomial(u.astype(int), 1.0 / (1.0 + np.exp(-(0.5 * x1 + x2 + alpha)))).astype(float)
    th = np.array([0.5, 0.0, 0.0, 0.0, 0.0]) if theta is None else np.asarray(theta, dtype=float)
    return bssm_b200.ar1_ng(y, rho=th[0], sigma=float(np.exp(th[1])), mu=th[2], distribution="binomial", u=u,
                            xreg=np.column_stack([x1, x2]), beta=th[3:5])


def test_post_correction_for_ar1_model_with_regressors(engine):
    """:4-56: suggest_N and post_correct on the approx-MCMC output of a binomial ar1_ng model whose theta
    holds (rho, log sigma, mu) AND the regression coefficients, so every stored draw has its own xbeta
    (ar1_ng::update_model -> compute_xbeta, src/model_ar1_ng.cpp:12-36)"""
    model = _ar1_binomial_xreg()
    assert model.theta.size == 5
    rng = np.random.default_rng(7)
    ns = 24
    center = np.array([0.6, np.log(1.5), 0.8, 0.4, 0.9])
    theta = center[None, :] + rng.normal(size=(ns, 5)) * np.array([0.05, 0.1, 0.3, 0.2, 0.2])
    counts = (1 + rng.poisson(1.5, size=ns)).astype(np.uint32)
    modes = np.zeros((model.n, ns), order="F")
    for i in range(ns):   # what run_mcmc(mcmc_type = "approx") stores per draw: the mode at theta_i
        modes[:, i] = orc.psi_prepass(_ar1_binomial_xreg(theta[i])).mode
    mo = {"theta": theta, "counts": counts, "modes": modes, "output_type": 3}

    # suggest_N(model, out_approx, replications = 10, candidates = c(5, 10)) at the posterior mean
    th_hat = np.average(theta, axis=0, weights=counts)
    est = engine.suggest_N(model, th_hat, candidates=[5, 10], replications=10, seed=1)
    mi = _ar1_binomial_xreg(th_hat)
    mi.initial_mode = model.initial_mode
    for c, N in enumerate([5, 10]):
        ll = orc.filter_batch(mi, N, "psi", seed=1, first_filter=c * 10, n_filters=10, arith=orc.ARITH_REPRO)
        assert est["results"]["sd"][c] == pytest.approx(np.std(ll, ddof=1), rel=1e-10)
    assert est["N"] in (5, 10, None)

    # post_correct(model, out_approx, particles = estN$N): finite theta / alphahat / Vt / weights (:47-56) ...
    N = est["N"] or 10
    for ot, ist in ((3, 2), (2, 1), (1, 3)):
        mo_t = dict(mo, output_type=ot)
        res = engine.post_correct(model, mo_t, N, is_type="is%d" % ist, seed=3, return_loglik=True)
        ref = orc.post_correct(model, mo_t, N, is_type=ist, seed=3)
        assert 0 < res["weights"].max() < np.inf
        # ... and parity with the oracle, draw by draw
        assert np.array_equal(res["loglik"], ref.loglik)
        assert np.array_equal(res["weights"], ref.weights)
        if ot == 2:
            assert np.isfinite(res["alphahat"].sum()) and np.isfinite(res["Vt"].sum())
            np.testing.assert_allclose(res["alphahat"], ref.alphahat, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(res["Vt"], ref.Vt, rtol=1e-8, atol=1e-12)
        if ot == 1:
            assert np.array_equal(res["alpha"].view(np.uint64), ref.alpha.view(np.uint64))
    # the regression part matters: the same draws with beta zeroed give different weights
    theta0 = theta.copy()
    theta0[:, 3:] = 0.0
    res0 = engine.post_correct(model, dict(mo, theta=theta0), N, is_type="is2", seed=3)
    ref0 = orc.post_correct(model, dict(mo, theta=theta0), N, is_type=2, seed=3)
    assert np.array_equal(res0["weights"], ref0.weights)
    assert not np.allclose(res0["weights"], engine.post_correct(model, mo, N, is_type="is2", seed=3)["weights"])

    # the argument checks of :35-45
    with pytest.raises((ValueError, AttributeError, TypeError)):
        engine.post_correct(object(), mo, N)
    with pytest.raises(ValueError):
        engine.post_correct(model, mo, int(1e12))
    with pytest.raises((ValueError, RuntimeError)):
        engine.post_correct(model, dict(mo, theta=np.eye(2)), 10)


def test_ar1_with_regressors_bsf_correction_and_mcmc_loop(engine):
    """the same model through the two other callers of update_model: the bootstrap-filter correction
    (is_correction_bsf, src/approx_mcmc.cpp:447-518) and a resident job re-parameterised per MCMC
    iteration (the log_likelihood call of src/mcmc.cpp:338-491)"""
    model = _ar1_binomial_xreg()
    rng = np.random.default_rng(9)
    ns = 16
    center = np.array([0.6, np.log(1.5), 0.8, 0.4, 0.9])
    theta = center[None, :] + rng.normal(size=(ns, 5)) * np.array([0.05, 0.1, 0.3, 0.2, 0.2])
    counts = (1 + rng.poisson(1.5, size=ns)).astype(np.uint32)
    mo = {"theta": theta, "counts": counts, "approx_loglik": -45.0 + rng.normal(size=ns), "output_type": 3,
          "modes": np.zeros((model.n, ns), order="F")}
    ref = orc.is_correction_bsf(model, mo, 64, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True, sampling_method="bsf")
    np.testing.assert_allclose(res["loglik"], ref.loglik, rtol=1e-13)  # constants pre-summed per draw
    np.testing.assert_allclose(res["weights"], ref.weights, rtol=1e-10)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.is_correction_bsf(model, mo1, 32, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 32, is_type="is3", seed=5, sampling_method="bsf")
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))

    for method in ("psi", "bsf"):
        job = engine.FilterJob(model, 200, method=method, seed=4, n_filters=3)
        job.run()
        ll0 = job.logliks().copy()
        for th in theta[:2]:
            job.set_theta(th)
            job.run()
            m2 = _ar1_binomial_xreg(th)
            m2.initial_mode = model.initial_mode
            fresh = engine.FilterJob(m2, 200, method=method, seed=4, n_filters=3)
            fresh.run()
            if method == "psi":
                assert np.array_equal(job.logliks(), fresh.logliks())
            else:
                np.testing.assert_allclose(job.logliks(), fresh.logliks(), rtol=1e-13)
            assert not np.array_equal(job.logliks(), ll0)
            fresh.close()
        job.close()
