"""The product against the reference's own golden values (see tests/test_reference_golden.py for where
they come from and how their inputs are regenerated without R)."""
import numpy as np
import pytest

import bssm_b200
import orc
from test_reference_golden import SVM_SMOOTHED, gbm_exact_loglik, gbm_kat_model, svm_kat_model

pytestmark = pytest.mark.gpu


def test_svm_approximation_golden_vector(engine):
    """tests/testthat/test_approx.R:38-53 through the product: the host Newton sweeps
    (bssm_b200_approximate_mode) reproduce the pinned smoothed mode, and the device pre-pass built on it
    gives the oracle's approximate log-likelihood (whose inputs the same golden vector pins)"""
    m = svm_kat_model()
    np.testing.assert_allclose(engine.gaussian_approx_mode(m), SVM_SMOOTHED, rtol=2e-14)
    ref = orc.psi_filter(m, 64, seed=2)
    res = engine.particle_smoother(m, 64, method="psi", seed=2, return_indices=True)
    assert res["approx_logLik"] == pytest.approx(ref.approx_loglik, rel=1e-13)
    assert np.array_equal(res["indices"], ref.indices)
    assert res["logLik"] == ref.loglik


def test_gbm_loglik_against_exact_value(engine):
    """tests/testthat/test_sde.R:7-25, 55-58 data set: the Milstein bootstrap filter on the device
    against the closed-form likelihood (log GBM = Gaussian random walk)"""
    m = gbm_kat_model()
    exact = gbm_exact_loglik(m)
    ll3 = engine.logLik(m, 10000, seed=1, L=3)
    assert abs(ll3 + 17.0) / abs(ll3) < 1.0  # the reference's own (loose) assertion
    ll6 = np.mean([engine.logLik(m, 100000, seed=s, L=6) for s in (1, 2)])
    assert abs(ll6 - exact) < 0.1
    out = engine.bootstrap_filter(m, 1000, seed=3, L=3)
    assert abs(out["logLik"] - ll3) / abs(ll3) < 1.0  # test_sde.R:57-58


def test_gbm_smoothed_states_against_the_published_posterior_means(engine):
    """test_sde.R:30-34 holds the IS-MCMC posterior means against the values of the paper the example
    comes from: theta = (0.053, 0.253, 1.058), states x_1 = 1.254 and x_50 = 2.960, tolerance 0.1
    (relative).  Plugging the published theta into the particle smoother on the regenerated data set
    reproduces the two state means within that tolerance (the MCMC itself stays in R)."""
    m = gbm_kat_model()
    mm = bssm_b200.ssm_sde(m.y, "gbm", theta=[0.053, 0.253, 1.058], x0=1.0, positive=True)
    target = np.array([1.254, 2.960])
    got = np.mean([np.ravel(engine.particle_smoother(mm, 50000, method="bsf", seed=s, L=5)["alphahat"])[[0, 49]]
                   for s in (1, 2)], axis=0)
    assert np.mean(np.abs(got - target)) / np.mean(np.abs(target)) < 0.1   # all.equal's measure
    assert np.all(np.abs(got / target - 1.0) < 0.1)
