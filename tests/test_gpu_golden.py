"""The product against the reference's own golden values (see tests/test_reference_golden.py for where
they come from and how their inputs are regenerated without R)."""
import numpy as np
import pytest

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
