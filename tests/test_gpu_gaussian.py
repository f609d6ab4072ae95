"""Linear-Gaussian models (ssm_ulg / bsm_lg): the `gaussian = TRUE` branch of bsf / bsf_smoother and
gaussian_psi_smoother.  Besides oracle parity these have an EXACT reference: the Kalman smoother
(tests/testthat/test_particle_smoother.R:25-45 compares the particle smoothers with it at tol 1e-2)."""
import numpy as np
import pytest

import data
import orc
from test_oracle import _rts_smoother

pytestmark = pytest.mark.gpu


def _cube(res):
    return np.transpose(res["alpha"], (1, 0, 2))


@pytest.mark.parametrize("N", [100, 3000])
def test_ulg_bsf_parity(engine, N):
    model = data.lg_model(60, missing=(5, 6))
    ref = orc.ulg_bsf_filter(model, N, seed=1)
    res = engine.bootstrap_filter(model, N, seed=1, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik


def test_bsm_lg_equals_equivalent_ssm_ulg(engine):
    """tests/testthat/test_bootstrap_filter.R:5-17: bsm_lg and the same model written as ssm_ulg
    give identical bootstrap-filter output"""
    a = data.lg_model(50)
    b = engine.ssm_ulg(a.y, a.Z, a.H, a.T, a.R, a.a1, a.P1)
    ra, rb = engine.bootstrap_filter(a, 200, seed=1), engine.bootstrap_filter(b, 200, seed=1)
    assert ra["logLik"] == rb["logLik"] and np.array_equal(ra["alpha"], rb["alpha"])
    assert np.array_equal(ra["att"], rb["att"])


def test_gaussian_psi_smoother_parity_and_kalman(engine):
    model = data.lg_model(60)
    ref = orc.ulg_psi_filter(model, 512, seed=3)
    res = engine.gaussian_psi_smoother(model, 512, seed=3)
    assert np.array_equal(_cube(res).view(np.uint64), ref.view(np.uint64))
    # exact answer: Kalman / RTS smoother of the same model (numpy)
    Z, T = model.Z[:, 0], model.T[:, :, 0]
    RR = model.R[:, :, 0] @ model.R[:, :, 0].T
    asm, Psm, *_ = _rts_smoother(model.y, Z, np.full(model.n, model.H[0] ** 2), T, RR, model.a1, model.P1)
    big = engine.gaussian_psi_smoother(model, 100000, seed=1)
    np.testing.assert_allclose(big["alphahat"], asm, atol=2e-2)
    np.testing.assert_allclose(np.moveaxis(big["Vt"], 2, 0), Psm, atol=2e-2)


def test_bsf_smoother_agrees_with_kalman_smoother(engine):
    """test_particle_smoother.R:25-31: BSF smoother at N = 1e5 vs the exact smoother, tol 1e-2 (relative)"""
    model = data.lg_model(40)
    Z, T = model.Z[:, 0], model.T[:, :, 0]
    RR = model.R[:, :, 0] @ model.R[:, :, 0].T
    asm, Psm, af, Pf, ap, Pp = _rts_smoother(model.y, Z, np.full(model.n, model.H[0] ** 2), T, RR, model.a1, model.P1)
    res = engine.particle_smoother(model, 100000, method="bsf", seed=1)
    np.testing.assert_allclose(res["alphahat"][:, 0], asm[:, 0], atol=0.05)
    flt = engine.bootstrap_filter(model, 100000, seed=2)
    np.testing.assert_allclose(flt["att"][:, 0], af[:, 0], atol=0.03)   # filtered means
    np.testing.assert_allclose(flt["at"][1:, 0], ap[1:, 0], atol=0.03)  # one-step predictions
    # exact log-likelihood of the linear-Gaussian model
    ll = 0.0
    for t in range(model.n):
        F = Z @ Pp[t] @ Z + model.H[0] ** 2
        v = model.y[t] - Z @ ap[t]
        ll -= 0.5 * (np.log(2 * np.pi) + np.log(F) + v * v / F)
    assert abs(flt["logLik"] - ll) < 0.05
