"""ssm_nlg (growth model, BASELINE config 3) and ssm_sde (GBM Milstein, config 5) bootstrap
filters / smoothers against the oracle's restatements of ssm_nlg::bsf_filter
(src/model_ssm_nlg.cpp:1066-1146) and ssm_sde::bsf_filter (src/model_ssm_sde.cpp:21-82)."""
import numpy as np
import pytest

import data
import orc

pytestmark = pytest.mark.gpu


def _cube(res):
    return np.transpose(res["alpha"], (1, 0, 2))


@pytest.mark.parametrize("N", [100, 1000, 5000])
def test_growth_bsf_parity(engine, N):
    model = data.growth_model(60)
    ref = orc.nlg_bsf_filter(model, N, seed=1)
    res = engine.bootstrap_filter(model, N, seed=1, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    at, att, Pt, Ptt = orc.filter_summary(ref.alpha, ref.weights)
    np.testing.assert_allclose(res["att"].T, att, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res["Ptt"], Ptt, rtol=1e-8, atol=1e-12)


def test_growth_smoother_parity_and_missing(engine):
    model = data.growth_model(50, missing=(0, 10, 49))
    ref = orc.nlg_bsf_filter(model, 2000, seed=3)
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    mean, cov = orc.summary(sm)
    res = engine.particle_smoother(model, 2000, method="bsf", seed=3, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices) and res["logLik"] == ref.loglik
    assert np.array_equal(_cube(res).view(np.uint64), sm.view(np.uint64))
    np.testing.assert_allclose(res["alphahat"].T, mean, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res["Vt"], cov, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("N,L", [(100, 1), (512, 3), (3000, 4)])
def test_gbm_bsf_parity(engine, N, L):
    model = data.gbm_model(40)
    ref = orc.sde_bsf_filter(model, N, L, seed=1)
    res = engine.bootstrap_filter(model, N, seed=1, return_indices=True, L=L)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert res["logLik"] == ref.loglik
    assert res["att"].shape == (model.n + 1, 1) and res["att"][-1, 0] == 0.0  # src/R_sde.cpp:56-60
    assert engine.logLik(model, N, seed=1, L=L) == ref.loglik  # loglik_sde path


def test_gbm_smoother_and_missing(engine):
    model = data.gbm_model(30, missing=(5,))
    ref = orc.sde_bsf_filter(model, 800, 2, seed=2)
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    res = engine.particle_smoother(model, 800, method="bsf", seed=2, L=2)
    assert np.array_equal(_cube(res).view(np.uint64), sm.view(np.uint64))
    assert res["logLik"] == ref.loglik


def test_sde_argument_checks(engine):
    model = data.gbm_model(10)
    with pytest.raises(ValueError, match="Discretization level"):
        engine.bootstrap_filter(model, 100)
    with pytest.raises(ValueError):
        engine.ssm_nlg(np.zeros(5), "nope", [0, 0, 0], [0] * 6)


def test_growth_loglik_close_to_large_n_reference(engine):
    """statistical anchor: N = 2^17 (cluster launch) agrees with an independent N = 4000 oracle run"""
    model = data.growth_model(40)
    ll_big = engine.logLik(model, 131072, method="bsf", seed=1)
    ll_ref = np.mean([orc.nlg_bsf_filter(model, 4000, seed=s).loglik for s in (1, 2, 3)])
    assert abs(ll_big - ll_ref) < 0.5


@pytest.mark.parametrize("N", [100, 3000])
def test_mng_bsf_parity(engine, N):
    """ssm_mng: three series (Poisson, negative binomial, Gaussian), 2-dim state, missing values"""
    model = data.mng_model()
    ref = orc.mng_bsf_filter(model, N, seed=1)
    res = engine.bootstrap_filter(model, N, seed=1, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    # the reference adds the constants term by term, the kernel adds their sum: few-ulp difference
    assert res["logLik"] == pytest.approx(ref.loglik, rel=1e-13)
    assert np.all(res["weights"][:, 7] == 1.0)  # all series missing at t = 7
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    res2 = engine.particle_smoother(model, N, method="bsf", seed=1)
    assert np.array_equal(_cube(res2).view(np.uint64), sm.view(np.uint64))


@pytest.mark.parametrize("N", [100, 2000])
def test_growth_psi_apf_parity(engine, N):
    """psi-APF for ssm_nlg (growth model): EKF-started Gaussian approximation on the host,
    particle loop with the transition-density ratio in the weights on the device"""
    model = data.growth_model(80)
    ref = orc.nlg_psi_filter(model, N, seed=1)
    res = engine.particle_smoother(model, N, method="psi", seed=1, return_indices=True)
    assert res["approx_logLik"] == ref.approx_loglik
    assert np.array_equal(res["indices"], ref.indices)
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), sm.view(np.uint64))
    job = engine.FilterJob(model, N, method="psi", seed=1, out_flags=2)
    job.run()
    b = job.fetch(0, weights=True)
    assert np.array_equal(b.weights.view(np.uint64), ref.weights.view(np.uint64))
    job.close()


def test_growth_psi_missing_and_statistics(engine):
    model = data.growth_model(60, missing=(0, 7, 8, 59))
    ref = orc.nlg_psi_filter(model, 500, seed=4)
    res = engine.particle_smoother(model, 500, method="psi", seed=4, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices) and res["logLik"] == ref.loglik
    # psi-APF and BSF estimate the same likelihood (vignettes/psi_pf.Rmd table 1: psi has far lower variance)
    full = data.growth_model(60)
    lp = np.array([engine.logLik(full, 200, method="psi", seed=s) for s in range(1, 7)])
    lb = np.array([engine.logLik(full, 20000, method="bsf", seed=s) for s in range(1, 4)])
    assert lp.std() < 0.1 and abs(lp.mean() - lb.mean()) < 0.3
