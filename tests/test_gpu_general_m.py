"""General state dimension: bsm_ng with a seasonal component has m = 2 + period - 1
states (src/model_bsm_ng.cpp:17-51, R/models.R:926-1114) -- 5 for quarterly, 13 for monthly data.  The
run-time-m policies (bssm_b200/csrc/models_ung_dyn.cuh, m <= 16, k <= m) must give the oracle's bits:
bootstrap filter, psi-APF, and the particle smoother built on either."""
import numpy as np
import pytest

import bssm_b200
import data
import orc

pytestmark = pytest.mark.gpu


def seasonal_model(period, n=72, dist="poisson", seed=3):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    season = 0.6 * np.sin(2 * np.pi * t / period)
    level = 1.0 + np.cumsum(rng.normal(0, 0.05, n))
    lam = np.exp(level + season)
    y = rng.poisson(lam).astype(float) if dist == "poisson" else rng.negative_binomial(5, 5 / (5 + lam)).astype(float)
    y[7] = np.nan  # a missing observation
    return bssm_b200.bsm_ng(y, sd_level=0.1, sd_slope=0.01, sd_seasonal=0.05, distribution=dist, phi=5.0,
                            period=period, P1=np.diag(np.full(2 + period - 1, 1.0)))


@pytest.mark.parametrize("period,N", [(4, 100), (4, 3000), (12, 100), (12, 1000)])
def test_bootstrap_filter_general_m(engine, period, N):
    model = seasonal_model(period)
    assert model.m == 2 + period - 1 and model.k == 3
    ref = orc.bsf_filter(model, N, seed=2)
    res = engine.bootstrap_filter(model, N, seed=2, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    assert np.array_equal(np.transpose(res["alpha"], (1, 0, 2)).view(np.uint64), ref.alpha.view(np.uint64))
    at, att, Pt, Ptt = orc.filter_summary(ref.alpha, ref.weights)
    np.testing.assert_allclose(res["att"], att.T, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res["Ptt"], Ptt, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res["at"], at.T, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("period,N,dist", [(4, 200, "poisson"), (12, 200, "poisson"), (12, 2500, "negative binomial")])
def test_psi_apf_and_smoother_general_m(engine, period, N, dist):
    model = seasonal_model(period, dist=dist)
    ref = orc.psi_filter_m(model, N, seed=4)
    res = engine.particle_smoother(model, N, method="psi", seed=4, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    assert np.array_equal(np.transpose(res["alpha"], (1, 0, 2)).view(np.uint64), sm.view(np.uint64))
    mean, cov = orc.summary(sm)
    np.testing.assert_allclose(res["alphahat"], mean.T, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res["Vt"], cov, rtol=1e-8, atol=1e-11)


def test_bsf_smoother_general_m_and_unusual_k(engine):
    """m = 4 with k = 2 (no specialised instantiation of (m, k)): the run-time policy takes over"""
    rng = np.random.default_rng(9)
    y = rng.poisson(3.0, 50).astype(float)
    model = bssm_b200.bsm_ng(y, sd_level=0.1, sd_seasonal=0.05, distribution="poisson", period=4,
                             P1=np.diag(np.full(4, 1.0)))
    assert (model.m, model.k) == (4, 2)
    ref = orc.bsf_filter(model, 500, seed=6)
    res = engine.particle_smoother(model, 500, method="bsf", seed=6, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices) and res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    assert np.array_equal(np.transpose(res["alpha"], (1, 0, 2)).view(np.uint64), sm.view(np.uint64))


@pytest.mark.parametrize("period", [4, 12])
def test_post_correction_with_a_seasonal_component(engine, period):
    """is_correction_psi for bsm_ng with a seasonal component (m = 5 / 13): bsm_ng::update_model per draw
    (src/model_bsm_ng.cpp:17-51), the run-time-m psi-APF per draw, trajectories and the weighted summary
    merge (src/approx_mcmc.cpp:394-424) at run-time m"""
    import data
    model = seasonal_model(period, n=48)
    mo = data.bsm_mcmc_output(model, 10, jitter=0.05)
    ref = orc.post_correct(model, mo, 64, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.post_correct(model, mo1, 32, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 32, is_type="is3", seed=5)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.post_correct(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-11)


def mng_structural_model(p, n=48, seed=5, season=True):
    """p Poisson / negative binomial series, each with its own level, one common slope and (optionally) one
    common quarterly seasonal: m = p + 1 (+ 3), the multivariate structural models of ssm_mng"""
    rng = np.random.default_rng(seed)
    m = p + 1 + (3 if season else 0)
    T = np.eye(m)
    T[:p, p] = 1.0  # every level moves with the common slope
    if season:
        T[p + 1:, p + 1:] = 0.0
        T[p + 1, p + 1:] = -1.0
        T[p + 2, p + 1] = 1.0
        T[p + 3, p + 2] = 1.0
    k = p + 1 + (1 if season else 0)
    R = np.zeros((m, k))
    for j in range(p):
        R[j, j] = 0.08
    R[p, p] = 0.01
    if season:
        R[p + 1, p + 1] = 0.03
    Z = np.zeros((p, m))
    Z[np.arange(p), np.arange(p)] = 1.0
    if season:
        Z[:, p + 1] = 1.0
    a = np.zeros((n, m))
    a[0, :p] = rng.normal(1.0, 0.3, size=p)
    if season:
        a[0, p + 1:] = [0.3, -0.1, -0.3]
    for t in range(1, n):
        a[t] = T @ a[t - 1] + R @ rng.normal(size=k)
    sig = a @ Z.T
    y = rng.poisson(np.exp(sig)).astype(float)
    y[5, 0] = np.nan
    y[9, :] = np.nan
    dist = ["poisson" if j % 2 == 0 else "negative binomial" for j in range(p)]
    P1 = np.diag(np.r_[np.full(p, 1.0), 0.01, np.full(m - p - 1, 0.25)])
    a1 = np.r_[np.full(p, 1.0), np.zeros(m - p)]
    return bssm_b200.ssm_mng(y, Z, T, R, a1=a1, P1=P1, distribution=dist, phi=np.full(p, 4.0))


@pytest.mark.parametrize("p,season,N", [(2, True, 100), (2, True, 3000), (4, False, 500), (10, False, 300)])
def test_mng_general_m(engine, p, season, N):
    """ssm_mng with m > 3 and p up to 16 (m = 6 / 5 / 11): bootstrap filter and psi-APF at run-time m,
    bit-exact against the oracle (src/model_ssm_mng.cpp:461-668)"""
    model = mng_structural_model(p, season=season)
    assert model.m > 3 and model.p == p
    ref = orc.mng_bsf_filter(model, N, seed=1)
    res = engine.bootstrap_filter(model, N, seed=1, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(np.transpose(res["alpha"], (1, 0, 2)).view(np.uint64), ref.alpha.view(np.uint64))
    assert res["logLik"] == pytest.approx(ref.loglik, rel=1e-13)  # constants summed once on the host
    at, att, Pt, Ptt = orc.filter_summary(ref.alpha, ref.weights)
    np.testing.assert_allclose(res["att"], att.T, rtol=1e-10, atol=1e-12)
    refp = orc.mng_psi_filter(model, N, seed=3)
    resp = engine.particle_smoother(model, N, method="psi", seed=3, return_indices=True)
    assert resp["approx_logLik"] == refp.approx_loglik
    assert np.array_equal(resp["indices"], refp.indices)
    assert resp["logLik"] == refp.loglik
    sm = orc.filter_smoother(refp.alpha, refp.indices)
    assert np.array_equal(np.transpose(resp["alpha"], (1, 0, 2)).view(np.uint64), sm.view(np.uint64))
    ah, Vt = orc.summary(sm)
    np.testing.assert_allclose(resp["alphahat"], ah.T, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(resp["Vt"], Vt, rtol=1e-8, atol=1e-12)
