"""SURVEY.md 8(f) rank 4, SPDK: importance_sample() and logLik(method = "spdk") -- the simulation smoother of
the approximating Gaussian model (ssm_ulg::simulate_states, src/model_ssm_ulg.cpp:141-217, with the stored
gains of fast_precomputing_smoother :359-413) and the importance weights of src/model_ssm_ung.cpp:199-211,
372-378.  One thread per antithetic pair of trajectories (bssm_b200/csrc/sim_smoother.cuh) must give the
oracle's bits: the oracle's restatement is pinned against dense numpy conditioning in tests/test_oracle.py."""
import numpy as np
import pytest

import bssm_b200
import data
import orc
from test_gpu_general_m import seasonal_model

pytestmark = pytest.mark.gpu


def _same(res, ref):
    alpha, w, ll, _ = ref
    got = np.transpose(res["alpha"], (1, 0, 2))
    assert np.array_equal(got.view(np.uint64), alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), w.view(np.uint64))
    assert res["logLik"] == ll[0] and res["approx_logLik"] == ll[1]


@pytest.mark.parametrize("nsim", [2, 3, 64, 257, 1000])
def test_importance_sample_svm(engine, nsim):
    model = data.svm_model(n=300, missing=(0, 17, 299))
    res = engine.importance_sample(model, nsim, seed=5, return_loglik=True)
    _same(res, orc.importance_sample(model, nsim, 5))


@pytest.mark.parametrize("nsim,antithetic", [(100, True), (101, True), (50, False), (51, False)])
def test_importance_sample_poisson_bsm(engine, nsim, antithetic):
    model = data.poisson_bsm_model(missing=(4, 50, 99))
    res = engine.importance_sample(model, nsim, use_antithetic=antithetic, seed=2, return_loglik=True)
    _same(res, orc.importance_sample(model, nsim, 2, use_antithetic=antithetic))
    if antithetic:  # the mirrored halves average to the smoothed mean of the approximating model
        h = nsim // 2
        a = res["alpha"]
        mid = 0.5 * (a[:, :, :h] + a[:, :, h:2 * h])
        assert np.abs(mid - mid[:, :, :1]).max() < 1e-12


@pytest.mark.parametrize("period,dist", [(4, "poisson"), (12, "negative binomial")])
def test_importance_sample_seasonal(engine, period, dist):
    """m = 5 and m = 13: the run-time-dimension instantiation"""
    model = seasonal_model(period, dist=dist)
    res = engine.importance_sample(model, 129, seed=8, return_loglik=True)
    _same(res, orc.importance_sample(model, 129, 8))


@pytest.mark.parametrize("dist", ["poisson", "binomial", "negative binomial", "gamma"])
def test_importance_sample_ar1_ng_with_regressors(engine, dist):
    rng = np.random.default_rng(4)
    n = 60
    x = rng.normal(size=n)
    u = np.full(n, 10.0) if dist == "binomial" else np.ones(n)
    lam = np.exp(0.4 + 0.3 * x)
    if dist == "binomial":
        y = rng.binomial(10, 1 / (1 + np.exp(-0.3 * x))).astype(float)
    elif dist == "gamma":
        y = rng.gamma(3.0, lam / 3.0)
    else:
        y = rng.poisson(lam).astype(float)
    y[9] = np.nan
    model = bssm_b200.ar1_ng(y, rho=0.8, sigma=0.3, mu=0.2, distribution=dist, phi=3.0, u=u, xreg=x[:, None],
                             beta=np.array([0.3]))
    res = engine.importance_sample(model, 200, seed=3, return_loglik=True)
    _same(res, orc.importance_sample(model, 200, 3))


def test_loglik_spdk(engine):
    model = data.poisson_bsm_model()
    ll = engine.logLik(model, 1000, method="spdk", seed=9)
    ref = orc.importance_sample(model, 1000, 9)[2]
    assert ll == ref[0]
    psi = engine.logLik(model, 1000, method="psi", seed=9)
    assert abs(ll - psi) < 0.05


def test_importance_sample_argument_checks(engine):
    model = data.poisson_bsm_model()
    with pytest.raises(ValueError):
        engine.importance_sample(model, 10, use_antithetic=1)
    with pytest.raises(ValueError):
        engine.importance_sample(model, 1)
    with pytest.raises(ValueError):
        engine.importance_sample(data.lg_model(), 10)


@pytest.mark.parametrize("m,tv,is_type", [(1, False, 2), (2, False, 1), (2, True, 3)])
def test_is_correction_spdk_parameter_sets(engine, m, tv, is_type):
    """approx_mcmc::is_correction_spdk (src/approx_mcmc.cpp:606-678): one model, one approximation at the stored
    mode and nsim (is1: nsim * count) simulated trajectories per stored draw, all draws in one launch"""
    from test_gpu_parset import draws, ung_update_model
    model = ung_update_model(m=m, tv=tv)
    mo = draws(model, 13)
    w, post = orc.correction_spdk_parset(model, mo, 30, is_type=is_type, seed=6)
    res = engine.post_correct(model, mo, 30, is_type="is%d" % is_type, seed=6, sampling_method="spdk")
    assert np.array_equal(res["weights"].view(np.uint64), w.view(np.uint64))
    assert np.array_equal(res["posterior"], post)
    assert len(set(res["weights"])) == len(w)  # every (expanded) draw has its own stream
    # same target as the psi-APF correction of the same draws (E_g[prod_t w_t]); these stored "modes" are random, so
    # the non-sequential sampler is heavy-tailed: only the typical draw is compared
    psi = engine.post_correct(model, mo, 2000, is_type="is2", seed=6)
    big = engine.post_correct(model, mo, 20000, is_type="is2", seed=6, sampling_method="spdk")
    assert np.median(np.abs(np.log(big["weights"]) - np.log(psi["weights"]))) < 0.4


def test_is_correction_spdk_with_regression_coefficients(engine):
    """beta in the parameter set: every draw has its own x beta (compute_xbeta, src/model_ssm_ung.h:83)"""
    rng = np.random.default_rng(8)
    n = 40
    x = rng.normal(size=(n, 2))
    y = rng.poisson(np.exp(0.5 + 0.2 * x[:, 0])).astype(float)
    y[11] = np.nan

    def update_fn(theta):
        return {"R": np.array([[0.1 * np.exp(theta[0])]]), "beta": np.array([0.2 + theta[1], theta[2]])}

    model = bssm_b200.ssm_ung(y, [1.0], [[0.95]], [[0.1]], [0.3], [[0.4]], "poisson", xreg=x, beta=np.array([0.2, 0.0]),
                              update_fn=update_fn)
    theta = 0.1 * rng.normal(size=(9, 3))
    mo = {"theta": theta, "counts": np.full(9, 2, dtype=np.uint32), "output_type": 3,
          "modes": np.asfortranarray(0.5 + 0.1 * rng.normal(size=(n, 9))), "posterior": np.zeros(9),
          "approx_loglik": -60.0 + rng.normal(size=9)}
    w, post = orc.correction_spdk_parset(model, mo, 21, is_type=2, seed=2)
    res = engine.post_correct(model, mo, 21, is_type="is2", seed=2, sampling_method="spdk")
    assert np.array_equal(res["weights"].view(np.uint64), w.view(np.uint64))
    assert np.array_equal(res["posterior"], post)


def test_is_correction_spdk_device_gains_equal_host_gains(engine, monkeypatch):
    """the per-draw approximation and smoother gains on the device (spdk_gains_kernel, one thread per draw) against
    the host loop over the draws (BSSM_B200_HOST_APPROX=1): same bits, m = 1, 2 and 5"""
    from test_gpu_parset import draws, ung_update_model
    for model in (ung_update_model(m=1), ung_update_model(m=2, tv=True)):
        mo = draws(model, 40)
        dev = engine.post_correct(model, mo, 16, is_type="is1", seed=3, sampling_method="spdk")
        monkeypatch.setenv("BSSM_B200_HOST_APPROX", "1")
        host = engine.post_correct(model, mo, 16, is_type="is1", seed=3, sampling_method="spdk")
        monkeypatch.delenv("BSSM_B200_HOST_APPROX")
        assert np.array_equal(dev["weights"].view(np.uint64), host["weights"].view(np.uint64))
    # m = 5 (run-time dimension): a seasonal model whose update_fn rescales the noise
    base = seasonal_model(4)
    R0 = np.array(base.R, dtype=float)

    def update_fn(theta):
        return {"R": R0 * np.exp(theta[0]), "P1": np.asarray(base.P1) * np.exp(theta[1])}

    model = bssm_b200.ssm_ung(base.y, base.Z, base.T, base.R, base.a1, base.P1, "poisson", u=base.u, update_fn=update_fn)
    rng = np.random.default_rng(2)
    ns = 10
    mo = {"theta": 0.1 * rng.normal(size=(ns, 2)), "counts": np.ones(ns, dtype=np.uint32), "output_type": 3,
          "modes": np.asfortranarray(1.0 + 0.1 * rng.normal(size=(model.n, ns))), "posterior": np.zeros(ns)}
    dev = engine.post_correct(model, mo, 20, is_type="is2", seed=4, sampling_method="spdk")
    w, _ = orc.correction_spdk_parset(model, mo, 20, is_type=2, seed=4)
    assert np.array_equal(dev["weights"].view(np.uint64), w.view(np.uint64))
