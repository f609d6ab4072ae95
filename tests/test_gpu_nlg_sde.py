"""ssm_nlg (growth model, BASELINE config 3) and ssm_sde (GBM Milstein, config 5) bootstrap
filters / smoothers against the oracle's restatements of ssm_nlg::bsf_filter
(src/model_ssm_nlg.cpp:1066-1146) and ssm_sde::bsf_filter (src/model_ssm_sde.cpp:21-82)."""
import numpy as np
import pytest

import data
import orc
import bssm_b200 as engine_mod

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


@pytest.mark.parametrize("N,L", [(200, 2), (3000, 3)])
def test_poisson_ou_bsf_parity(engine, N, L):
    """second functor of the ssm_sde registry: OU state, Poisson observations (sde_poisson_ou)"""
    model = data.poisson_ou_model(40, missing=(6,))
    ref = orc.sde_bsf_filter(model, N, L, seed=4)
    res = engine.bootstrap_filter(model, N, seed=4, return_indices=True, L=L)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    assert np.all(res["weights"][:, 6] == 1.0)
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    res2 = engine.particle_smoother(model, N, method="bsf", seed=4, L=L)
    assert np.array_equal(_cube(res2).view(np.uint64), sm.view(np.uint64))


@pytest.mark.parametrize("N", [100, 3000])
def test_ar_exp_bsf_parity(engine, N):
    """second functor of the ssm_nlg registry: AR(1) state observed through exp(alpha) (nlg_ar_exp), m = 1"""
    model = data.ar_exp_model(70, missing=(4, 5))
    ref = orc.nlg_bsf_filter(model, N, seed=2)
    res = engine.bootstrap_filter(model, N, seed=2, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    res2 = engine.particle_smoother(model, N, method="bsf", seed=2)
    assert np.array_equal(_cube(res2).view(np.uint64), sm.view(np.uint64))


@pytest.mark.parametrize("N,iekf", [(100, 0), (2000, 0), (500, 5)])
def test_ar_exp_psi_apf_parity(engine, N, iekf):
    """psi-APF for the one-state registry model: dmvnorm takes its scalar branch for the transition
    densities (src/dmvnorm.cpp:9-16); iekf_iter > 0 iterates the EKF start (src/model_ssm_nlg.cpp:325-356)"""
    model = data.ar_exp_model(70, missing=(4, 5), iekf_iter=iekf)
    ref = orc.nlg_psi_filter(model, N, seed=3, iekf_iter=iekf)
    res = engine.particle_smoother(model, N, method="psi", seed=3, return_indices=True)
    assert res["approx_logLik"] == ref.approx_loglik
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), sm.view(np.uint64))
    if iekf == 0:
        assert engine.logLik(model, N, method="psi", seed=3) == ref.loglik
        # psi-APF and bootstrap filter estimate the same likelihood
        lb = engine.logLik(model, 50000, method="bsf", seed=1)
        assert abs(res["logLik"] - lb) < 1.0


@pytest.mark.parametrize("N", [100, 2500])
def test_sin_exp_parity(engine, N):
    """third functor of the ssm_nlg registry (nlg_sin_exp): a genuinely nonlinear transition, so the
    psi-APF weights carry a non-trivial ratio of true and linearised transition densities"""
    model = data.sin_exp_model(60, missing=(7,))
    ref = orc.nlg_bsf_filter(model, N, seed=2)
    res = engine.bootstrap_filter(model, N, seed=2, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert res["logLik"] == ref.loglik
    refp = orc.nlg_psi_filter(model, N, seed=3)
    resp = engine.particle_smoother(model, N, method="psi", seed=3, return_indices=True)
    assert resp["approx_logLik"] == refp.approx_loglik
    assert np.array_equal(resp["indices"], refp.indices)
    assert np.array_equal(resp["weights"].view(np.uint64), refp.weights.view(np.uint64))
    assert resp["logLik"] == refp.loglik
    sm = orc.filter_smoother(refp.alpha, refp.indices)
    assert np.array_equal(_cube(resp).view(np.uint64), sm.view(np.uint64))
    lb = engine.logLik(model, 50000, method="bsf", seed=1)
    lp = np.mean([engine.logLik(model, 2000, method="psi", seed=s) for s in range(1, 6)])
    assert abs(lp - lb) < 0.5


@pytest.mark.parametrize("name,N", [("growth", 100), ("growth", 3000), ("ar_exp", 1000), ("sin_exp", 1000)])
def test_ekpf_parity(engine, name, N):
    """EKF-based particle filter (ssm_nlg::ekf_filter, src/model_ssm_nlg.cpp:1150-1260): every particle
    proposes from its own EKF update; 2 x 2 linear algebra per particle for the growth model"""
    model = {"growth": lambda: data.growth_model(60, missing=(4,)),
             "ar_exp": lambda: data.ar_exp_model(70, missing=(4, 5)),
             "sin_exp": lambda: data.sin_exp_model(60, missing=(7,))}[name]()
    ref = orc.nlg_ekf_filter(model, N, seed=5)
    res = engine.ekpf_filter(model, N, seed=5, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert np.array_equal(_cube(res).view(np.uint64), ref.alpha.view(np.uint64))
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    res2 = engine.particle_smoother(model, N, method="ekf", seed=5)
    assert np.array_equal(_cube(res2).view(np.uint64), sm.view(np.uint64))
    # the three filters estimate the same likelihood
    lb = engine.logLik(model, 50000, method="bsf", seed=1)
    le = np.mean([engine.logLik(model, 5000, method="ekf", seed=s) for s in range(1, 5)])
    assert abs(le - lb) < 0.5


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


def _mng_variant(kind):
    if kind == "mixed_missing":      # Gaussian series with NAs: the reference stops its sweeps after one
        return data.mng_model(60)
    if kind == "mixed":
        m = data.mng_model(60, missing=False)
        m.y[3, 0] = np.nan
        m.y[11, :2] = np.nan
        return m
    rng = np.random.default_rng(11)
    if kind == "m1p2":               # one state, Poisson + binomial
        n = 50
        a = np.cumsum(0.15 * rng.normal(size=n))
        y = np.column_stack([rng.poisson(np.exp(a + 1.0)), rng.binomial(10, 1 / (1 + np.exp(-a)))]).astype(float)
        y[6, 1] = np.nan
        u = np.column_stack([np.ones(n), np.full(n, 10.0)])
        return engine_mod.ssm_mng(y, np.array([[1.0], [1.0]]), [[1.0]], [[0.15]], a1=[0.0], P1=[[1.0]],
                                  distribution=["poisson", "binomial"], u=u, D=[1.0, 0.0])
    # m = 3, p = 3 with a gamma series and time-varying D
    n = 40
    a = np.zeros((n, 3))
    for t in range(1, n):
        a[t] = [0.9 * a[t - 1, 0] + 0.2 * rng.normal(), a[t - 1, 1] + a[t - 1, 2] + 0.05 * rng.normal(),
                0.8 * a[t - 1, 2] + 0.02 * rng.normal()]
    Z = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.5, 0.5, 1.0]])
    D = np.vstack([np.full(n, 0.5), np.linspace(0, 0.3, n), np.zeros(n)])
    sig = a @ Z.T + D.T
    y = np.column_stack([rng.poisson(np.exp(sig[:, 0])), rng.gamma(2.0, np.exp(sig[:, 1]) / 2.0),
                         rng.poisson(3 * np.exp(sig[:, 2]))]).astype(float)
    y[0, 0] = np.nan
    u = np.column_stack([np.ones(n), np.ones(n), np.full(n, 3.0)])
    T = np.array([[0.9, 0, 0], [0, 1, 1], [0, 0, 0.8]])
    return engine_mod.ssm_mng(y, Z, T, np.diag([0.2, 0.05, 0.02]), a1=[0, 0, 0], P1=np.diag([0.3, 1.0, 0.1]),
                              distribution=["poisson", "gamma", "poisson"], phi=[1.0, 2.0, 1.0], u=u, D=D)


@pytest.mark.parametrize("kind,N", [("mixed", 100), ("mixed", 3000), ("mixed_missing", 500), ("m1p2", 1000),
                                    ("m3p3", 700)])
def test_mng_psi_apf_parity(engine, kind, N):
    """psi-APF for ssm_mng (src/model_ssm_mng.cpp:461-559): multivariate Gaussian approximation on
    the host, particle loop on the device; bit-exact against the oracle"""
    model = _mng_variant(kind)
    ref = orc.mng_psi_filter(model, N, seed=3)
    res = engine.particle_smoother(model, N, method="psi", seed=3, return_indices=True)
    assert res["approx_logLik"] == ref.approx_loglik
    assert np.array_equal(res["indices"], ref.indices)
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    assert res["logLik"] == ref.loglik
    sm = orc.filter_smoother(ref.alpha, ref.indices)
    assert np.array_equal(_cube(res).view(np.uint64), sm.view(np.uint64))
    ah, Vt = orc.summary(sm)
    np.testing.assert_allclose(res["alphahat"], ah.T, rtol=1e-12, atol=1e-14)
    assert engine.logLik(model, N, method="psi", seed=3) == ref.loglik


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


@pytest.mark.parametrize("case", ["growth_bsf", "growth_ekf", "growth_psi", "ar_exp_ekf", "ar_exp_psi", "sin_exp_bsf",
                                  "sin_exp_psi", "gbm", "poisson_ou"])
def test_job_set_theta_nlg_sde_matches_fresh_job(engine, case):
    """the filter inside nonlinear_pm_mcmc / sde_pm_mcmc (one bootstrap, EKF or psi particle filter per
    iteration at a new theta): re-parameterising a resident job equals building the job at that theta, bit
    for bit (psi-APF jobs redo approximate() on the host and refill the same device buffers)"""
    import bssm_b200
    method = "bsf"
    if case in ("gbm", "poisson_ou"):
        model = data.gbm_model(30) if case == "gbm" else data.poisson_ou_model(30)
        make = lambda th: bssm_b200.ssm_sde(model.y, case, th, model.x0, model.positive)
        kw = {"L": 3}
        thetas = [np.abs(np.asarray(model.theta) * s) for s in (1.1, 0.85)]
    else:
        kind, method = case.rsplit("_", 1)
        model = {"growth": data.growth_model, "ar_exp": data.ar_exp_model, "sin_exp": data.sin_exp_model}[kind](40)
        known = getattr(model, "known_params", ())
        make = lambda th: bssm_b200.ssm_nlg(model.y, kind, theta=th, known_params=known)
        kw = {}
        thetas = [np.asarray(model.theta) + d for d in (0.07, -0.05)]
        if kind == "ar_exp":
            thetas = [np.asarray(model.theta) + np.array([d, -d / 4, d, d]) for d in (0.07, -0.05)]
    job = engine.FilterJob(model, 300, method=method, seed=4, n_filters=3, **kw)
    job.run()
    ll0 = job.logliks().copy()
    for th in thetas:
        job.set_theta(th)
        job.run()
        fresh = engine.FilterJob(make(th), 300, method=method, seed=4, n_filters=3, **kw)
        fresh.run()
        assert np.array_equal(job.logliks(), fresh.logliks())
        assert not np.array_equal(job.logliks(), ll0)
        fresh.close()
    job.close()


def _kalman_loglik_random_walk(y, H):
    """exact log-likelihood of alpha_{t+1} = alpha_t + eta, y_t = alpha_t + H eps, a1 = 0, P1 = 1 (numpy, independent)"""
    a, P, ll = 0.0, 1.0, 0.0
    for yt in y:
        if np.isfinite(yt):
            F = P + H * H
            v = yt - a
            ll += -0.5 * (np.log(2 * np.pi) + np.log(F) + v * v / F)
            K = P / F
            a, P = a + K * v, P - K * P
        P = P + 1.0
    return ll


def test_linear_gaussian_functor(engine):
    """the sixth shipped example model, nlg_linear_gaussian (R/cpp_example_models.R:512-634): bootstrap filter,
    psi-APF and EKF particle filter bit-exact against the oracle; and pinned INDEPENDENTLY of the oracle: for a
    linear-Gaussian model the psi-APF proposes from the exact smoothing distribution, so its log-likelihood
    estimate IS the Kalman log-likelihood (up to rounding) whatever the number of particles"""
    model = data.linear_gaussian_model(70, missing=(11, 12))
    N = 600
    ref = orc.nlg_bsf_filter(model, N, seed=3)
    res = engine.bootstrap_filter(model, N, seed=3, return_indices=True)
    assert np.array_equal(res["indices"], ref.indices) and res["logLik"] == ref.loglik
    assert np.array_equal(res["weights"].view(np.uint64), ref.weights.view(np.uint64))
    refp = orc.nlg_psi_filter(model, N, seed=4)
    resp = engine.particle_smoother(model, N, method="psi", seed=4, return_indices=True)
    assert np.array_equal(resp["indices"], refp.indices) and resp["logLik"] == refp.loglik
    exact = _kalman_loglik_random_walk(model.y, np.exp(model.theta[0]))
    assert abs(resp["logLik"] - exact) < 1e-8 * abs(exact)
    assert abs(res["logLik"] - exact) < 1.0  # bootstrap filter: Monte Carlo error only
    refe = orc.nlg_ekf_filter(model, N, seed=5)
    rese = engine.FilterJob(model, N, method="ekf", seed=5, out_flags=engine_mod.api.OUT_INDICES)
    rese.run()
    be = rese.fetch(0, indices=True)
    assert np.array_equal(be.indices, refe.indices) and rese.logliks()[0] == refe.loglik
    rese.close()
    # for a linear model the EKF proposal is the optimal proposal: its estimate is close to exact as well
    assert abs(refe.loglik - exact) < 0.5
