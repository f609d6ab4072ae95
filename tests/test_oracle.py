"""Pins the CPU oracle: hand-derived known answers for stratified_sample
(SURVEY.md 8c(3)), agreement of its two arithmetic modes, invariants the
reference's own tests assert (tests/testthat/test_bootstrap_filter.R:21-107,
test_particle_smoother.R:109-125), and small independent numpy restatements of
filter_smoother / summary / filter_summary."""
import numpy as np
import pytest

import bssm_b200
import data
import orc


def test_stratified_sample_hand_vectors():
    # cumsum (.1,.3,.6,1), thresholds (.125,.375,.625,.875) -> (1,2,3,3)
    assert list(orc.stratified_sample([.1, .2, .3, .4], [.5] * 4)) == [1, 2, 3, 3]
    assert list(orc.stratified_sample([0, 0, 1], [.5] * 3)) == [2, 2, 2]
    # p[last] is forced to 1 even if the cumsum rounds below 1
    p = np.full(10, 0.1)
    assert list(orc.stratified_sample(p * (1 - 1e-12), [0.999999999999] * 10))[-1] == 9
    # uniform weights: r = .5 gives the identity; r = 0 hits the `<=` (threshold j/N equals p_{j-1})
    assert list(orc.stratified_sample(np.full(8, 1 / 8), np.full(8, .5))) == list(range(8))
    assert list(orc.stratified_sample(np.full(8, 1 / 8), np.zeros(8))) == [0, 0, 1, 2, 3, 4, 5, 6]
    # the repro (fixed-point) cumsum gives the same answers
    assert list(orc.stratified_sample([.1, .2, .3, .4], [.5] * 4, repro=True)) == [1, 2, 3, 3]
    assert list(orc.stratified_sample([0, 0, 1], [.5] * 3, repro=True)) == [2, 2, 2]


def test_stratified_sample_matches_numpy_definition():
    rng = np.random.default_rng(0)
    for N in (5, 64, 1000):
        w = rng.exponential(size=N)
        w /= w.sum()
        r = rng.uniform(size=N)
        got = orc.stratified_sample(w, r)
        p = np.cumsum(w)
        p[-1] = 1
        u = (r + np.arange(N)) * (1.0 / N)
        want = np.searchsorted(p, u, side="left")
        assert np.array_equal(got, want)
        assert np.all(np.diff(got.astype(np.int64)) >= 0)


@pytest.mark.parametrize("N", [10, 100, 1000])
def test_arith_modes_agree_svm_psi(N):
    """reference-verbatim arithmetic (sequential cumsum, arma accu) and the
    order-independent fixed-point arithmetic give the same ancestors"""
    model = data.svm_model(60)
    a = orc.psi_filter(model, N, seed=1, arith=orc.ARITH_REFERENCE)
    b = orc.psi_filter(model, N, seed=1, arith=orc.ARITH_REPRO)
    assert np.array_equal(a.indices, b.indices)
    assert abs(a.loglik - b.loglik) <= 1e-10 * abs(a.loglik)
    np.testing.assert_allclose(a.weights, b.weights, rtol=0, atol=0)


def test_arith_modes_agree_poisson_bsf():
    model = data.poisson_bsm_model()
    a = orc.bsf_filter(model, 100, seed=1, arith=orc.ARITH_REFERENCE)
    b = orc.bsf_filter(model, 100, seed=1, arith=orc.ARITH_REPRO)
    assert np.array_equal(a.indices, b.indices)
    assert abs(a.loglik - b.loglik) <= 1e-10 * abs(a.loglik)


def test_bsf_invariants_poisson_series():
    """tests/testthat/test_bootstrap_filter.R:40-60: finite, weights in [0, 1]"""
    model = data.poisson_bsm_model()
    r = orc.bsf_filter(model, 100, seed=1)
    assert np.isfinite(r.loglik) and r.stopped_at == -1
    assert np.all(r.weights >= 0) and np.all(r.weights <= 1) and np.all(np.isfinite(r.alpha))
    assert np.all(r.weights.max(axis=0) == 1.0)
    assert np.all(r.weights[:, -1] == 1.0)  # column n: pure prediction
    assert np.all(np.diff(r.indices.astype(np.int64), axis=0) >= 0)
    # same-seed determinism (test_mcmc.R:332-349), different seed differs
    r2 = orc.bsf_filter(model, 100, seed=1)
    assert r2.loglik == r.loglik and np.array_equal(r2.indices, r.indices)
    assert orc.bsf_filter(model, 100, seed=2).loglik != r.loglik


def test_missing_observations():
    model = data.poisson_bsm_model(missing=(0, 5, 6, 99))
    r = orc.bsf_filter(model, 50, seed=3)
    assert np.isfinite(r.loglik)
    for t in (0, 5, 6, 99):
        assert np.all(r.weights[:, t] == 1.0)
    ms = data.svm_model(40, missing=(3, 4))
    rp = orc.psi_filter(ms, 50, seed=3)
    assert np.isfinite(rp.loglik) and np.all(rp.weights[:, 3] == 1.0)


def test_psi_and_bsf_agree_statistically_svm():
    """test_particle_smoother.R:109-125: psi(N=100) vs BSF(N=1e4) on svm"""
    model = data.svm_model(100)
    lp = np.array([orc.psi_filter(model, 100, seed=s).loglik for s in range(1, 9)])
    lb = orc.bsf_filter(model, 10000, seed=1).loglik
    assert lp.std() < 0.1
    assert abs(lp.mean() - lb) < 0.5


def test_psi_loglik_poisson_ar1():
    import bssm_b200
    rng = np.random.default_rng(5)
    a = np.zeros(80)
    for t in range(1, 80):
        a[t] = 0.9 * a[t - 1] + 0.3 * rng.normal()
    y = rng.poisson(np.exp(a + 1.0)).astype(float)
    model = bssm_b200.ar1_ng(y, rho=0.9, sigma=0.3, mu=1.0, distribution="poisson")
    lp = np.array([orc.psi_filter(model, 200, seed=s).loglik for s in range(1, 6)])
    lb = np.array([orc.bsf_filter(model, 4000, seed=s).loglik for s in range(1, 4)])
    assert lp.std() < 0.05
    assert abs(lp.mean() - lb.mean()) < 0.3


def test_filter_smoother_and_summaries_vs_numpy():
    model = data.poisson_bsm_model(30)
    r = orc.bsf_filter(model, 64, seed=2)
    m, n1, N = r.alpha.shape
    sm = orc.filter_smoother(r.alpha, r.indices)
    # independent restatement: follow ancestors backwards
    want = np.array(r.alpha)
    b = np.arange(N)
    for t in range(n1 - 2, -1, -1):
        b = r.indices[b, t]
        want[:, t, :] = r.alpha[:, t, b]
    assert np.array_equal(sm, want)
    mean, cov = orc.summary(sm)
    np.testing.assert_allclose(mean, sm.mean(axis=2), rtol=1e-12, atol=1e-12)
    for t in range(n1):
        np.testing.assert_allclose(cov[:, :, t], np.cov(sm[:, t, :], bias=True), rtol=1e-9, atol=1e-12)
    at, att, Pt, Ptt = orc.filter_summary(r.alpha, r.weights)
    W = r.weights / r.weights.sum(axis=0)
    np.testing.assert_allclose(at, r.alpha.mean(axis=2), rtol=1e-12, atol=1e-12)
    for t in range(n1 - 1):
        np.testing.assert_allclose(att[:, t], r.alpha[:, t, :] @ W[:, t], rtol=1e-11, atol=1e-12)
        dv = r.alpha[:, t, :] - att[:, t][:, None]
        np.testing.assert_allclose(Ptt[:, :, t], (dv * W[:, t]) @ dv.T, rtol=1e-9, atol=1e-12)


def test_premature_stop_returns_minus_inf():
    """all weights underflow -> -Inf, later columns stay zero (src/model_ssm_ung.cpp:638-642)"""
    import bssm_b200
    y = np.full(10, 5.0)
    y[4] = 1e6  # impossible count under the prior: y*s - exp(s) is fine, so use huge negative signal
    model = bssm_b200.ar1_ng(y, rho=0.5, sigma=0.1, mu=-1e5, distribution="poisson")
    r = orc.bsf_filter(model, 20, seed=1)
    # exp(signal) underflows, y*s ~ -5e5 for all particles: max finite -> still fine;
    # what must hold in any case: stopped_at consistent with loglik
    assert (r.stopped_at >= 0) == (r.loglik == -np.inf)
    if r.stopped_at >= 0:
        assert np.all(r.weights[:, r.stopped_at + 1:] == 0)


def test_general_m_approximation_reduces_to_scalar_restatement():
    """include/bssm_b200/ung_gauss_approx.h (any m) == the oracle's independent scalar
    restatement of approximate / smoother_ccov / conditional_cov, bit for bit, at m = 1"""
    for model in (data.svm_model(80), data.svm_model(60, missing=(0, 3, 4))):
        a = orc.psi_filter(model, 150, seed=1)
        b = orc.psi_filter_m(model, 150, seed=1)
        pre = orc.psi_prepass(model)
        assert a.loglik == b.loglik and a.approx_loglik == b.approx_loglik
        assert np.array_equal(a.indices, b.indices) and np.array_equal(a.alpha, b.alpha)
        assert np.array_equal(pre.alphahat, b.alphahat[0])
        assert np.array_equal(pre.L, b.L[0, 0]) and np.array_equal(pre.Ct, b.Ct[0, 0])


def _rts_smoother(y, Z, H2, T, RR, a1, P1):
    """textbook Kalman filter + RTS smoother for y_t = Z'a_t + eps, a_{t+1} = T a_t + eta (numpy)"""
    n, m = len(y), len(a1)
    ap, Pp = np.zeros((n + 1, m)), np.zeros((n + 1, m, m))
    af, Pf = np.zeros((n, m)), np.zeros((n, m, m))
    ap[0], Pp[0] = a1, P1
    for t in range(n):
        if np.isfinite(y[t]):
            F = Z @ Pp[t] @ Z + H2[t]
            K = Pp[t] @ Z / F
            af[t] = ap[t] + K * (y[t] - Z @ ap[t])
            Pf[t] = Pp[t] - np.outer(K, K) * F
        else:
            af[t], Pf[t] = ap[t], Pp[t]
        ap[t + 1] = T @ af[t]
        Pp[t + 1] = T @ Pf[t] @ T.T + RR
    asm, Psm = np.zeros((n + 1, m)), np.zeros((n + 1, m, m))
    asm[n], Psm[n] = ap[n], Pp[n]
    for t in range(n - 1, -1, -1):
        J = Pf[t] @ T.T @ np.linalg.pinv(Pp[t + 1])
        asm[t] = af[t] + J @ (asm[t + 1] - ap[t + 1])
        Psm[t] = Pf[t] + J @ (Psm[t + 1] - Pp[t + 1]) @ J.T
    # smoothed state BEFORE seeing... the reference smooths the one-step predictions: a_t | y_1..n
    return asm, Psm, af, Pf, ap, Pp


def test_general_m_approximation_against_numpy_smoother():
    """alphahat and L L' + (conditional structure) of the m = 2 Gaussian approximation agree
    with an independent numpy Kalman/RTS smoother run on the same pseudo-observations"""
    model = data.poisson_bsm_model(60)
    r = orc.psi_filter_m(model, 16, seed=1)
    # recover the pseudo-observations at the converged mode: Poisson laplace_iter (:313-317)
    n, m = model.n, model.m
    # marginal smoothed covariance from the conditional factors: V_t = L_t L_t' + C_t V_{t-1} C_t'
    V = np.zeros((n + 1, m, m))
    V[0] = r.L[:, :, 0] @ r.L[:, :, 0].T
    for t in range(1, n + 1):
        V[t] = r.L[:, :, t] @ r.L[:, :, t].T + r.Ct[:, :, t] @ V[t - 1] @ r.Ct[:, :, t].T
    # pseudo-observation model at the mode implied by alphahat: signal = Z' alphahat
    Z = model.Z[:, 0]
    T = model.T[:, :, 0]
    RR = model.R[:, :, 0] @ model.R[:, :, 0].T
    # iterate the numpy smoother to its own fixed point (Newton on the mode), like approximate()
    mode = model.initial_mode.copy()
    for _ in range(100):
        HH = 1.0 / (np.exp(mode) * model.u)
        yt = model.y * HH + mode - 1.0
        asm, Psm, *_ = _rts_smoother(yt, Z, HH, T, RR, model.a1, model.P1)
        new = asm[:n] @ Z
        if np.mean((new - mode) ** 2) < 1e-8:
            mode = new
            break
        mode = new
    HH = 1.0 / (np.exp(mode) * model.u)
    yt = model.y * HH + mode - 1.0
    # the reference linearises at the last-but-one sweep; converged to 1e-8 both agree to ~1e-4
    asm, Psm, *_ = _rts_smoother(yt, Z, HH, T, RR, model.a1, model.P1)
    np.testing.assert_allclose(r.alphahat.T, asm, atol=2e-3)
    np.testing.assert_allclose(V, Psm, atol=2e-3, rtol=1e-3)


def test_psi_m2_agrees_with_bsf():
    model = data.poisson_bsm_model()
    lp = np.array([orc.psi_filter_m(model, 100, seed=s).loglik for s in range(1, 9)])
    lb = orc.bsf_filter(model, 5000, seed=1).loglik
    assert lp.std() < 0.08
    assert abs(lp.mean() - lb) < 0.5


def _mv_kalman(y, Z, H2, T, RR, a1, P1, D):
    """textbook multivariate Kalman filter + RTS smoother with missing series (numpy):
    y n x p, Z p x m, H2 n x p observation variances.  Returns (loglik, smoothed a_t | y_1..n for t = 0..n)."""
    n, p = y.shape
    m = len(a1)
    ap, Pp = np.zeros((n + 1, m)), np.zeros((n + 1, m, m))
    af, Pf = np.zeros((n, m)), np.zeros((n, m, m))
    ap[0], Pp[0] = a1, P1
    ll = 0.0
    for t in range(n):
        o = np.isfinite(y[t])
        if o.any():
            Zo = Z[o]
            F = Zo @ Pp[t] @ Zo.T + np.diag(H2[t, o])
            v = y[t, o] - D[o] - Zo @ ap[t]
            K = Pp[t] @ Zo.T @ np.linalg.inv(F)
            af[t] = ap[t] + K @ v
            Pf[t] = Pp[t] - K @ F @ K.T
            ll -= 0.5 * (o.sum() * np.log(2 * np.pi) + np.linalg.slogdet(F)[1] + v @ np.linalg.solve(F, v))
        else:
            af[t], Pf[t] = ap[t], Pp[t]
        ap[t + 1] = T @ af[t]
        Pp[t + 1] = T @ Pf[t] @ T.T + RR
    asm = np.zeros((n + 1, m))
    asm[n] = ap[n]
    for t in range(n - 1, -1, -1):
        J = Pf[t] @ T.T @ np.linalg.pinv(Pp[t + 1])
        asm[t] = af[t] + J @ (asm[t + 1] - ap[t + 1])
    # the reference smooths the one-step predictions a_t | y_1..n; asm[t] above is the
    # filtered-to-smoothed estimate of the same state
    return ll, asm


def _gaussian_mng(n=40, gen_seed=3):
    rng = np.random.default_rng(gen_seed)
    a = np.zeros((n, 2))
    for t in range(1, n):
        a[t] = [a[t - 1, 0] + a[t - 1, 1] + 0.3 * rng.normal(), 0.9 * a[t - 1, 1] + 0.1 * rng.normal()]
    Z = np.array([[1.0, 0.0], [0.5, 1.0]])
    D = np.array([0.2, -0.1])
    y = a @ Z.T + D + rng.normal(size=(n, 2)) * np.array([0.5, 0.8])
    y[4, 1] = np.nan
    y[9, :] = np.nan
    T = np.array([[1.0, 1.0], [0.0, 0.9]])
    R = np.diag([0.3, 0.1])
    mdl = bssm_b200.ssm_mng(y, Z, T, R, a1=[0.0, 0.0], P1=np.diag([2.0, 0.5]),
                            distribution="gaussian", phi=[0.5, 0.8], D=D)
    return mdl, (y, Z, T, R, D)


def test_mng_gaussian_approximation_is_exact():
    """all-Gaussian ssm_mng: the 'approximation' is the model itself, so approx_loglik is the exact
    Kalman log-likelihood, alphahat the exact smoother, and every psi weight is 1"""
    mdl, (y, Z, T, R, D) = _gaussian_mng()
    n = mdl.n
    H2 = np.broadcast_to(np.array([0.25, 0.64]), (n, 2))
    ll, asm = _mv_kalman(y, Z, H2, T, R @ R.T, mdl.a1, mdl.P1, D)
    al, ah, mode, sc = orc.mng_approx(mdl)
    assert al == pytest.approx(ll, rel=1e-10)
    np.testing.assert_allclose(ah.T, asm, atol=1e-9)
    r = orc.mng_psi_filter(mdl, 64, seed=1)
    assert r.approx_loglik == al
    assert np.all(np.abs(r.weights - 1.0) < 1e-9)
    assert r.loglik == pytest.approx(ll, abs=1e-8)


def test_mng_approximation_is_a_fixed_point():
    """mixed Poisson / negative binomial / Gaussian ssm_mng: the converged mode reproduces itself
    through an independent numpy smoother run on the pseudo-observations built from it"""
    # no missing value in the Gaussian series: its initial mode is y itself (R/init_mode.R:24-26), so a
    # missing value there makes the convergence measure NaN and stops the sweeps after the first one
    # (src/model_ssm_mng.cpp:98-112) -- restated faithfully, but then there is no fixed point to check
    mdl = data.mng_model(50, missing=False)
    mdl.y[3, 0] = np.nan
    mdl.y[11, 1] = np.nan
    al, ah, mode, sc = orc.mng_approx(mdl)
    n, p = mdl.n, mdl.p
    Z, T, R, D = mdl.Z[:, :, 0], mdl.T[:, :, 0], mdl.R[:, :, 0], mdl.D[:, 0]
    s = mode.T  # n x p
    y, u = mdl.y, mdl.u
    HH, yt = np.zeros((n, p)), np.zeros((n, p))
    HH[:, 0] = 1.0 / (np.exp(s[:, 0]) * u[:, 0]); yt[:, 0] = y[:, 0] * HH[:, 0] + s[:, 0] - 1.0
    e = np.exp(s[:, 1]) * u[:, 1]; phi = mdl.phi[1]
    HH[:, 1] = (phi + e) ** 2 / (phi * e * (y[:, 1] + phi)); yt[:, 1] = s[:, 1] + (phi + e) * (y[:, 1] - e) / ((y[:, 1] + phi) * e)
    HH[:, 2] = mdl.phi[2] ** 2; yt[:, 2] = y[:, 2]
    ll, asm = _mv_kalman(yt, Z, HH, T, R @ R.T, mdl.a1, mdl.P1, D)
    np.testing.assert_allclose(ah.T, asm, atol=5e-4)
    np.testing.assert_allclose(asm[:n] @ Z.T + D, s, atol=5e-4)


def test_mng_psi_agrees_with_bsf():
    mdl = data.mng_model(40, missing=False)
    mdl.y[5, 0] = np.nan
    mdl.y[9, :2] = np.nan
    lp = np.array([orc.mng_psi_filter(mdl, 100, seed=s).loglik for s in range(1, 9)])
    lb = np.array([orc.mng_bsf_filter(mdl, 4000, seed=s).loglik for s in range(1, 4)])
    assert lp.std() < 0.1
    assert abs(lp.mean() - lb.mean()) < 0.5


def test_reference_order_vs_repro_first_divergence_at_the_headline_size():
    """the kernels compute the sum and the cumulative sum of the weights in exact
    fixed-point arithmetic (ORC_ARITH_REPRO); the reference adds binary64 numbers in Armadillo's order
    (ORC_ARITH_REFERENCE: two-accumulator accu, sequential cumsum).  At the headline size (C2: N = 65536,
    T = 1000) run both and REPORT the first time step at which an ancestor index differs.  DESIGN.md
    section 4 estimates a flip probability of ~6e-5 per step; the log-likelihoods must agree to 1e-12
    whatever the answer, and the report goes to profiles/r2_first_divergence.json (tools/first_divergence.py
    does the same for more seeds)."""
    import json, os
    model = data.svm_model(1000)
    a = orc.psi_filter(model, 65536, seed=1, arith=orc.ARITH_REPRO)
    b = orc.psi_filter(model, 65536, seed=1, arith=orc.ARITH_REFERENCE)
    diff = np.any(a.indices != b.indices, axis=0)
    first = int(np.argmax(diff)) if diff.any() else None
    print("first divergence step (seed 1, N = 65536, T = 1000):", first)
    assert abs(a.loglik - b.loglik) <= 1e-12 * abs(b.loglik)
    if first is None:
        assert np.array_equal(a.weights.view(np.uint64), b.weights.view(np.uint64))
    else:  # up to the first flipped ancestor the two runs are the same run
        assert np.array_equal(a.indices[:, :first], b.indices[:, :first])
        assert np.array_equal(a.weights[:, :first + 1].view(np.uint64), b.weights[:, :first + 1].view(np.uint64))


def test_parset_correction_reproduces_the_svm_update_model_path():
    """two restatements of is_correction_psi meet: the svm specialisation (update_model in C++,
    src/model_svm.cpp:12-29, scalar m = 1 approximation) and the ssm_ung specialisation with a parameter set
    (src/approx_mcmc.cpp:685-804, src/parset_ng.cpp) whose update_fn does what svm::update_model does"""
    import bssm_b200
    model = data.svm_model(60, missing=(9,))
    mo = data.svm_mcmc_output(model, 10)
    ref = orc.post_correct(model, mo, 48, is_type=2, seed=3)

    def update_fn(theta):  # svm_type 0: theta = (rho, sd_ar, sigma)
        return {"T": [[theta[0]]], "R": [[theta[1]]], "P1": [[theta[1] ** 2 / (1 - theta[0] ** 2)]], "phi": theta[2]}

    ung = bssm_b200.ssm_ung(model.y, model.Z, model.T, model.R, model.a1, model.P1, "svm", phi=model.phi, u=model.u,
                            D=model.D, C=model.C, initial_mode=model.initial_mode, update_fn=update_fn)
    res = orc.correction_parset(ung, mo, 48, is_type=2, seed=3)
    assert np.array_equal(res.loglik, ref.loglik)
    assert np.array_equal(res.weights, ref.weights)


def _tv_smoother(y, Z, D, H2, T, C, RR, a1, P1):
    """numpy Kalman filter + RTS smoother of a time-varying linear-Gaussian model with scalar observations:
    y_t = D_t + Z_t' a_t + eps_t (var H2_t), a_{t+1} = C_t + T_t a_t + eta_t (cov RR_t).  Returns a_t | y_1..n, t = 0..n-1"""
    n, m = len(y), len(a1)
    ap, Pp = np.zeros((n + 1, m)), np.zeros((n + 1, m, m))
    af, Pf = np.zeros((n, m)), np.zeros((n, m, m))
    ap[0], Pp[0] = a1, P1
    for t in range(n):
        if np.isfinite(y[t]):
            F = Z[t] @ Pp[t] @ Z[t] + H2[t]
            K = Pp[t] @ Z[t] / F
            af[t] = ap[t] + K * (y[t] - D[t] - Z[t] @ ap[t])
            Pf[t] = Pp[t] - np.outer(K, K) * F
        else:
            af[t], Pf[t] = ap[t], Pp[t]
        ap[t + 1] = C[t] + T[t] @ af[t]
        Pp[t + 1] = T[t] @ Pf[t] @ T[t].T + RR[t]
    asm = np.zeros((n, m))
    nxt = None
    for t in range(n - 1, -1, -1):
        if nxt is None:
            asm[t] = af[t]
        else:
            J = Pf[t] @ T[t].T @ np.linalg.pinv(Pp[t + 1])
            asm[t] = af[t] + J @ (nxt - ap[t + 1])
        nxt = asm[t]
    return asm


@pytest.mark.parametrize("kind", ["growth", "ar_exp", "sin_exp"])
def test_nlg_approximation_is_a_fixed_point(kind):
    """ssm_nlg::approximate (src/model_ssm_nlg.cpp:62-150) linearises Z_fn / T_fn at the mode and smooths the
    linearised model until the mode reproduces itself.  Independent check of include/bssm_b200/nlg_gauss_approx.h:
    the converged mode is a fixed point of a numpy smoother of the model linearised (numerical Jacobians) at it."""
    if kind == "growth":
        mdl = data.growth_model(60, missing=(7,))
        H, R0, R1 = np.exp(mdl.theta)
        dT, K, a11, a12, P11, P12 = mdl.known_params[:6]
        Zf = lambda a: a[1]
        def Tf(a):
            r = np.exp(a[0]) / (1 + np.exp(a[0]))
            e = np.exp(r * dT)
            return np.array([a[0], K * a[1] * e / (K + a[1] * (e - 1))])
        RR, a1, P1 = np.diag([R0 ** 2, R1 ** 2]), np.array([a11, a12]), np.diag([P11, P12])
    elif kind == "ar_exp":
        mdl = data.ar_exp_model(60, missing=(3,))
        mu, rho, ls, lH = mdl.theta
        H = np.exp(lH)
        Zf = lambda a: np.exp(a[0])
        Tf = lambda a: np.array([mu * (1 - rho) + rho * a[0]])
        RR, a1, P1 = np.array([[np.exp(ls) ** 2]]), np.array([mu]), np.array([[np.exp(ls) ** 2 / (1 - rho ** 2)]])
    else:
        mdl = data.sin_exp_model(60)
        H, R = np.exp(mdl.theta)
        Zf = lambda a: np.exp(a[0])
        Tf = lambda a: np.array([np.sin(a[0])])
        RR, a1, P1 = np.array([[R ** 2]]), np.array([0.0]), np.array([[1.0]])
    r = orc.nlg_psi_filter(mdl, 8, seed=1)
    mode = r.mode.T.copy()  # n x m
    n, m = mode.shape
    h = 1e-6
    Z, D, T, C = np.zeros((n, m)), np.zeros(n), np.zeros((n, m, m)), np.zeros((n, m))
    for t in range(n):
        for d in range(m):
            e = np.zeros(m); e[d] = h
            Z[t, d] = (Zf(mode[t] + e) - Zf(mode[t] - e)) / (2 * h)
            T[t, :, d] = (Tf(mode[t] + e) - Tf(mode[t] - e)) / (2 * h)
        D[t] = Zf(mode[t]) - Z[t] @ mode[t]
        C[t] = Tf(mode[t]) - T[t] @ mode[t]
    asm = _tv_smoother(mdl.y, Z, D, np.full(n, H ** 2), T, C, [RR] * n, a1, P1)
    np.testing.assert_allclose(asm, mode, atol=2e-3, rtol=1e-3)


def test_bootstrap_correction_of_nlg_models_is_the_filter_per_draw():
    """orc.is_correction_bsf_nlg (src/approx_mcmc.cpp:442-518) = update_model(theta_i) + bsf_filter with stream i:
    weight_i = exp(loglik_i - approx_loglik_i), expanded draws for is3"""
    model = data.growth_model(30)
    rng = np.random.default_rng(2)
    ns = 5
    theta = np.asarray(model.theta)[None, :] + 0.05 * rng.normal(size=(ns, 3))
    counts = np.array([1, 3, 2, 1, 2], dtype=np.uint32)
    approx = -60.0 + rng.normal(size=ns)
    mo = {"theta": theta, "counts": counts, "approx_loglik": approx, "output_type": 3, "posterior": np.zeros(ns)}
    r = orc.is_correction_bsf_nlg(model, mo, 50, is_type=2, seed=4)
    for i in range(ns):
        mi = bssm_b200.ssm_nlg(model.y, "growth", theta=theta[i], known_params=model.known_params)
        ll = orc.nlg_bsf_filter(mi, 50, seed=4, filt=i).loglik
        assert r.loglik[i] == ll
        assert r.weights[i] == pytest.approx(np.exp(ll - approx[i]), rel=1e-12)
        assert r.posterior[i] == pytest.approx(ll - approx[i], abs=1e-10)
    r3 = orc.is_correction_bsf_nlg(model, mo, 50, is_type=3, seed=4)
    assert r3.weights.size == counts.sum()
    src = np.repeat(np.arange(ns), counts)
    mi = bssm_b200.ssm_nlg(model.y, "growth", theta=theta[src[4]], known_params=model.known_params)
    assert r3.loglik[4] == orc.nlg_bsf_filter(mi, 50, seed=4, filt=4).loglik


def test_oracle_runs_user_sde_functions_like_the_registry_model(tmp_path):
    """the oracle takes an ssm_sde model as four function pointers (src/model_ssm_sde.h:8-11); the GBM example
    compiled from source with g++ and passed that way gives the registry model's bits"""
    import ctypes as C
    import os
    import subprocess
    src = r'''
    double drift(double x, const double* theta) { return theta[0] * x; }
    double diffusion(double x, const double* theta) { const double tx = theta[1] * x; return tx > 0.0 ? tx : 0.0; }
    double ddiffusion(double x, const double* theta) { return theta[1] * (x > 0.0 ? 1.0 : 0.0); }
    double log_obs_density(double y, double x, const double* theta) {
      const double q = (y - pf_log(x)) / theta[2];
      return -(0.918938533204672741780329736406 + 0.5 * q * q + pf_log(theta[2]));
    }'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cpp = tmp_path / "gbm.cpp"
    cpp.write_text('#include "bssm_b200/pf_math.h"\nextern "C" {\n' + src + "\n}\n")
    so = tmp_path / "gbm.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I" + os.path.join(root, "include"), str(cpp),
                    "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    f = orc.L().orc_sde_set_user
    f.restype = None
    f.argtypes = [C.c_void_p] * 4
    f(*[C.cast(getattr(lib, nm), C.c_void_p) for nm in ("drift", "diffusion", "ddiffusion", "log_obs_density")])
    reg = data.gbm_model(25)
    user = bssm_b200.ssm_sde(reg.y, 1000, reg.theta, reg.x0, reg.positive)
    a = orc.sde_bsf_filter(reg, 64, 3, seed=5)
    b = orc.sde_bsf_filter(user, 64, 3, seed=5)
    assert a.loglik == b.loglik and np.array_equal(a.indices, b.indices) and np.array_equal(a.alpha, b.alpha)


def test_jacobi_stops_at_its_fixed_point_with_the_same_bits(tmp_path):
    """pf_jacobi_eig_sym (include/bssm_b200/pf_smallmat.h, the stand-in for arma::svd_econ on symmetric matrices,
    src/conditional_dist.cpp:14) leaves the sweep loop at the first sweep that changes nothing instead of always
    running 60: same eigenvalues and eigenvectors, bit for bit, as the plain rule kept in tests/aux/"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "jac"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "aux", "jacobi_fixed_point_check.cpp"), "-o", str(exe)], check=True)
    tot, bad = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert tot == 13 * 1500 and bad == 0


def _dense_conditional_mean(model, ytilde, Htilde, ysim):
    """E[alpha_0..alpha_n | y] of the linear-Gaussian model (ytilde's structure, sd Htilde) by conditioning the
    joint normal of all states and the observed y with numpy: mean and covariance of the stacked states from
    the recursion alpha_{t+1} = C + T alpha_t + R eta, then the textbook formula."""
    n, m = model.n, model.m
    Z = np.asarray(model.Z).reshape(m, -1)
    T = np.asarray(model.T).reshape(m, m, -1)
    R = np.asarray(model.R).reshape(m, model.k, -1)
    Cm = np.asarray(model.C).reshape(m, -1)
    D = np.ravel(model.D)
    tv = lambda a, t: a[..., t if a.shape[-1] > 1 else 0]
    mu = np.zeros((n + 1, m))
    S = np.zeros((n + 1, m, n + 1, m))
    mu[0] = np.ravel(model.a1)
    S[0, :, 0, :] = np.asarray(model.P1).reshape(m, m)
    for t in range(n):
        Tt, Rt = tv(T, t), tv(R, t)
        mu[t + 1] = tv(Cm, t) + Tt @ mu[t]
        for s in range(t + 1):
            S[t + 1, :, s, :] = Tt @ S[t, :, s, :]
            S[s, :, t + 1, :] = S[t + 1, :, s, :].T
        S[t + 1, :, t + 1, :] = Tt @ S[t, :, t, :] @ Tt.T + Rt @ Rt.T
    Sf = S.reshape((n + 1) * m, (n + 1) * m)
    obs = [t for t in range(n) if np.isfinite(ysim[t])]
    A = np.zeros((len(obs), (n + 1) * m))
    for r, t in enumerate(obs):
        A[r, t * m:(t + 1) * m] = tv(Z, t)
    d = np.array([tv(D, t) for t in obs])
    Syy = A @ Sf @ A.T + np.diag(Htilde[obs] ** 2)
    resid = ysim[obs] - (A @ mu.ravel() + d)
    return (mu.ravel() + Sf @ A.T @ np.linalg.solve(Syy, resid)).reshape(n + 1, m).T


def test_simulation_smoother_against_dense_gaussian_conditioning():
    """The draws of the oracle's simulation smoother (ssm_ulg::simulate_states, src/model_ssm_ulg.cpp:141-217)
    are alpha+ - E[alpha | y+] + E[alpha | ytilde]: rebuilt here with numpy from the same normals by conditioning
    the dense joint normal -- pins the three-sweep smoother with stored gains, missing values included."""
    rng = np.random.default_rng(11)
    n = 12
    y = rng.poisson(4, size=n).astype(float)
    y[3] = np.nan
    y[n - 1] = np.nan
    model = bssm_b200.bsm_ng(y, sd_level=0.3, sd_slope=0.05, distribution="poisson", P1=np.diag([1.0, 0.5]))
    m, k = model.m, model.k
    seed, nsim = 5, 7
    alpha, w, ll, ahat, yt, Ht = orc.importance_sample(model, nsim, seed, approx=True)
    assert np.abs(ahat - _dense_conditional_mean(model, yt, Ht, yt)).max() < 1e-9
    Z = np.asarray(model.Z).reshape(m, -1)[:, 0]
    T = np.asarray(model.T).reshape(m, m, -1)[:, :, 0]
    R = np.asarray(model.R).reshape(m, k, -1)[:, :, 0]
    L1 = np.linalg.cholesky(np.asarray(model.P1).reshape(m, m))
    nsim2 = nsim // 2
    for i in list(range(nsim2)) + [nsim2]:
        ap = np.zeros((m, n + 1))
        yp = np.full(n, np.nan)
        ap[:, 0] = np.ravel(model.a1) + L1 @ np.array([orc.sim_normal(seed, 0, i * m + d) for d in range(m)])
        for t in range(n):
            if np.isfinite(y[t]):
                yp[t] = Z @ ap[:, t] + Ht[t] * orc.sim_normal(seed, t + 1, i * (k + 1))
            uk = np.array([orc.sim_normal(seed, t + 1, i * (k + 1) + 1 + j) for j in range(k)])
            ap[:, t + 1] = T @ ap[:, t] + R @ uk
        draw = ap - _dense_conditional_mean(model, yt, Ht, yp) + ahat
        if i < nsim2:
            assert np.abs(alpha[:, :, i] - draw).max() < 1e-9
            assert np.abs(alpha[:, :, i + nsim2] - (2 * ahat - draw)).max() < 1e-9
        else:
            assert np.abs(alpha[:, :, nsim - 1] - draw).max() < 1e-9  # the unpaired last draw of an odd nsim


def test_spdk_loglik_agrees_with_the_psi_filter():
    """SPDK and psi-APF estimate the same likelihood (src/model_ssm_ung.cpp:196-211): both correct approx_loglik
    by importance weights of the same form"""
    model = data.poisson_bsm_model()
    _, w, ll, _ = orc.importance_sample(model, 4000, 3)
    r = orc.psi_filter_m(model, 4000, 3)
    assert ll[1] == r.approx_loglik
    assert abs(ll[0] - r.loglik) < 0.05
    assert w.max() == 1.0 and w.min() > 0.0
    # svm (m = 1)
    svm = data.svm_model(n=200)
    _, _, ll, _ = orc.importance_sample(svm, 4000, 3)
    r = orc.psi_filter(svm, 4000, 3)
    assert ll[1] == r.approx_loglik
    assert abs(ll[0] - r.loglik) < 0.3


def test_spdk_correction_oracle_properties():
    """is_correction_spdk restatement (src/approx_mcmc.cpp:606-678, :923-1036): results do not depend on the number
    of threads (draw i owns stream i), is1 with unit counts is is2, is3 repeats a stored draw on fresh streams, and
    a draw whose stored mode is the converged mode reproduces logLik(method = "spdk") - approx_loglik"""
    from test_gpu_parset import draws, ung_update_model
    model = ung_update_model(m=2, tv=True)
    mo = draws(model, 9)
    w1, p1 = orc.correction_spdk_parset(model, mo, 12, is_type=2, seed=5, threads=1)
    w4, p4 = orc.correction_spdk_parset(model, mo, 12, is_type=2, seed=5, threads=4)
    assert np.array_equal(w1, w4) and np.array_equal(p1, p4)
    np.testing.assert_array_equal(p1, mo["approx_loglik"] + np.log(w1))
    ones = dict(mo, counts=np.ones(9, dtype=np.uint32))
    assert np.array_equal(orc.correction_spdk_parset(model, ones, 12, is_type=1, seed=5)[0],
                          orc.correction_spdk_parset(model, ones, 12, is_type=2, seed=5)[0])
    w3, _ = orc.correction_spdk_parset(model, mo, 12, is_type=3, seed=5)
    assert w3.size == int(mo["counts"].sum()) and len(set(w3)) == w3.size
    assert w3[0] == w1[0]  # expanded draw 0 is stored draw 0 on stream 0
    # one stored draw = the model itself at its converged mode: weight = exp(logLik_spdk - approx_loglik)
    base = data.poisson_bsm_model()
    same = bssm_b200.ssm_ung(base.y, base.Z, base.T, base.R, base.a1, base.P1, "poisson", u=base.u,
                             update_fn=lambda theta: {"R": np.array(base.R, dtype=float)})
    mode = np.zeros(base.n)
    orc.L().orc_ung_approx_mode_m.argtypes = [orc.C.POINTER(orc.UngModel), orc.c_double_p]
    ms = same.c_struct()
    orc.L().orc_ung_approx_mode_m(orc.C.byref(ms), orc.dptr(mode))
    one = {"theta": np.zeros((1, 1)), "counts": np.ones(1, dtype=np.uint32), "modes": mode[:, None], "output_type": 3,
           "posterior": np.zeros(1)}
    w, _ = orc.correction_spdk_parset(same, one, 500, is_type=2, seed=7)
    _, _, ll, _ = orc.importance_sample(base, 500, 7)
    # (to the tolerance of approximate(): its pseudo-observations are linearised at the last-but-one iterate)
    assert abs(np.log(w[0]) - (ll[0] - ll[1])) < 1e-4
