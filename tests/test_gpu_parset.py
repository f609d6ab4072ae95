"""Post-correction of models whose theta -> model map is a user function (ssm_ung / ssm_mng with update_fn):
postcorrection_nongaussian(model_type = 0 | 1) -> is_correction_psi / is_correction_bsf with the parameter
sets of src/parset_ng.cpp (src/approx_mcmc.cpp:685-921, :1038-1273).  CUDA path vs the oracle, bit for bit."""
import numpy as np
import pytest

import data
import orc

pytestmark = pytest.mark.gpu


def ung_update_model(n=50, m=2, tv=False, seed=3):
    """an ssm_ung Poisson model (level + slope, or AR(1) for m = 1) whose update_fn moves phi-free pieces:
    Z scale, T damping, R sds, a1, P1, D and C -- every field parset_ung can carry"""
    import bssm_b200
    rng = np.random.default_rng(seed)
    y = rng.poisson(3.0, size=n).astype(float)
    y[4] = np.nan
    u = np.full(n, 1.5)
    if m == 1:
        Z, T, R, a1, P1 = [1.0], [[0.9]], [[0.2]], [0.1], [[0.5]]
    else:
        Z, T, R = [1.0, 0.0], [[1.0, 1.0], [0.0, 0.9]], np.diag([0.1, 0.05])
        a1, P1 = [0.5, 0.0], np.diag([0.5, 0.1])
    if tv:  # time-varying Z and D: the slices then carry n columns
        Z = np.tile(np.reshape(Z, (m, 1)), (1, n)) * (1.0 + 0.01 * np.arange(n))
        D = 0.02 * np.arange(n)
    else:
        D = 0.3
    base = dict(Z=np.array(Z, dtype=float), T=np.array(T, dtype=float).reshape(m, m), R=np.array(R, dtype=float).reshape(m, -1),
                a1=np.array(a1, dtype=float), P1=np.array(P1, dtype=float).reshape(m, m))

    def update_fn(theta):
        out = {"T": base["T"].copy(), "R": base["R"] * np.exp(theta[1]), "a1": base["a1"] + theta[2],
               "P1": base["P1"] * np.exp(2 * theta[1]), "Z": base["Z"] * (1.0 + 0.1 * theta[0]),
               "D": np.asarray(D, dtype=float) + theta[2], "C": np.full((m, 1), 0.01 * theta[0])}
        out["T"][m - 1, m - 1] = 0.9 * np.tanh(1.5 + theta[0])
        return out

    return bssm_b200.ssm_ung(y, base["Z"], base["T"], base["R"], base["a1"], base["P1"], "poisson", u=u, D=D,
                             update_fn=update_fn)


def draws(model, ns, n_par=3, seed=11, rows=None):
    rng = np.random.default_rng(seed)
    theta = 0.1 * rng.normal(size=(ns, n_par))
    counts = (1 + rng.poisson(2, size=ns)).astype(np.uint32)
    rows = model.n if rows is None else rows
    modes = np.log(3.0) + 0.1 * rng.normal(size=(rows, ns))  # approximate_for_is takes any mode
    return {"theta": theta, "counts": counts, "modes": np.asfortranarray(modes), "output_type": 3,
            "posterior": np.zeros(ns), "approx_loglik": -120.0 + rng.normal(size=ns)}


@pytest.mark.parametrize("m,tv", [(1, False), (2, False), (2, True)])
def test_ung_parset_psi(engine, m, tv):
    model = ung_update_model(m=m, tv=tv)
    mo = draws(model, 12)
    ref = orc.correction_parset(model, mo, 96, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 96, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    assert len(set(res["loglik"])) == 12
    mo1 = dict(mo, output_type=1)
    ref1 = orc.correction_parset(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.correction_parset(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("m,tv", [(1, False), (2, True)])
def test_ung_parset_bsf(engine, m, tv):
    model = ung_update_model(m=m, tv=tv)
    mo = draws(model, 10)
    ref = orc.correction_parset(model, mo, 128, is_type=2, seed=4, bsf=True)
    res = engine.post_correct(model, mo, 128, is_type="is2", seed=4, return_loglik=True, sampling_method="bsf")
    np.testing.assert_allclose(res["loglik"], ref.loglik, rtol=1e-13)  # constants summed per draw on the host
    np.testing.assert_allclose(res["weights"], ref.weights, rtol=1e-10)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.correction_parset(model, mo1, 64, is_type=3, seed=6, bsf=True)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=6, sampling_method="bsf")
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.correction_parset(model, mo2, 32, is_type=1, seed=8, bsf=True)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=8, sampling_method="bsf")
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


def test_ung_parset_seasonal(engine):
    """m = 5 (run-time state dimension): weights, trajectories and summaries"""
    import bssm_b200
    base = data.poisson_bsm_model(60)
    y = base.y
    m = 5
    T = np.zeros((m, m)); T[0, 0] = T[0, 1] = T[1, 1] = 1.0
    T[2, 2:] = -1.0; T[3, 2] = 1.0; T[4, 3] = 1.0
    R = np.zeros((m, 3)); R[0, 0], R[1, 1], R[2, 2] = 0.1, 0.01, 0.05
    Z = np.array([1.0, 0.0, 1.0, 0.0, 0.0])

    def update_fn(theta):
        Ri = R.copy()
        Ri[0, 0], Ri[1, 1], Ri[2, 2] = 0.1 * np.exp(theta[0]), 0.01 * np.exp(theta[1]), 0.05 * np.exp(theta[2])
        return {"R": Ri}

    model = bssm_b200.ssm_ung(y, Z, T, R, np.zeros(m), np.diag([1.0, 0.1, 0.5, 0.5, 0.5]), "poisson", update_fn=update_fn)
    mo = draws(model, 8)
    mo["modes"] = np.asfortranarray(np.log(np.maximum(np.nan_to_num(y, nan=1.0), 0.5))[:, None] + 0.05 * mo["modes"])
    ref = orc.correction_parset(model, mo, 64, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    refb = orc.correction_parset(model, mo, 64, is_type=2, seed=3, bsf=True)
    resb = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True, sampling_method="bsf")
    np.testing.assert_allclose(resb["loglik"], refb.loglik, rtol=1e-13)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.correction_parset(model, mo1, 32, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 32, is_type="is3", seed=5)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.correction_parset(model, mo2, 32, is_type=1, seed=7, bsf=True)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7, sampling_method="bsf")
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-11)


def mng_update_model():
    base = data.mng_model(40)

    def update_fn(theta):
        T = base.T[:, :, 0].copy()
        T[1, 1] = 0.95 * np.tanh(2.0 + theta[0])
        return {"T": T, "R": base.R[:, :, 0] * np.exp(theta[1]), "phi": base.phi * np.exp(0.5 * theta[2]),
                "Z": base.Z[:, :, 0] * (1.0 + 0.05 * theta[0]), "D": base.D[:, 0] + 0.1 * theta[2],
                "a1": base.a1 + theta[1], "P1": base.P1 * np.exp(theta[2]), "C": np.full((2, 1), 0.01 * theta[1])}

    base.update_fn = update_fn
    return base


def test_mng_parset_psi_and_bsf(engine):
    model = mng_update_model()
    _, _, mode, _ = orc.mng_approx(model)
    mo = draws(model, 10, rows=model.n * model.p)
    mo["modes"] = np.asfortranarray(mode.ravel(order="F")[:, None] + 0.05 * (mo["modes"] - np.log(3.0)))
    ref = orc.correction_parset(model, mo, 96, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 96, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.correction_parset(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.correction_parset(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)
    mo["approx_loglik"] = orc.mng_bsf_filter(model, 256, seed=9).loglik + 0.3 * np.random.default_rng(1).normal(size=10)
    refb = orc.correction_parset(model, mo, 128, is_type=2, seed=4, bsf=True)
    resb = engine.post_correct(model, mo, 128, is_type="is2", seed=4, return_loglik=True, sampling_method="bsf")
    np.testing.assert_allclose(resb["loglik"], refb.loglik, rtol=1e-13)
    np.testing.assert_allclose(resb["weights"], refb.weights, rtol=1e-10)
    refb1 = orc.correction_parset(model, mo1, 64, is_type=3, seed=6, bsf=True)
    resb1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=6, sampling_method="bsf")
    assert np.array_equal(resb1["alpha"].view(np.uint64), refb1.alpha.view(np.uint64))


def test_parset_path_meets_the_svm_path(engine):
    """the same draws through both routes of postcorrection_nongaussian: model_type 3 (svm::update_model in the
    library, approximation per draw on the device) and model_type 1 (parameter set built from an update_fn that does
    what svm::update_model does, approximation per draw on the host)"""
    import bssm_b200
    model = data.svm_model(80, missing=(9,))
    mo = data.svm_mcmc_output(model, 12)

    def update_fn(theta):
        return {"T": [[theta[0]]], "R": [[theta[1]]], "P1": [[theta[1] ** 2 / (1 - theta[0] ** 2)]], "phi": theta[2]}

    ung = bssm_b200.ssm_ung(model.y, model.Z, model.T, model.R, model.a1, model.P1, "svm", phi=model.phi, u=model.u,
                            D=model.D, C=model.C, initial_mode=model.initial_mode, update_fn=update_fn)
    a = engine.post_correct(model, mo, 96, is_type="is2", seed=3, return_loglik=True)
    b = engine.post_correct(ung, mo, 96, is_type="is2", seed=3, return_loglik=True)
    ref = orc.correction_parset(ung, mo, 96, is_type=2, seed=3)
    assert np.array_equal(b["loglik"], ref.loglik)
    np.testing.assert_allclose(a["loglik"], b["loglik"], rtol=1e-12)


def test_suggest_N_for_models_with_update_fn(engine):
    """suggest_n_nongaussian(model_type = 1 | 0): model.update_model(theta, update_fn) then the sd of replicate
    psi-APF log-likelihoods per candidate N (src/R_postcorrection.cpp:14-104)"""
    import copy
    reps, cands = 6, [16, 64]
    model = ung_update_model(m=2)
    theta = np.array([0.05, -0.1, 0.02])
    res = engine.suggest_N(model, theta, candidates=cands, replications=reps, seed=5)
    mi = copy.copy(model)
    for name, val in model.update_fn(theta).items():
        setattr(mi, name, np.asfortranarray(np.reshape(np.asarray(val, dtype=np.float64), np.shape(getattr(model, name)), order="F")))
    for c, N in enumerate(cands):
        ll = [orc.psi_filter_m(mi, N, seed=5, filt=c * reps + r).loglik for r in range(reps)]
        assert res["results"]["sd"][c] == pytest.approx(np.std(ll, ddof=1), rel=1e-10)
    g = mng_update_model()
    thg = np.array([0.1, -0.05, 0.2])
    resg = engine.suggest_N(g, thg, candidates=cands, replications=reps, seed=7)
    gi = copy.copy(g)
    for name, val in g.update_fn(thg).items():
        setattr(gi, name, np.asfortranarray(np.reshape(np.asarray(val, dtype=np.float64), np.shape(getattr(g, name)), order="F")))
    for c, N in enumerate(cands):
        ll = [orc.mng_psi_filter(gi, N, seed=7, filt=c * reps + r).loglik for r in range(reps)]
        assert resg["results"]["sd"][c] == pytest.approx(np.std(ll, ddof=1), rel=1e-10)


def test_parset_argument_errors(engine):
    model = ung_update_model(m=2)
    mo = draws(model, 4)
    good = model.update_fn
    model.update_fn = lambda th: {"T": np.zeros((3, 3))}  # wrong shape
    with pytest.raises(ValueError, match="elements"):
        engine.post_correct(model, mo, 32, is_type="is2", seed=1)
    model.update_fn = lambda th: {"Q": 1.0}  # not a model component
    with pytest.raises(ValueError, match="unknown components"):
        engine.post_correct(model, mo, 32, is_type="is2", seed=1)
    model.update_fn = good
    bad = dict(mo, modes=mo["modes"][:, :3])
    with pytest.raises(ValueError, match="disagree"):
        engine.post_correct(model, bad, 32, is_type="is2", seed=1)
    with pytest.raises(Exception, match="k <= m <= 4"):
        engine.nlg_register("__device__ double Z_fn(int t, const double* a, const double* th, const double* kp) { return 0; }", 5, 1)
    g = data.mng_model(30)
    with pytest.raises(ValueError, match="update_fn"):
        engine.post_correct(g, mo, 32, is_type="is2", seed=1)


def test_mng_device_approximation_equals_host_approximation(engine, monkeypatch):
    """the per-draw approximation of ssm_mng parameter sets (multivariate Kalman filter with masked missing series,
    Cholesky of F_t, backward recursion, conditional_cov) on the device vs the host loop over mng_gauss_approx.h"""
    model = mng_update_model()
    _, _, mode, _ = orc.mng_approx(model)
    mo = draws(model, 10, rows=model.n * model.p)
    mo["modes"] = np.asfortranarray(mode.ravel(order="F")[:, None] + 0.05 * (mo["modes"] - np.log(3.0)))
    mo = dict(mo, output_type=2)
    dev = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    monkeypatch.setenv("BSSM_B200_HOST_APPROX", "1")
    host = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    monkeypatch.delenv("BSSM_B200_HOST_APPROX")
    assert np.array_equal(dev["loglik"], host["loglik"])
    assert np.array_equal(dev["alphahat"], host["alphahat"]) and np.array_equal(dev["Vt"], host["Vt"])
