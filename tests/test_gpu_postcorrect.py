"""post_correct (IS-MCMC correction, BASELINE config 4) against the oracle's
restatement of approx_mcmc::is_correction_psi (src/approx_mcmc.cpp:353-427)."""
import numpy as np
import pytest

import data
import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(engine):
    model = data.svm_model(120)
    return engine, model, data.svm_mcmc_output(model, 40)


def test_weights_only_is2(setup):
    engine, model, mo = setup
    ref = orc.post_correct(model, mo, 128, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 128, threads=1, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    np.testing.assert_allclose(res["posterior"], ref.posterior, rtol=1e-13, atol=1e-13)
    assert res["launches"] == 2  # one pre-pass + one filter launch for all draws
    # threads = number of GPUs of this process to use: more than there are is clamped, results identical
    res2 = engine.post_correct(model, mo, 128, threads=2, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res2["weights"], ref.weights) and np.array_equal(res2["loglik"], ref.loglik)


def test_trajectories_is3(setup):
    """is3 expands every stored draw `count` times; output_type 1 keeps one trajectory per draw"""
    engine, model, mo = setup
    mo1 = dict(mo, output_type=1)
    ref = orc.post_correct(model, mo1, 64, is_type=3, seed=5)
    res = engine.post_correct(model, mo1, 64, is_type="is3", seed=5)
    assert res["weights"].size == int(mo["counts"].sum())
    assert np.array_equal(res["weights"], ref.weights)
    assert np.array_equal(res["alpha"].view(np.uint64), ref.alpha.view(np.uint64))


def test_summary_is1(setup):
    """is1: nsim * count particles per draw; output_type 2: weighted running mean / variance"""
    engine, model, mo = setup
    mo2 = dict(mo, output_type=2)
    ref = orc.post_correct(model, mo2, 32, is_type=1, seed=7)
    res = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    assert np.array_equal(res["weights"], ref.weights)
    np.testing.assert_allclose(res["alphahat"], ref.alphahat, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res["Vt"], ref.Vt, rtol=1e-9, atol=1e-13)


def test_summary_is2_chunked(setup, monkeypatch):
    engine, model, mo = setup
    mo2 = dict(mo, output_type=2)
    ref = orc.post_correct(model, mo2, 64, is_type=2, seed=9)
    whole = engine.post_correct(model, mo2, 64, is_type="is2", seed=9)
    np.testing.assert_allclose(whole["alphahat"], ref.alphahat, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(whole["Vt"], ref.Vt, rtol=1e-9, atol=1e-13)


def test_sharding_gives_identical_results(setup):
    """draw d always uses stream d: two shards concatenated == one call"""
    engine, model, mo = setup
    mo1 = dict(mo, output_type=1)
    whole = engine.post_correct(model, mo1, 64, is_type="is2", seed=11)
    a = engine.post_correct(model, mo1, 64, is_type="is2", seed=11, draw_range=(0, 17))
    b = engine.post_correct(model, mo1, 64, is_type="is2", seed=11, draw_range=(17, 40))
    assert np.array_equal(np.concatenate([a["weights"], b["weights"]]), whole["weights"])
    assert np.array_equal(np.concatenate([a["alpha"], b["alpha"]], axis=2), whole["alpha"])


def test_resident_job_matches(setup):
    engine, model, mo = setup
    ref = orc.post_correct(model, mo, 256, is_type=2, seed=1)
    job = engine.PostcorrJob(model, mo, 256, is_type="is2", seed=1, draw_range=(5, 30))
    job.run()
    assert np.array_equal(job.logliks(), ref.loglik[5:30])
    job.close()


def test_ar1_ng_postcorrection(engine):
    rng = np.random.default_rng(3)
    y = rng.poisson(4.0, size=60).astype(float)
    model = engine.ar1_ng(y, rho=0.7, sigma=0.3, mu=1.2, distribution="poisson")
    th0 = np.asarray(model.theta)
    theta = th0[None, :] + 0.05 * rng.normal(size=(12, 3))
    modes = np.zeros((60, 12), order="F")
    for i in range(12):
        mi = engine.ar1_ng(y, rho=theta[i, 0], sigma=float(np.exp(theta[i, 1])), mu=theta[i, 2], distribution="poisson")
        modes[:, i] = orc.psi_prepass(mi).mode
    mo = {"theta": theta, "counts": np.ones(12, dtype=np.uint32), "modes": modes, "output_type": 3}
    ref = orc.post_correct(model, mo, 100, is_type=2, seed=2)
    res = engine.post_correct(model, mo, 100, is_type="is2", seed=2)
    assert np.array_equal(res["weights"], ref.weights)


def test_nonlinear_postcorrection_growth(engine):
    """postcorrection_nonlinear (src/R_postcorrection.cpp:252-317) on the growth model: weights,
    one trajectory per draw, and the weighted summary, against the oracle"""
    model = data.growth_model(40)
    mo = data.growth_mcmc_output(model, 12)
    ref = orc.post_correct(model, mo, 64, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.post_correct(model, mo1, 32, is_type=3, seed=4)
    res1 = engine.post_correct(model, mo1, 32, is_type="is3", seed=4)
    assert np.array_equal(res1["weights"], ref1.weights)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.post_correct(model, mo2, 32, is_type=1, seed=5)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=5)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


def _bsm_case(kind):
    import bssm_b200
    y = data.poisson_series()
    if kind == "level_slope":      # m = 2, the documented bsm_ng example
        return data.poisson_bsm_model(80)
    if kind == "level":            # m = 1: goes through the m = 1 device pre-pass
        return bssm_b200.bsm_ng(y[:70], sd_level=0.1, P1=[[1.0]], distribution="poisson")
    if kind == "level_xreg":       # m = 1 with two regression coefficients in theta: device pre-pass, one xbeta per draw
        n = 64
        rng = np.random.default_rng(5)
        return bssm_b200.bsm_ng(y[:n], sd_level=0.1, P1=[[1.0]], distribution="poisson",
                                xreg=rng.normal(size=(n, 2)), beta=[0.15, -0.1])
    if kind == "negbin_xreg":      # m = 2, negative binomial with phi and one regression coefficient in theta
        n = 60
        rng = np.random.default_rng(4)
        xreg = rng.normal(size=(n, 1))
        return bssm_b200.bsm_ng(y[:n], sd_level=0.1, sd_slope=0.02, P1=np.diag([1.0, 0.5]),
                                distribution="negative binomial", phi=4.0, xreg=xreg, beta=[0.2])
    # m = 3: level + slope + extra noise state
    return bssm_b200.bsm_ng(y[:50], sd_level=0.1, sd_slope=0.02, sd_noise=0.1, P1=np.diag([1.0, 0.5, 1.0]),
                            distribution="poisson")


@pytest.mark.parametrize("kind", ["level_slope", "level", "level_xreg", "negbin_xreg", "noise"])
def test_bsm_ng_post_correction(engine, kind):
    """is_correction_psi for bsm_ng (src/approx_mcmc.cpp:685-804): bsm_ng::update_model per draw
    (src/model_bsm_ng.cpp:17-51), Gaussian approximation at the stored mode (host, OpenMP for m > 1)"""
    model = _bsm_case(kind)
    mo = data.bsm_mcmc_output(model, 24)
    ref = orc.post_correct(model, mo, 96, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 96, is_type="is2", seed=3, return_loglik=True)
    if kind in ("level", "level_xreg"):
        # m = 1 runs the device pre-pass; the general-m host approximation of the oracle reduces to it
        np.testing.assert_allclose(res["loglik"], ref.loglik, rtol=1e-12)
    else:
        assert np.array_equal(res["loglik"], ref.loglik)
        assert np.array_equal(res["weights"], ref.weights)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.post_correct(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5)
    if kind not in ("level", "level_xreg"):
        assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.post_correct(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


def test_suggest_N(engine):
    """suggest_N (R/post_correction.R:85-134, src/R_postcorrection.cpp:14-104): sd of replicate psi-APF
    log-likelihoods per candidate N, at theta applied through svm::update_model"""
    import bssm_b200
    model = data.svm_model(100)
    theta = np.array([0.97, 0.25, 1.1])
    cands, reps = [16, 64, 256], 20
    res = engine.suggest_N(model, theta, candidates=cands, replications=reps, seed=11)
    mi = bssm_b200.svm(model.y, theta[0], theta[1], sigma=theta[2])
    mi.initial_mode = model.initial_mode  # update_model does not touch initial_mode (src/model_svm.cpp:12-29)
    for c, N in enumerate(cands):
        ll = orc.filter_batch(mi, N, "psi", seed=11, first_filter=c * reps, n_filters=reps, arith=orc.ARITH_REPRO)
        assert res["results"]["sd"][c] == pytest.approx(np.std(ll, ddof=1), rel=1e-10)
    sd = res["results"]["sd"]
    assert sd[0] > sd[2]
    assert res["N"] == next((N for N, s in zip(cands, sd) if s < 1), None)
    with pytest.raises(ValueError, match="Length of 'theta'"):
        engine.suggest_N(model, theta[:2])


def test_suggest_N_bsm_and_growth(engine):
    model = data.poisson_bsm_model(60)
    res = engine.suggest_N(model, np.asarray(model.theta) + 0.1, candidates=[20, 200], replications=12, seed=2)
    assert np.all(np.isfinite(res["results"]["sd"])) and res["results"]["sd"][1] < res["results"]["sd"][0]
    g = data.growth_model(40)
    res = engine.suggest_N(g, np.asarray(g.theta), candidates=[50, 500], replications=12, seed=2)
    assert np.all(np.isfinite(res["results"]["sd"])) and res["results"]["sd"][1] < res["results"]["sd"][0]


@pytest.mark.parametrize("kind", ["gbm", "poisson_ou"])
def test_sde_is_correction(engine, kind):
    """correction phase of run_mcmc.ssm_sde: fine-mesh bootstrap filter per stored draw
    (is_correction_bsf<ssm_sde>, src/approx_mcmc.cpp:522-591), weight = exp(loglik_f - approx_loglik)"""
    model = data.gbm_model(30) if kind == "gbm" else data.poisson_ou_model(30)
    rng = np.random.default_rng(3)
    ns = 20
    th0 = np.asarray(model.theta)
    theta = th0[None, :] + 0.05 * rng.normal(size=(ns, 3))
    if kind == "gbm":
        theta[:, 1:] = np.abs(theta[:, 1:])
    counts = (1 + rng.poisson(2, size=ns)).astype(np.uint32)
    # coarse-mesh estimates as the MCMC phase would have stored them (any numbers do for parity)
    approx = np.array([orc.sde_bsf_filter(engine.ssm_sde(model.y, kind, theta[i], model.x0, model.positive),
                                          64, 1, seed=100 + i).loglik for i in range(ns)])
    mo = {"theta": theta, "counts": counts, "approx_loglik": approx, "output_type": 3, "posterior": np.zeros(ns)}
    ref = orc.is_correction_bsf_sde(model, mo, 128, 3, is_type=2, seed=4)
    res = engine.post_correct(model, mo, 128, is_type="is2", seed=4, return_loglik=True, L=3)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    np.testing.assert_allclose(res["posterior"], ref.posterior, rtol=1e-13, atol=1e-13)
    assert np.all(np.abs(np.log(res["weights"])) < 5)  # fine vs coarse mesh estimates of the same likelihood
    mo1 = dict(mo, output_type=1)
    ref1 = orc.is_correction_bsf_sde(model, mo1, 64, 2, is_type=3, seed=6)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=6, L=2)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.is_correction_bsf_sde(model, mo2, 32, 2, is_type=1, seed=8)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=8, L=2)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-9, atol=1e-13)


def test_bsf_is_correction(setup):
    """correction phase of run_mcmc(sampling_method = "bsf"): bootstrap filter per stored draw,
    weight = exp(loglik_bsf - approx_loglik) (is_correction_bsf, src/approx_mcmc.cpp:447-518)"""
    engine, model, mo = setup
    rng = np.random.default_rng(5)
    mo = dict(mo, approx_loglik=-221.0 + rng.normal(size=mo["counts"].size))
    ref = orc.is_correction_bsf(model, mo, 128, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 128, is_type="is2", seed=3, return_loglik=True, sampling_method="bsf")
    np.testing.assert_allclose(res["loglik"], ref.loglik, rtol=1e-13)  # constants pre-summed per draw
    np.testing.assert_allclose(res["weights"], ref.weights, rtol=1e-10)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.is_correction_bsf(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5, sampling_method="bsf")
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.is_correction_bsf(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7, sampling_method="bsf")
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


def test_ar_exp_post_correction_and_suggest_N(engine):
    """postcorrection_nonlinear / suggest_n_nonlinear for the one-state registry model: (mu, rho) of the
    true transition differ per draw, so they travel with the per-draw approximation"""
    model = data.ar_exp_model(50, missing=(3,))
    mo = data.growth_mcmc_output(model, 16, jitter=0.04)
    ref = orc.post_correct(model, mo, 96, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 96, is_type="is2", seed=3, return_loglik=True)
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    mo1 = dict(mo, output_type=1)
    ref1 = orc.post_correct(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5)
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.post_correct(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7)
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)
    sn = engine.suggest_N(model, np.asarray(model.theta), candidates=[20, 400], replications=12, seed=2)
    assert np.all(np.isfinite(sn["results"]["sd"])) and sn["results"]["sd"][1] < sn["results"]["sd"][0]


def test_summary_with_a_draw_whose_filter_stops(engine):
    """a stored draw whose filter stops at once (sigma so small that (y / sigma)^2 overflows)
    has weight 0; its unreached columns must read as zeros (the reference's zero-filled buffers), not as
    whatever the recycled device blocks held, so that the merged summary stays finite and equals the oracle's"""
    model = data.svm_model(80)
    mo = data.svm_mcmc_output(model, 12, output_type=2)
    mo["theta"] = np.array(mo["theta"], copy=True)
    mo["theta"][5, 2] = 1e-200
    # fill the pool with non-zero garbage of the right size classes first
    engine.post_correct(model, dict(data.svm_mcmc_output(model, 12, output_type=2)), 48, is_type="is2", seed=2)
    ref = orc.post_correct(model, mo, 48, is_type=2, seed=4)
    res = engine.post_correct(model, mo, 48, is_type="is2", seed=4, return_loglik=True)
    assert res["loglik"][5] == -np.inf and res["weights"][5] == 0.0
    assert np.array_equal(res["weights"], ref.weights)
    assert np.all(np.isfinite(res["alphahat"])) and np.all(np.isfinite(res["Vt"]))
    np.testing.assert_allclose(res["alphahat"], ref.alphahat, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res["Vt"], ref.Vt, rtol=1e-9, atol=1e-13)


def test_post_correct_inside_the_library_on_two_gpus(engine):
    """ONE call of the C ABI with n_threads = 2 shards the draws over two GPUs (one host
    thread per device, contiguous static blocks like the reference's schedule(static),
    src/approx_mcmc.cpp:364-369), gathers the per-draw log-likelihoods with one ncclAllGather over a
    persistent communicator, and returns exactly what the one-GPU call returns"""
    import bssm_b200
    if engine.lib().bssm_b200_device_count() < 2:
        pytest.skip("needs two visible GPUs (gpurun --gpus 2)")
    model = data.svm_model(200)
    mo = data.svm_mcmc_output(model, 301)
    one = engine.post_correct(model, mo, 128, threads=1, is_type="is2", seed=5, return_loglik=True)
    two = engine.post_correct(model, mo, 128, threads=2, is_type="is2", seed=5, return_loglik=True)
    note = engine.lib().bssm_b200_last_gather().decode()
    assert note.startswith("nccl: 2 ranks"), note
    assert np.array_equal(one["weights"], two["weights"]) and np.array_equal(one["loglik"], two["loglik"])
    assert np.array_equal(one["posterior"], two["posterior"])
    ref = orc.post_correct(model, mo, 128, is_type=2, seed=5)
    assert np.array_equal(two["weights"], ref.weights)
    # trajectories (output_type 1): every device writes its slice of the caller's buffer
    mo1 = dict(mo, output_type=1)
    a1 = engine.post_correct(model, mo1, 64, threads=1, is_type="is3", seed=6)
    a2 = engine.post_correct(model, mo1, 64, threads=2, is_type="is3", seed=6)
    assert np.array_equal(a1["alpha"].view(np.uint64), a2["alpha"].view(np.uint64))
    assert np.array_equal(a1["weights"], a2["weights"])


@pytest.mark.parametrize("kind", ["growth", "ar_exp", "sin_exp", "linear_gaussian"])
def test_bsf_is_correction_nlg(engine, kind):
    """correction phase of run_mcmc(sampling_method = "bsf") for ssm_nlg: update_model(theta_i), bootstrap
    filter, weight = exp(loglik_bsf - approx_loglik_i) (is_correction_bsf<ssm_nlg>, src/approx_mcmc.cpp:442-518)"""
    model = {"growth": lambda: data.growth_model(40, missing=(5,)), "ar_exp": lambda: data.ar_exp_model(50),
             "sin_exp": lambda: data.sin_exp_model(40), "linear_gaussian": lambda: data.linear_gaussian_model(40)}[kind]()
    rng = np.random.default_rng(13)
    ns = 14
    th0 = np.asarray(model.theta, dtype=np.float64)
    theta = th0[None, :] + 0.04 * rng.normal(size=(ns, th0.size))
    if kind == "ar_exp":
        theta[:, 1] = np.clip(theta[:, 1], -0.98, 0.98)
    counts = (1 + rng.poisson(2, size=ns)).astype(np.uint32)
    base = orc.nlg_bsf_filter(model, 256, seed=77).loglik
    approx = base + 0.5 * rng.normal(size=ns)
    mo = {"theta": theta, "counts": counts, "approx_loglik": approx, "output_type": 3, "posterior": np.zeros(ns)}
    ref = orc.is_correction_bsf_nlg(model, mo, 128, is_type=2, seed=3)
    res = engine.post_correct(model, mo, 128, is_type="is2", seed=3, return_loglik=True, sampling_method="bsf")
    assert np.array_equal(res["loglik"], ref.loglik)
    assert np.array_equal(res["weights"], ref.weights)
    np.testing.assert_allclose(res["posterior"], ref.posterior, rtol=1e-13, atol=1e-13)
    assert len(set(res["loglik"])) == ns  # every draw ran at its own theta
    mo1 = dict(mo, output_type=1)
    ref1 = orc.is_correction_bsf_nlg(model, mo1, 64, is_type=3, seed=5)
    res1 = engine.post_correct(model, mo1, 64, is_type="is3", seed=5, sampling_method="bsf")
    assert np.array_equal(res1["alpha"].view(np.uint64), ref1.alpha.view(np.uint64))
    mo2 = dict(mo, output_type=2)
    ref2 = orc.is_correction_bsf_nlg(model, mo2, 32, is_type=1, seed=7)
    res2 = engine.post_correct(model, mo2, 32, is_type="is1", seed=7, sampling_method="bsf")
    np.testing.assert_allclose(res2["alphahat"], ref2.alphahat, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res2["Vt"], ref2.Vt, rtol=1e-8, atol=1e-12)


def test_device_approximation_equals_host_approximation(engine, monkeypatch):
    """per-draw Gaussian approximations for m > 1 (approximate_for_is: laplace_iter, scales, smoother_ccov,
    conditional_cov by SVD) run on the device, one thread per draw (approx_device.cuh); BSSM_B200_HOST_APPROX=1
    sends the same call through the host's OpenMP loop over include/bssm_b200/ung_gauss_approx.h.  Same bits."""
    for kind in ("level_slope", "negbin_xreg", "noise"):
        model = _bsm_case(kind)
        mo = dict(data.bsm_mcmc_output(model, 24), output_type=2)
        dev = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
        monkeypatch.setenv("BSSM_B200_HOST_APPROX", "1")
        host = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
        monkeypatch.delenv("BSSM_B200_HOST_APPROX")
        assert np.array_equal(dev["loglik"], host["loglik"]), kind
        assert np.array_equal(dev["alphahat"], host["alphahat"]) and np.array_equal(dev["Vt"], host["Vt"]), kind


@pytest.mark.parametrize("kind", ["growth", "ar_exp", "sin_exp"])
def test_device_approximation_equals_host_approximation_nlg(engine, kind, monkeypatch):
    """postcorrection_nonlinear: approximate_for_is per draw (relinearisation at the stored mode, smoother,
    transition-density factors) on the device vs the host loop over include/bssm_b200/nlg_gauss_approx.h"""
    model = {"growth": lambda: data.growth_model(50, missing=(4,)), "ar_exp": lambda: data.ar_exp_model(50, missing=(3,)),
             "sin_exp": lambda: data.sin_exp_model(40)}[kind]()
    mo = dict(data.growth_mcmc_output(model, 12, jitter=0.04), output_type=2)
    dev = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    monkeypatch.setenv("BSSM_B200_HOST_APPROX", "1")
    host = engine.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
    monkeypatch.delenv("BSSM_B200_HOST_APPROX")
    assert np.array_equal(dev["loglik"], host["loglik"])
    assert np.array_equal(dev["alphahat"], host["alphahat"]) and np.array_equal(dev["Vt"], host["Vt"])
    ref = orc.post_correct(model, dict(mo, output_type=3), 64, is_type=2, seed=3)
    assert np.array_equal(dev["loglik"], ref.loglik)
