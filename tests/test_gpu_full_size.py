"""BASELINE.json configurations 2-5 at their FULL sizes.

Two kinds of checks: (i) the whole run compared with the CPU oracle BIT FOR BIT (ancestor indices,
weights, log-likelihood) -- C2 and C3 at the sizes BASELINE names, C5 at its full particle count on a
shortened series (the scalar oracle needs 16 Milstein sub-steps per particle-step); (ii)
size-independent properties: sorted ancestors, weight normalisation, same-seed determinism, sharding
invariance, agreement of the log-likelihood with the closed form where one exists."""
import numpy as np
import pytest

import bssm_b200
import data
import orc

pytestmark = pytest.mark.gpu


def test_full_size_c3_growth_bsf_and_smoother(engine):
    """ssm_nlg growth model, bootstrap filter + particle smoother, N = 2^20, T = 200"""
    model = data.growth_model(200)
    N = 1 << 20
    job = engine.FilterJob(model, N, method="bsf", seed=1, out_flags=bssm_b200.api.OUT_INDICES)
    job.run()
    ll = job.logliks()[0]
    b = job.fetch(0, indices=True)
    idx = b.indices
    assert b.out.stopped_at == -1 and np.isfinite(ll)
    assert idx.max() < N and np.all(np.diff(idx.astype(np.int64), axis=0) >= 0), "ancestors sorted per column"
    job.run()
    assert job.logliks()[0] == ll, "same seed, same bits"
    job.close()
    ref = np.mean([orc.nlg_bsf_filter(model, 4096, seed=s).loglik for s in (1, 2, 3)])
    assert abs(ll - ref) < 0.3
    # smoother summary at full size against a 2^14-particle run of the same engine (Monte Carlo error only)
    sm = engine.FilterJob(model, N, method="bsf", seed=2, post=bssm_b200.api.POST_SMOOTHER)
    sm.run()
    big = sm.fetch(0, smooth=True)
    sm.close()
    small = engine.particle_smoother(model, 1 << 14, method="bsf", seed=3)
    np.testing.assert_allclose(big.alphahat.T[:, 1], small["alphahat"][:, 1], rtol=0.02)  # population level
    assert np.all(np.isfinite(big.Vt))


def test_full_size_c5_gbm(engine):
    """ssm_sde GBM, Milstein with 2^4 sub-steps, N = 2^18, T = 500: exact likelihood available"""
    from test_reference_golden import gbm_exact_loglik
    model = data.gbm_model(500)
    N = 1 << 18
    job = engine.FilterJob(model, N, method="bsf", seed=1, out_flags=bssm_b200.api.OUT_INDICES | bssm_b200.api.OUT_WEIGHTS,
                           L=4)
    job.run()
    ll = job.logliks()[0]
    b = job.fetch(0, weights=True, indices=True)
    assert b.out.stopped_at == -1
    assert np.all(np.diff(b.indices.astype(np.int64), axis=0) >= 0)
    assert np.all(b.weights.max(axis=0) == 1.0) and b.weights.min() >= 0
    job.run()
    assert job.logliks()[0] == ll
    job.close()
    assert abs(ll - gbm_exact_loglik(model)) < 0.5  # Milstein bias at L = 4 plus Monte Carlo error
    # stratified resampling recomputed from the stored weights at full size
    for t in (0, 250, 498):
        r = orc.resample_uniforms(1, 0, t, N)
        assert np.array_equal(b.indices[:, t], orc.stratified_sample(b.weights[:, t], r, repro=True)), t


def test_full_size_c4_post_correct_sharding(engine):
    """post_correct on svm, 15 000 stored draws x psi-APF with N = 256, T = 1000: eight shards (the 8-GPU
    partition) concatenate to the single-call result bit for bit; a sample of draws equals the oracle"""
    model = data.svm_model(1000)
    mo = bssm_b200.synthetic_svm_mcmc_output(model, 15000)
    whole = engine.post_correct(model, mo, 256, is_type="is2", seed=1, return_loglik=True)
    assert whole["weights"].size == 15000 and np.all(np.isfinite(whole["loglik"])) and np.all(whole["weights"] > 0)
    from bssm_b200.distributed import shard_range
    parts = []
    for rank in range(8):
        b, e = shard_range(15000, rank, 8)
        parts.append(engine.post_correct(model, mo, 256, is_type="is2", seed=1, draw_range=(b, e))["weights"])
    assert np.array_equal(np.concatenate(parts), whole["weights"])
    # oracle on 24 of the draws (each draw d uses stream d, so a subset needs the same draw indices)
    sub = {k: (v[:24] if k in ("theta", "counts", "posterior") else v) for k, v in mo.items()}
    sub["modes"] = mo["modes"][:, :24]
    ref = orc.post_correct(model, sub, 256, is_type=2, seed=1)
    assert np.array_equal(whole["loglik"][:24], ref.loglik)


def test_c2_full_run_is_bit_identical_to_the_oracle(engine):
    """BASELINE config 2 exactly: svm psi-APF, N = 65536, T = 1000 -- every ancestor index, every weight,
    every particle and the log-likelihood equal the oracle's (6.5e7 particle-steps, ~13 s of oracle time)"""
    model = data.svm_model(1000)
    N = 65536
    ref = orc.psi_filter(model, N, seed=1)
    job = engine.FilterJob(model, N, method="psi", seed=1, out_flags=7)
    job.run()
    b = job.fetch(0, alpha=True, weights=True, indices=True)
    job.close()
    assert b.out.stopped_at == -1 == ref.stopped_at
    assert np.array_equal(b.indices, ref.indices), "ancestor indices must be bit-exact at the headline size"
    assert np.array_equal(b.weights.view(np.uint64), ref.weights.view(np.uint64))
    assert np.array_equal(b.alpha.view(np.uint64), ref.alpha.view(np.uint64))
    assert b.out.loglik == ref.loglik
    assert abs(b.out.loglik - ref.loglik) <= 1e-6 * abs(ref.loglik)  # the north-star tolerance, met with margin zero


def test_c3_full_run_is_bit_identical_to_the_oracle(engine):
    """BASELINE config 3: ssm_nlg growth model, bootstrap filter, N = 2^20, T = 200 (2.1e8 particle-steps,
    about a minute of oracle time): ancestors, weights and log-likelihood bit for bit"""
    model = data.growth_model(200)
    N = 1 << 20
    ref = orc.nlg_bsf_filter(model, N, seed=1)
    job = engine.FilterJob(model, N, method="bsf", seed=1,
                           out_flags=bssm_b200.api.OUT_INDICES | bssm_b200.api.OUT_WEIGHTS)
    job.run()
    b = job.fetch(0, weights=True, indices=True)
    ll = job.logliks()[0]
    job.close()
    assert np.array_equal(b.indices, ref.indices)
    assert np.array_equal(b.weights.view(np.uint64), ref.weights.view(np.uint64))
    assert ll == ref.loglik


def test_c5_full_particle_count_is_bit_identical_to_the_oracle(engine):
    """BASELINE config 5 at its full particle count: ssm_sde GBM, Milstein L = 4, N = 2^18; the series is
    cut to T = 100 (4.2e8 Milstein sub-steps, ~30 s of oracle time; the full T = 500 is covered by the
    properties above)"""
    model = data.gbm_model(100)
    N = 1 << 18
    ref = orc.sde_bsf_filter(model, N, 4, seed=1)
    job = engine.FilterJob(model, N, method="bsf", seed=1,
                           out_flags=bssm_b200.api.OUT_INDICES | bssm_b200.api.OUT_WEIGHTS, L=4)
    job.run()
    b = job.fetch(0, weights=True, indices=True)
    ll = job.logliks()[0]
    job.close()
    assert np.array_equal(b.indices, ref.indices)
    assert np.array_equal(b.weights.view(np.uint64), ref.weights.view(np.uint64))
    assert ll == ref.loglik
