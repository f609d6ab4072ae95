"""ssm_sde models given by their four functions (drift, diffusion, ddiffusion, log_obs_density): the
reference takes compiled function pointers (src/model_ssm_sde.h:8-11), the engine takes CUDA source and
compiles it at run time into its own filter kernels (bssm_b200/csrc/user_models.cu).  Parity: (1) the
shipped GBM example written as user source gives the registry model's bits, (2) a model the engine has
never seen equals the oracle running the SAME source compiled with g++ through function pointers."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import data
import orc

pytestmark = pytest.mark.gpu

GBM_SRC = r"""
// sde_gbm of R/cpp_example_models.R:109-193: theta = (mu, sigma_x, sigma_y)
__device__ double drift(double x, const double* theta) { return theta[0] * x; }
__device__ double diffusion(double x, const double* theta) { const double tx = theta[1] * x; return tx > 0.0 ? tx : 0.0; }
__device__ double ddiffusion(double x, const double* theta) { return theta[1] * (x > 0.0 ? 1.0 : 0.0); }
__device__ double log_obs_density(double y, double x, const double* theta) {
  const double q = (y - pf_log(x)) / theta[2];
  return -(0.918938533204672741780329736406 + 0.5 * q * q + pf_log(theta[2]));
}
"""

# Cox-Ingersoll-Ross state observed with Gaussian noise: theta = (log kappa, mu, log sigma, log sd_y, unused)
CIR_SRC = r"""
__device__ double drift(double x, const double* theta) { return pf_exp(theta[0]) * (theta[1] - x); }
__device__ double diffusion(double x, const double* theta) {
  const double s = pf_exp(theta[2]);
  return x > 0.0 ? s * pf_exp(0.5 * pf_log(x)) : 0.0;
}
__device__ double ddiffusion(double x, const double* theta) {
  const double s = pf_exp(theta[2]);
  return x > 0.0 ? 0.5 * s * pf_exp(-0.5 * pf_log(x)) : 0.0;
}
__device__ double log_obs_density(double y, double x, const double* theta) {
  const double q = (y - x) * pf_exp(-theta[3]);
  return -(0.918938533204672741780329736406 + 0.5 * q * q + theta[3]);
}
"""


def host_functions(src):
    """the same source text compiled for the host: addresses of the four functions for the oracle"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = tempfile.mkdtemp(prefix="bssm_user_")
    cpp = os.path.join(d, "user.cpp")
    with open(cpp, "w") as f:
        f.write('#include "bssm_b200/pf_math.h"\n#define __device__\nextern "C" {\n' + src + "\n}\n")
    so = os.path.join(d, "user.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I" + os.path.join(root, "include"), cpp, "-o", so],
                   check=True)
    lib = C.CDLL(so)
    return lib, [C.cast(getattr(lib, nm), C.c_void_p) for nm in ("drift", "diffusion", "ddiffusion", "log_obs_density")]


def test_user_gbm_equals_registry_gbm(engine):
    reg = data.gbm_model(40)
    mid = engine.sde_register(GBM_SRC)
    assert mid >= 1000 and engine.sde_register(GBM_SRC) == mid  # same source, same id
    user = engine.ssm_sde(reg.y, mid, reg.theta, reg.x0, reg.positive)
    for N, L in ((100, 3), (3000, 2), (9000, 2)):  # latency kernel, one-CTA kernel, group of CTAs
        a = engine.bootstrap_filter(reg, N, seed=2, L=L, return_indices=True)
        b = engine.bootstrap_filter(user, N, seed=2, L=L, return_indices=True)
        for key in ("alpha", "weights", "indices", "at", "att", "Pt", "Ptt"):
            assert np.array_equal(a[key], b[key]), (N, key)
        assert a["logLik"] == b["logLik"]
    sa = engine.particle_smoother(reg, 500, method="bsf", seed=4, L=2)
    sb = engine.particle_smoother(user, 500, method="bsf", seed=4, L=2)
    assert np.array_equal(sa["alphahat"], sb["alphahat"]) and np.array_equal(sa["Vt"], sb["Vt"])
    # correction phase of run_mcmc.ssm_sde: one theta per filter
    rng = np.random.default_rng(3)
    ns = 12
    theta = np.abs(np.asarray(reg.theta)[None, :] + 0.05 * rng.normal(size=(ns, 3)))
    mo = {"theta": theta, "counts": (1 + rng.poisson(2, size=ns)).astype(np.uint32),
          "approx_loglik": -30.0 + rng.normal(size=ns), "output_type": 1, "posterior": np.zeros(ns)}
    ra = engine.post_correct(reg, mo, 64, is_type="is2", seed=5, L=3, return_loglik=True)
    rb = engine.post_correct(user, mo, 64, is_type="is2", seed=5, L=3, return_loglik=True)
    assert np.array_equal(ra["loglik"], rb["loglik"]) and np.array_equal(ra["alpha"], rb["alpha"])


def test_user_model_against_the_oracle_running_the_same_source(engine):
    rng = np.random.default_rng(8)
    n = 40
    x, y = 1.2, np.zeros(n)
    for t in range(n):
        for _ in range(8):
            x = abs(x + 0.8 * (1.0 - x) / 8 + 0.3 * np.sqrt(max(x, 0.0)) * rng.normal() / np.sqrt(8))
        y[t] = x + 0.2 * rng.normal()
    y[5] = np.nan
    theta = np.array([np.log(0.8), 1.0, np.log(0.3), np.log(0.2), 0.0])
    mid = engine.sde_register(CIR_SRC)
    model = engine.ssm_sde(y, mid, theta, 1.2, True)
    lib, fns = host_functions(CIR_SRC)
    f = orc.L().orc_sde_set_user
    f.restype = None
    f.argtypes = [C.c_void_p] * 4
    f(*fns)
    for N, L in ((64, 3), (2000, 2)):
        ref = orc.sde_bsf_filter(model, N, L, seed=7)
        res = engine.bootstrap_filter(model, N, seed=7, L=L, return_indices=True)
        assert np.array_equal(res["indices"], ref.indices)
        # the engine returns alpha as (n + 1) x m x N (R's layout), the oracle holds m x (n + 1) x N
        assert np.array_equal(np.transpose(res["alpha"], (1, 0, 2)).view(np.uint64), ref.alpha.view(np.uint64))
        assert res["logLik"] == ref.loglik
    # a resident job inside an MCMC loop: new theta per iteration
    job = engine.FilterJob(model, 128, method="bsf", seed=3, out_flags=0, L=3)
    th2 = theta + [0.1, 0.05, -0.1, 0.0, 0.0]
    job.set_theta(th2)
    job.run()
    m2 = engine.ssm_sde(y, mid, th2, 1.2, True)
    assert job.logliks()[0] == orc.sde_bsf_filter(m2, 128, 3, seed=3).loglik
    job.close()


def test_user_source_that_does_not_compile(engine):
    with pytest.raises(Exception, match="does not compile"):
        engine.sde_register("__device__ double drift(double x, const double* theta) { return undefined_symbol; }")


GROWTH_SRC = r"""
// nlg_growth of R/cpp_example_models.R:336-511: state (logit r, p), theta = (log H, log R1, log R2),
// known_params = (dT, K, a11, a12, P11, P12)
__device__ double Z_fn(int t, const double* a, const double* th, const double* kp) { return a[1]; }
__device__ double H_fn(int t, const double* a, const double* th, const double* kp) { return pf_exp(th[0]); }
__device__ void T_fn(int t, const double* a, const double* th, const double* kp, double* o) {
  const double e0 = pf_exp(a[0]);
  const double r = e0 / (1.0 + e0);
  const double erdt = pf_exp(r * kp[0]);
  o[0] = a[0];
  o[1] = kp[1] * a[1] * erdt / (kp[1] + a[1] * (erdt - 1));
}
__device__ void R_fn(int t, const double* a, const double* th, const double* kp, double* o) {
  o[0] = pf_exp(th[1]); o[1] = 0.0; o[2] = 0.0; o[3] = pf_exp(th[2]);
}
__device__ void a1_fn(const double* th, const double* kp, double* o) { o[0] = kp[2]; o[1] = kp[3]; }
__device__ void P1_fn(const double* th, const double* kp, double* o) { o[0] = kp[4]; o[1] = 0.0; o[2] = 0.0; o[3] = kp[5]; }
"""

SIN_EXP_SRC = r"""
// nlg_sin_exp of R/cpp_example_models.R:635-757: theta = (log H, log R), a1 = 0, P1 = 1
__device__ double Z_fn(int t, const double* a, const double* th, const double* kp) { return pf_exp(a[0]); }
__device__ double H_fn(int t, const double* a, const double* th, const double* kp) { return pf_exp(th[0]); }
__device__ void T_fn(int t, const double* a, const double* th, const double* kp, double* o) { o[0] = pf_sin(a[0]); }
__device__ void R_fn(int t, const double* a, const double* th, const double* kp, double* o) { o[0] = pf_exp(th[1]); }
__device__ void a1_fn(const double* th, const double* kp, double* o) { o[0] = 0.0; }
__device__ void P1_fn(const double* th, const double* kp, double* o) { o[0] = 1.0; }
"""


@pytest.mark.parametrize("kind", ["growth", "sin_exp"])
def test_user_nlg_equals_registry_model(engine, kind):
    """the shipped ssm_nlg examples written as user source: bootstrap filter, smoother, MCMC-style set_theta and the
    bootstrap correction phase give the registry policies' bits (which the other tests pin to the oracle)"""
    if kind == "growth":
        reg, src, m, k = data.growth_model(50, missing=(4,)), GROWTH_SRC, 2, 2
    else:
        reg, src, m, k = data.sin_exp_model(50), SIN_EXP_SRC, 1, 1
    mid = engine.nlg_register(src, m, k)
    user = engine.ssm_nlg(reg.y, mid, reg.theta, reg.known_params, n_states=m, n_etas=k)
    for N in (100, 3000, 9000):  # latency kernel, one-CTA kernel, group of CTAs
        a = engine.bootstrap_filter(reg, N, seed=2, return_indices=True)
        b = engine.bootstrap_filter(user, N, seed=2, return_indices=True)
        for key in ("alpha", "weights", "indices", "at", "att", "Pt", "Ptt"):
            assert np.array_equal(a[key], b[key]), (N, key)
        assert a["logLik"] == b["logLik"]
    sa = engine.particle_smoother(reg, 500, method="bsf", seed=4)
    sb = engine.particle_smoother(user, 500, method="bsf", seed=4)
    assert np.array_equal(sa["alphahat"], sb["alphahat"]) and np.array_equal(sa["Vt"], sb["Vt"])
    th2 = np.asarray(reg.theta) + 0.05
    ja = engine.FilterJob(reg, 128, method="bsf", seed=3, out_flags=0)
    jb = engine.FilterJob(user, 128, method="bsf", seed=3, out_flags=0)
    for job in (ja, jb):
        job.set_theta(th2)
        job.run()
    assert ja.logliks()[0] == jb.logliks()[0]
    ja.close(); jb.close()
    rng = np.random.default_rng(3)
    ns = 10
    theta = np.asarray(reg.theta)[None, :] + 0.04 * rng.normal(size=(ns, np.size(reg.theta)))
    mo = {"theta": theta, "counts": (1 + rng.poisson(2, size=ns)).astype(np.uint32),
          "approx_loglik": a["logLik"] + rng.normal(size=ns), "output_type": 1, "posterior": np.zeros(ns)}
    ra = engine.post_correct(reg, mo, 64, is_type="is2", seed=5, return_loglik=True, sampling_method="bsf")
    rb = engine.post_correct(user, mo, 64, is_type="is2", seed=5, return_loglik=True, sampling_method="bsf")
    assert np.array_equal(ra["loglik"], rb["loglik"]) and np.array_equal(ra["alpha"], rb["alpha"])
    with pytest.raises(Exception, match="bootstrap filter"):
        engine.particle_smoother(user, 100, method="psi", seed=1)


def test_new_correction_paths_on_two_gpus(engine):
    """parameter sets, the ssm_nlg bootstrap correction and a user SDE model through ONE C-ABI call with
    n_threads = 2: shards on two devices (the run-time compiled kernels load on both), same results"""
    if engine.lib().bssm_b200_device_count() < 2:
        pytest.skip("needs two visible GPUs (gpurun --gpus 2)")
    import test_gpu_parset as tp
    model = tp.ung_update_model(m=2, tv=True)
    mo = tp.draws(model, 40)
    one = engine.post_correct(model, mo, 64, threads=1, is_type="is2", seed=3, return_loglik=True)
    two = engine.post_correct(model, mo, 64, threads=2, is_type="is2", seed=3, return_loglik=True)
    assert engine.lib().bssm_b200_last_gather().decode().startswith("nccl: 2 ranks")
    assert np.array_equal(one["loglik"], two["loglik"]) and np.array_equal(one["weights"], two["weights"])
    reg = data.gbm_model(40)
    user = engine.ssm_sde(reg.y, engine.sde_register(GBM_SRC), reg.theta, reg.x0, reg.positive)
    rng = np.random.default_rng(3)
    ns = 30
    theta = np.abs(np.asarray(reg.theta)[None, :] + 0.05 * rng.normal(size=(ns, 3)))
    mos = {"theta": theta, "counts": (1 + rng.poisson(2, size=ns)).astype(np.uint32),
           "approx_loglik": -30.0 + rng.normal(size=ns), "output_type": 1, "posterior": np.zeros(ns)}
    ra = engine.post_correct(reg, mos, 64, threads=2, is_type="is2", seed=5, L=3, return_loglik=True)
    rb = engine.post_correct(user, mos, 64, threads=2, is_type="is2", seed=5, L=3, return_loglik=True)
    assert np.array_equal(ra["loglik"], rb["loglik"]) and np.array_equal(ra["alpha"], rb["alpha"])
    g = data.growth_model(40)
    thg = np.asarray(g.theta)[None, :] + 0.04 * rng.normal(size=(ns, 3))
    mog = {"theta": thg, "counts": mos["counts"], "approx_loglik": -100.0 + rng.normal(size=ns), "output_type": 3,
           "posterior": np.zeros(ns)}
    g1 = engine.post_correct(g, mog, 64, threads=1, is_type="is2", seed=5, return_loglik=True, sampling_method="bsf")
    g2 = engine.post_correct(g, mog, 64, threads=2, is_type="is2", seed=5, return_loglik=True, sampling_method="bsf")
    assert np.array_equal(g1["loglik"], g2["loglik"])


def test_user_models_survive_shutdown(engine):
    """bssm_b200_shutdown unloads the run-time compiled kernels; registered ids stay valid (compiled again on use)"""
    reg = data.gbm_model(30)
    mid = engine.sde_register(GBM_SRC)
    user = engine.ssm_sde(reg.y, mid, reg.theta, reg.x0, reg.positive)
    a = engine.bootstrap_filter(user, 100, seed=2, L=2)["logLik"]
    engine.lib().bssm_b200_shutdown()
    b = engine.bootstrap_filter(user, 100, seed=2, L=2)["logLik"]
    assert a == b == engine.bootstrap_filter(reg, 100, seed=2, L=2)["logLik"]
