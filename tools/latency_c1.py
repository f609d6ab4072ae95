"""C1 (BASELINE configs[0]): bootstrap_filter(bsm_ng Poisson, n=100, m=2, N=100) -- call latency of the
one-shot C-ABI path and of a resident job, next to the CPU oracle port."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bssm_b200, data, orc

model = data.poisson_bsm_model(100)
for N in (100, 1000, 10000):
    for _ in range(3):
        bssm_b200.bootstrap_filter(model, N, seed=1)
    t0 = time.perf_counter()
    for i in range(20):
        r = bssm_b200.bootstrap_filter(model, N, seed=1 + i)
    one = (time.perf_counter() - t0) / 20 * 1e3
    job = bssm_b200.FilterJob(model, N, method="bsf", seed=1, out_flags=0)
    for _ in range(3):
        job.run()
    t0 = time.perf_counter()
    for i in range(20):
        ms = job.run()
    res = (time.perf_counter() - t0) / 20 * 1e3
    dev = job.kernel_ms()
    job.close()
    for _ in range(2):
        orc.bsf_filter(model, N, seed=1)
    t0 = time.perf_counter()
    for i in range(5):
        orc.bsf_filter(model, N, seed=1 + i)
    cpu = (time.perf_counter() - t0) / 5 * 1e3
    print("N=%6d: one-shot call %.3f ms | resident job run %.3f ms (device %.3f ms) | CPU port %.3f ms" % (N, one, res, sum(dev), cpu))
for N in (100, 1000):
    mdl = data.svm_model(1000)
    job = bssm_b200.FilterJob(mdl, N, method="psi", seed=1, out_flags=0)
    for _ in range(3):
        job.run()
    t0 = time.perf_counter()
    for i in range(20):
        job.run()
    res = (time.perf_counter() - t0) / 20 * 1e3
    print("svm psi n=1000 N=%5d: resident job run %.3f ms (device %s)" % (N, res, ["%.3f" % v for v in job.kernel_ms()]))
    job.close()
# the filter inside an MCMC loop: set_theta + run per iteration on a resident job
mdl = data.svm_model(1000)
for N in (100, 1000):
    job = bssm_b200.FilterJob(mdl, N, method="psi", seed=1, out_flags=0)
    rng = np.random.default_rng(1)
    th0 = np.asarray(mdl.theta)
    for _ in range(3):
        job.set_theta(th0); job.run()
    t0 = time.perf_counter()
    for i in range(20):
        job.set_theta(th0 * np.exp(0.02 * rng.normal(size=3)) * [1, 1, 1]); job.run()
    it = (time.perf_counter() - t0) / 20 * 1e3
    print("svm psi n=1000 N=%5d: MCMC-style iteration (set_theta + run) %.3f ms" % (N, it))
    job.close()
# the same loop for the ssm_nlg / ssm_sde registry models (nonlinear_pm_mcmc / sde_pm_mcmc)
for name, mdl, method, kw in (("growth bsf n=200", data.growth_model(200), "bsf", {}),
                              ("growth ekf n=200", data.growth_model(200), "ekf", {}),
                              ("growth psi n=200", data.growth_model(200), "psi", {}),
                              ("gbm bsf n=50 L=4", data.gbm_model(50), "bsf", {"L": 4})):
    for N in (100, 1000):
        job = bssm_b200.FilterJob(mdl, N, method=method, seed=1, out_flags=0, **kw)
        rng = np.random.default_rng(1)
        th0 = np.asarray(mdl.theta)
        for _ in range(3):
            job.set_theta(th0); job.run()
        t0 = time.perf_counter()
        for i in range(20):
            job.set_theta(th0 + 0.01 * rng.normal(size=th0.size) * (mdl.cls != "ssm_sde")); job.run()
        it = (time.perf_counter() - t0) / 20 * 1e3
        print("%s N=%5d: MCMC-style iteration (set_theta + run) %.3f ms" % (name, N, it))
        job.close()
