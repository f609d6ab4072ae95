"""Time of importance_sample() / logLik(method = "spdk") on the device against the oracle on one host core.
Run on the GPU box: python tools/spdk_time.py > gpurun_out/spdk_time.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200  # noqa: E402
import data  # noqa: E402
import orc  # noqa: E402

orc.build()


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


rows = []
for name, model, nsim in (("svm n=1000 (m=1)", data.svm_model(n=1000), 65536),
                          ("svm n=1000 (m=1)", data.svm_model(n=1000), 1000),
                          ("bsm_ng poisson n=100 (m=2)", data.poisson_bsm_model(), 65536),
                          ("bsm_ng poisson n=100 (m=2)", data.poisson_bsm_model(), 100)):
    bssm_b200.logLik(model, nsim, method="spdk", seed=1)  # warm-up
    t_ll = best(lambda: bssm_b200.logLik(model, nsim, method="spdk", seed=1))
    t_all = best(lambda: bssm_b200.importance_sample(model, nsim, seed=1), reps=3)
    n_cpu = min(nsim, 4096)
    t_cpu = best(lambda: orc.importance_sample(model, n_cpu, 1), reps=2) * nsim / n_cpu
    ll = bssm_b200.logLik(model, nsim, method="spdk", seed=1)
    ref = orc.importance_sample(model, n_cpu, 1)[2][0] if n_cpu == nsim else float("nan")
    rows.append((name, nsim, t_ll * 1e3, t_all * 1e3, t_cpu * 1e3, ll, ref))
print("SPDK (simulation smoother + importance weights), wall time per call, best of 5; oracle on ONE host core")
print("(scaled from %d draws where nsim is larger)" % 4096)
print("%-30s %8s %14s %18s %14s %16s" % ("model", "nsim", "logLik ms", "with alpha D2H ms", "oracle ms", "logLik"))
for r in rows:
    print("%-30s %8d %14.2f %18.2f %14.1f %16.6f%s" % (r[0], r[1], r[2], r[3], r[4], r[5],
                                                       "" if np.isnan(r[6]) else ("  (= oracle)" if r[5] == r[6] else "  (!= oracle %r)" % r[6])))

# is_correction_spdk: one model + approximation per stored draw, all draws in one launch
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parset import draws, ung_update_model  # noqa: E402

print()
print("is_correction_spdk (theta output), ssm_ung parameter sets, n = 200; oracle on all host cores (%d threads)"
      % orc.L().orc_max_threads())
print("(update_fn is evaluated for every draw in Python before either arm runs: 'update_fn ms'; the two arms are timed")
print(" with that parameter set already built; host approx = the same call with BSSM_B200_HOST_APPROX=1)")
print("%-4s %7s %6s %13s %15s %15s %11s %10s" % ("m", "draws", "nsim", "update_fn ms", "device call ms", "host approx ms",
                                                "oracle ms", "same bits"))
_build = bssm_b200.api.build_parset
for m, ns, nsim in ((1, 2000, 10), (2, 2000, 10), (2, 2000, 100), (2, 15000, 10)):
    model = ung_update_model(n=200, m=m)
    mo = draws(model, ns)
    ps_t = best(lambda: _build(model, mo["theta"]), reps=2)
    ps = _build(model, mo["theta"])
    bssm_b200.api.build_parset = lambda mdl, th: ps  # both arms reuse the parameter set
    run = lambda: bssm_b200.post_correct(model, mo, nsim, is_type="is2", seed=1, sampling_method="spdk")
    run()
    t_dev = best(run, reps=5)
    os.environ["BSSM_B200_HOST_APPROX"] = "1"
    run()
    t_host = best(run, reps=3)
    del os.environ["BSSM_B200_HOST_APPROX"]
    t_cpu = best(lambda: orc.correction_spdk_parset(model, mo, nsim, is_type=2, seed=1), reps=2)
    same = np.array_equal(run()["weights"], orc.correction_spdk_parset(model, mo, nsim, is_type=2, seed=1)[0])
    bssm_b200.api.build_parset = _build
    print("%-4d %7d %6d %13.1f %15.1f %15.1f %11.1f %10s" % (m, ns, nsim, ps_t * 1e3, t_dev * 1e3, t_host * 1e3,
                                                           t_cpu * 1e3, same))
