"""c3 / c5 workload drivers: growth-model BSF + smoother at N=2^20, T=200; GBM Milstein BSF N=2^18, T=500, L=4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200
import data

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "c3":
    N, n = int(os.environ.get("PROF_N", 1 << 20)), 200
    model = data.growth_model(n)
    job = bssm_b200.FilterJob(model, N, method="bsf", seed=1, n_filters=1, out_flags=0, post=2)
else:
    N, n = int(os.environ.get("PROF_N", 1 << 18)), 500
    model = data.gbm_model(n)
    job = bssm_b200.FilterJob(model, N, method="bsf", seed=1, n_filters=1, out_flags=7, post=0, L=4)
for _ in range(runs):
    ms = job.run()
    k = job.kernel_ms()
    print(which, "run ms %.2f" % ms, ["%.2f" % x for x in k], "filter p-s/s %.3e" % (N * n / k[1] * 1e3),
          "total p-s/s %.3e" % (N * n / ms * 1e3), "loglik", job.logliks()[0])
job.close()
