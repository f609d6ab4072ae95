"""c4 workload driver: post_correct on svm, 15000 draws x psi-APF N=256, T=1000 (weights only).
usage: python tools/profile_c4.py [n_draws] [runs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200
import data

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(os.environ.get("PROF_N", "256"))
model = data.svm_model(1000)
t0 = time.time()
mo = bssm_b200.synthetic_svm_mcmc_output(model, nd)
print("inputs", time.time() - t0, "s")
job = bssm_b200.PostcorrJob(model, mo, N, is_type="is2", seed=1)
print("max concurrent filters", job.max_concurrent_filters())
for _ in range(runs):
    ms = job.run()
    print("run ms", ms, job.kernel_ms(), "draws/s", nd / ms * 1e3, "p-s/s", nd * N * 1000 / ms * 1e3, job.logliks()[:2])
job.close()
