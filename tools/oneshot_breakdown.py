"""host-side breakdown of a small one-shot call: job creation / run / fetch / destroy (C1: bsm_ng, N = 100)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bssm_b200, data
from bssm_b200.api import FilterJob, POST_FILTER_SUMMARY
model = data.poisson_bsm_model(100)
for N in (100, 1000):
    acc = np.zeros(4)
    for it in range(30):
        t0 = time.perf_counter()
        job = FilterJob(model, N, method="bsf", seed=1 + it, out_flags=3, post=POST_FILTER_SUMMARY)
        t1 = time.perf_counter()
        job.run()
        t2 = time.perf_counter()
        b = job.fetch(0, alpha=True, weights=True, filt=True)
        t3 = time.perf_counter()
        job.close()
        t4 = time.perf_counter()
        if it >= 5:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    print("N=%d: create %.3f ms, run %.3f ms, fetch %.3f ms, destroy %.3f ms" % ((N,) + tuple(acc / 25 * 1e3)))
