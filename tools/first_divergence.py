"""first time step at which the ancestors of the order-independent fixed-point
arithmetic (what the kernels compute) differ from the reference-order binary64 arithmetic, at the
headline size (svm psi-APF, N = 65536, T = 1000), for a few seeds.  CPU only (the oracle).
usage: python tools/first_divergence.py [seeds...] -> profiles/r2_first_divergence.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import data
import orc

seeds = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
model = data.svm_model(1000)
out = {"config": "svm psi-APF, N = 65536, T = 1000 (BASELINE config 2)", "runs": []}
for seed in seeds:
    a = orc.psi_filter(model, 65536, seed=seed, arith=orc.ARITH_REPRO)
    b = orc.psi_filter(model, 65536, seed=seed, arith=orc.ARITH_REFERENCE)
    diff = np.any(a.indices != b.indices, axis=0)
    first = int(np.argmax(diff)) if diff.any() else None
    out["runs"].append({"seed": seed, "first_divergence_step": first, "steps": 1000,
                        "loglik_repro": a.loglik, "loglik_reference_order": b.loglik,
                        "rel_diff_loglik": abs(a.loglik - b.loglik) / abs(b.loglik)})
    print(out["runs"][-1])
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_first_divergence.json"), "w"), indent=1)
