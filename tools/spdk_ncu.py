"""the workload of the ncu capture in profiles/r2_sim_smoother.md: logLik(method = "spdk"), svm n = 1000, 65 536 draws"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200  # noqa: E402
import data  # noqa: E402

model = data.svm_model(n=1000)
for _ in range(2):
    print(bssm_b200.logLik(model, 65536, method="spdk", seed=1))
