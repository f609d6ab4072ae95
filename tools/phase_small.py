import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bssm_b200, data
for N in (100, 1000):
    n=1000
    model = data.svm_model(n)
    job = bssm_b200.FilterJob(model, N, method="psi", seed=1, n_filters=1, out_flags=0)
    for _ in range(3): ms = job.run()
    pc = job.phase_cycles(0)
    print("N", N, "ms", ms, job.kernel_ms(), "phase cycles/step", [round(c/n) for c in pc], "sum", round(sum(pc)/n))
    job.close()
