"""Where the cycles of one step of the latency kernel (pf_small.cuh) go: consumer thread 0's phases and the
producers' busy time.  Needs a -DPF_PHASE_TIMING=1 build (tools/build_variants.py) and BSSM_B200_PHASES=1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200, data
names = ["weights+max", "exp+scan", "ancestors", "gather+propagate", "wait for producers", "producers busy"]
for label, model, method, n in (("svm psi", data.svm_model(1000), "psi", 1000), ("bsm_ng bsf", data.poisson_bsm_model(100), "bsf", 100)):
    for N in (100, 1000):
        job = bssm_b200.FilterJob(model, N, method=method, seed=1, n_filters=1, out_flags=0)
        for _ in range(3):
            ms = job.run()
        pc = job.phase_cycles(0)
        print("%s N=%d: kernel %s ms; cycles/step: %s; consumer sum %d" % (
            label, N, ["%.3f" % v for v in job.kernel_ms()], ", ".join("%s %d" % (a, round(c / n)) for a, c in zip(names, pc)),
            round(sum(pc[:5]) / n)))
        job.close()
