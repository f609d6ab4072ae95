import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bssm_b200, data
names = ["observed", "begin_weight+logw", "warp max+st", "bar", "2nd max", "exp+store"]
for label, model, method, n in (("svm psi", data.svm_model(1000), "psi", 1000),):
    for N in (100, 1000):
        job = bssm_b200.FilterJob(model, N, method=method, seed=1, n_filters=1, out_flags=0)
        for _ in range(3): job.run()
        pc = job.phase_cycles(0)
        print(label, N, ", ".join("%s %d" % (a, round(c / n)) for a, c in zip(names, pc)))
        job.close()
