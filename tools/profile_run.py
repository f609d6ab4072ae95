"""Small driver for ncu: runs the c2 workload (svm psi-APF, N=65536, T=1000) a few times.
usage: python tools/profile_run.py [filters] [runs] [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bssm_b200
import data

filters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
N = int(os.environ.get("PROF_N", "65536"))
model = data.svm_model(n)
if filters == 0:  # auto: as many filters as the device runs concurrently
    probe = bssm_b200.FilterJob(model, N, method="psi", seed=1, n_filters=1, out_flags=0)
    filters = probe.max_concurrent_filters()
    probe.close()
job = bssm_b200.FilterJob(model, N, method="psi", seed=1, n_filters=filters,
                         out_flags=int(os.environ.get("PROF_FLAGS", "7")))
print("filters", filters)
print("max concurrent filters", job.max_concurrent_filters())
for _ in range(runs):
    ms = job.run()
    print("run ms", ms, job.kernel_ms(), "p-s/s %.3e" % (filters * N * n / ms * 1e3), "loglik", job.logliks()[:2])
if os.environ.get('BSSM_B200_PHASES'):
    for cta in (0, 7, 15):
        pc = job.phase_cycles(cta)
        print('phase cycles/step cta', cta, [round(c / n) for c in pc], 'sum', round(sum(pc) / n))
job.close()
