"""C4 through ONE call of the C ABI per pass: bssm_b200_postcorrection_nongaussian with n_threads = G shards
the 15000 draws over G GPUs of this process (one host thread per device, persistent NCCL communicator, one
ncclAllGather of the log-likelihoods).  Prints wall time per call (host model upload + kernels + gather + D2H)
and the device time of the slowest shard.  usage: python tools/c4_inlib.py [G ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bssm_b200, data

nd, N = 15000, 256
model = data.svm_model(1000)
mo = bssm_b200.synthetic_svm_mcmc_output(model, nd)
Gs = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
ref = None
for G in Gs:
    if G > bssm_b200.lib().bssm_b200_device_count():
        continue
    for _ in range(2):
        r = bssm_b200.post_correct(model, mo, N, threads=G, is_type="is2", seed=1)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = bssm_b200.post_correct(model, mo, N, threads=G, is_type="is2", seed=1)
    wall = (time.perf_counter() - t0) / reps * 1e3
    note = bssm_b200.lib().bssm_b200_last_gather().decode()
    if ref is None:
        ref = r["weights"]
    print("G=%d: wall %.1f ms per call, device (slowest shard) %.1f ms, gather: %s, identical to G=%d: %s"
          % (G, wall, r["device_ms"], note or "single device", Gs[0], bool(np.array_equal(ref, r["weights"]))))
