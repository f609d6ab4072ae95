import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, bssm_b200, data
n, nd = 500, 4000
rng = np.random.default_rng(1)
y = rng.poisson(4.0, size=n).astype(float)
model = bssm_b200.bsm_ng(y, sd_level=0.1, sd_slope=0.01, P1=np.diag([1.0, 0.1]), distribution="poisson")
th0 = np.asarray(model.theta)
theta = th0[None, :] + 0.05 * rng.normal(size=(nd, th0.size))
mode = np.log(np.maximum(y, 0.5))
mo = {"theta": theta, "counts": np.ones(nd, dtype=np.uint32), "modes": np.asfortranarray(np.tile(mode[:, None], (1, nd))),
      "output_type": 3, "posterior": np.zeros(nd)}
for _ in range(2):
    t0 = time.perf_counter()
    r = bssm_b200.post_correct(model, mo, 256, is_type="is2", seed=1)
    w = time.perf_counter() - t0
    print("bsm_ng m=2: %d draws x N=256 x n=%d: wall %.1f ms, device %.1f ms, host cores %d" % (nd, n, w * 1e3, r["device_ms"], os.cpu_count()))
