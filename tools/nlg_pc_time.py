"""postcorrection_nonlinear, growth model: wall time of one call with the per-draw approximation on the device
(default) and on the host (BSSM_B200_HOST_APPROX=1)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bssm_b200, data, orc
n, nd = 200, 4000
model = data.growth_model(n)
rng = np.random.default_rng(1)
th0 = np.asarray(model.theta)
theta = th0[None, :] + 0.02 * rng.normal(size=(nd, th0.size))
mode = orc.nlg_psi_filter(model, 4, seed=1).mode.ravel(order="F")
mo = {"theta": theta, "counts": np.ones(nd, dtype=np.uint32), "modes": np.asfortranarray(np.tile(mode[:, None], (1, nd))),
      "output_type": 3, "posterior": np.zeros(nd)}
for label, env in (("device", None), ("host", "1")):
    if env: os.environ["BSSM_B200_HOST_APPROX"] = env
    for _ in range(2):
        t0 = time.perf_counter()
        r = bssm_b200.post_correct(model, mo, 256, is_type="is2", seed=1)
        w = time.perf_counter() - t0
    print("growth: %d draws x N=256 x n=%d, approximation on the %s: wall %.1f ms, filters %.1f ms (host cores %d)" % (
        nd, n, label, w * 1e3, r["device_ms"], os.cpu_count()))
