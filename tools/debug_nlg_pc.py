import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, bssm_b200, data, orc
model = data.growth_model(40)
mo = data.growth_mcmc_output(model, 12)
job = bssm_b200.PostcorrJob(model, mo, 64, is_type="is2", seed=3, out_flags=7)
job.run()
ll = job.logliks()
res = bssm_b200.post_correct(model, mo, 64, is_type="is2", seed=3, return_loglik=True)
print("job vs function", ll - res["loglik"])
for i in range(12):
    mi = bssm_b200.ssm_nlg(model.y, "growth", theta=mo["theta"][i], known_params=model.known_params)
    r = orc.nlg_psi_filter(mi, 64, seed=3, filt=i, mode=mo["modes"][:, i].reshape(2, -1, order="F"))
    b = job.fetch(i, alpha=True, weights=True, indices=True)
    weq = b.weights.view(np.uint64) == r.weights.view(np.uint64)
    msg = ""
    if not weq.all():
        t = np.where(~weq.all(axis=0))[0][0]
        msg = "first bad col %d max rel diff %.3g n_bad_cols %d" % (t, np.max(np.abs(b.weights[:, t] - r.weights[:, t]) / r.weights[:, t]), (~weq.all(axis=0)).sum())
    print(i, "loglik dev-orc", ll[i] - r.loglik, "weights equal", weq.all(), "alpha eq", np.array_equal(b.alpha, r.alpha), "idx eq", np.array_equal(b.indices, r.indices), msg)
