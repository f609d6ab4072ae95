"""ctypes wrapper of the CPU oracle (oracle/_build/liborc.so).  Test infrastructure."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from bssm_b200._lib import MngModel, NlgModel, SdeModel, ThetaMap, UngModel, c_double_p, c_uint32_p, dptr, uptr  # struct layouts only

ARITH_REFERENCE, ARITH_REPRO = 0, 1
_L = None


def build():
    """compile the oracle with its own Makefile (g++, no dependencies): building the checker is not using it"""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True, text=True)
    return os.path.join(ROOT, "oracle", "_build", "liborc.so")


def L():
    global _L
    if _L is None:
        _L = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liborc.so"))
        d, u32, i32, sz = C.c_double, C.c_uint32, C.c_int32, C.c_size_t
        for fn in ("orc_exp", "orc_log", "orc_log1p"):
            getattr(_L, fn).restype = d
            getattr(_L, fn).argtypes = [d]
        for fn in ("orc_exp_v", "orc_log_v", "orc_log1p_v", "orc_sin_v", "orc_cos_v"):
            getattr(_L, fn).argtypes = [c_double_p, c_double_p, sz]
        _L.orc_sincos2pi_v.argtypes = [c_double_p, c_double_p, c_double_p, sz]
        _L.orc_philox.argtypes = [c_uint32_p, c_uint32_p, c_uint32_p]
        _L.orc_normals.argtypes = [u32, u32, u32, u32, u32, c_double_p]
        _L.orc_resample_uniforms.argtypes = [u32, u32, u32, u32, c_double_p]
        _L.orc_stratified_sample.argtypes = [c_double_p, u32, c_double_p, u32, c_uint32_p]
        _L.orc_stratified_sample_repro.argtypes = [c_double_p, u32, c_double_p, u32, c_uint32_p]
        _L.orc_ung_bsf_filter.restype = d
        _L.orc_ung_bsf_filter.argtypes = [C.POINTER(UngModel), u32, u32, u32, C.c_int, c_double_p, c_double_p,
                                          c_uint32_p, C.POINTER(i32)]
        _L.orc_ung_psi_filter.restype = d
        _L.orc_ung_psi_filter.argtypes = [C.POINTER(UngModel), c_double_p, u32, u32, u32, C.c_int, c_double_p,
                                          c_double_p, c_uint32_p, c_double_p, C.POINTER(i32)]
        _L.orc_ung_psi_filter_m.restype = d
        _L.orc_ung_psi_filter_m.argtypes = [C.POINTER(UngModel), c_double_p, u32, u32, u32, C.c_int, c_double_p,
                                            c_double_p, c_uint32_p, c_double_p, C.POINTER(i32), c_double_p,
                                            c_double_p, c_double_p]
        _L.orc_ung_psi_prepass.argtypes = [C.POINTER(UngModel), c_double_p] + [c_double_p] * 8
        _L.orc_ung_filter_batch.argtypes = [C.POINTER(UngModel), C.c_int, c_double_p, u32, u32, u32, u32, C.c_int,
                                            C.c_int, c_double_p]
        _L.orc_postcorrection_ung.argtypes = [C.POINTER(UngModel), C.POINTER(ThetaMap), C.c_int, u32, u32, C.c_int,
                                              C.c_int, c_uint32_p, c_double_p, u32, c_double_p, u32, C.c_int,
                                              c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        _L.orc_ulg_bsf_filter.restype = d
        _L.orc_ulg_bsf_filter.argtypes = [C.POINTER(UngModel), u32, u32, u32, C.c_int, c_double_p, c_double_p,
                                          c_uint32_p, C.POINTER(i32)]
        _L.orc_ulg_psi_filter.argtypes = [C.POINTER(UngModel), u32, u32, c_double_p]
        _L.orc_mng_bsf_filter.restype = d
        _L.orc_mng_bsf_filter.argtypes = [C.POINTER(MngModel), u32, u32, u32, C.c_int, c_double_p, c_double_p,
                                          c_uint32_p, C.POINTER(i32)]
        _L.orc_nlg_bsf_filter.restype = d
        _L.orc_nlg_bsf_filter.argtypes = [C.POINTER(NlgModel), u32, u32, u32, C.c_int, c_double_p, c_double_p,
                                          c_uint32_p, C.POINTER(i32)]
        _L.orc_nlg_psi_filter.restype = d
        _L.orc_nlg_psi_filter.argtypes = [C.POINTER(NlgModel), c_double_p, u32, d, u32, u32, u32, C.c_int, c_double_p,
                                          c_double_p, c_uint32_p, c_double_p, C.POINTER(i32), c_double_p]
        _L.orc_sde_bsf_filter.restype = d
        _L.orc_sde_bsf_filter.argtypes = [C.POINTER(SdeModel), u32, u32, u32, u32, C.c_int, c_double_p, c_double_p,
                                          c_uint32_p, C.POINTER(i32)]
        _L.orc_postcorrection_nlg.argtypes = [C.POINTER(NlgModel), C.c_int, u32, u32, C.c_int, C.c_int, c_uint32_p,
                                              c_double_p, u32, c_double_p, u32, C.c_int, c_double_p, c_double_p,
                                              c_double_p, c_double_p, c_double_p, c_double_p]
        _L.orc_max_threads.restype = C.c_int
        _L.orc_filter_smoother.argtypes = [c_double_p, c_uint32_p, C.c_int, C.c_int, u32]
        _L.orc_summary.argtypes = [c_double_p, C.c_int, C.c_int, u32, c_double_p, c_double_p]
        _L.orc_filter_summary.argtypes = [c_double_p, C.c_int, C.c_int, u32, c_double_p] + [c_double_p] * 4
    return _L


def vec(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(L(), fn)(dptr(x), dptr(y), x.size)
    return y


def sincos2pi(u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    s, c = np.empty_like(u), np.empty_like(u)
    L().orc_sincos2pi_v(dptr(u), dptr(s), dptr(c), u.size)
    return s, c


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    L().orc_philox(c, k, o)
    return list(o)


def normals(seed, filt, t, count, sub=0):
    z = np.empty(count)
    L().orc_normals(seed, filt, t, sub, count, dptr(z))
    return z


def resample_uniforms(seed, filt, t, N):
    r = np.empty(N)
    L().orc_resample_uniforms(seed, filt, t, N, dptr(r))
    return r


def stratified_sample(p, r, repro=False):
    p = np.array(p, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    xp = np.zeros(r.size, dtype=np.uint32)
    fn = L().orc_stratified_sample_repro if repro else L().orc_stratified_sample
    fn(dptr(p), p.size, dptr(r), r.size, uptr(xp))
    return xp


class Result:
    pass


def _alloc(model, N):
    n, m = model.n, model.m
    r = Result()
    r.alpha = np.zeros((m, n + 1, N), order="F")
    r.weights = np.zeros((N, n + 1), order="F")
    r.indices = np.zeros((N, n), dtype=np.uint32, order="F")
    return r


def bsf_filter(model, N, seed, filt=0, arith=ARITH_REPRO):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    r.loglik = L().orc_ung_bsf_filter(C.byref(ms), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                                      uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def psi_filter(model, N, seed, filt=0, arith=ARITH_REPRO, mode=None):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    al = C.c_double(0)
    mode_a = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
    r.loglik = L().orc_ung_psi_filter(C.byref(ms), dptr(mode_a), N, seed, filt, arith, dptr(r.alpha),
                                      dptr(r.weights), uptr(r.indices), C.byref(al), C.byref(st))
    r.approx_loglik = al.value
    r.stopped_at = st.value
    return r


def psi_prepass(model, mode=None):
    n = model.n
    ms = model.c_struct()
    r = Result()
    r.mode, r.ytilde, r.Htilde, r.scales = (np.zeros(n) for _ in range(4))
    r.alphahat, r.L, r.Ct = (np.zeros(n + 1) for _ in range(3))
    al = C.c_double(0)
    mode_a = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
    rc = L().orc_ung_psi_prepass(C.byref(ms), dptr(mode_a), dptr(r.mode), dptr(r.ytilde), dptr(r.Htilde),
                                 dptr(r.scales), dptr(r.alphahat), dptr(r.L), dptr(r.Ct), C.byref(al))
    assert rc == 0
    r.approx_loglik = al.value
    return r


def filter_smoother(alpha, indices):
    a = np.array(alpha, order="F")
    m, n1, N = a.shape
    ind = np.asfortranarray(indices, dtype=np.uint32)
    L().orc_filter_smoother(dptr(a), uptr(ind), m, n1, N)
    return a


def summary(alpha):
    a = np.asfortranarray(alpha)
    m, n1, N = a.shape
    mean, cov = np.zeros((m, n1), order="F"), np.zeros((m, m, n1), order="F")
    L().orc_summary(dptr(a), m, n1, N, dptr(mean), dptr(cov))
    return mean, cov


def filter_summary(alpha, weights):
    a = np.asfortranarray(alpha)
    w = np.asfortranarray(weights)
    m, n1, N = a.shape
    n = n1 - 1
    at, att = np.zeros((m, n1), order="F"), np.zeros((m, n), order="F")
    Pt, Ptt = np.zeros((m, m, n1), order="F"), np.zeros((m, m, n), order="F")
    L().orc_filter_summary(dptr(a), m, n1, N, dptr(w), dptr(at), dptr(att), dptr(Pt), dptr(Ptt))
    return at, att, Pt, Ptt


def filter_batch(model, N, method="psi", seed=1, first_filter=0, n_filters=1, arith=ARITH_REFERENCE, threads=0,
                 mode=None):
    ms = model.c_struct()
    ll = np.zeros(n_filters)
    mode_a = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
    L().orc_ung_filter_batch(C.byref(ms), 1 if method == "psi" else 2, dptr(mode_a), N, seed, first_filter,
                             n_filters, arith, threads, dptr(ll))
    return ll


def max_threads():
    return L().orc_max_threads()


def post_correct(model, mcmc_output, N, is_type=2, seed=1, threads=0, arith=ARITH_REPRO):
    """oracle of approx_mcmc::is_correction_psi; mcmc_output as in bssm_b200.post_correct"""
    from bssm_b200.api import _postcorr_inputs, n_draws, theta_map
    theta, counts, modes = _postcorr_inputs(model, mcmc_output)
    ot = int(mcmc_output.get("output_type", 3))
    nd = n_draws(mcmc_output, is_type)
    n, m = model.n, model.m
    r = Result()
    r.weights, r.posterior, r.loglik = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    r.alpha = np.zeros((n + 1, m, nd), order="F") if ot == 1 else None
    r.alphahat = np.zeros((n + 1, m), order="F") if ot == 2 else None
    r.Vt = np.zeros((m, m, n + 1), order="F") if ot == 2 else None
    if model.cls == "ssm_nlg":
        ms = model.c_struct()
        rc = L().orc_postcorrection_nlg(C.byref(ms), ot, N, seed, threads, is_type, uptr(counts), dptr(theta),
                                        theta.shape[0], dptr(modes), counts.size, arith, dptr(r.weights),
                                        dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat), dptr(r.Vt),
                                        dptr(r.loglik))
        assert rc == 0
        return r
    ms, tm = model.c_struct(), theta_map(model)
    if model.cls == "bsm_ng":
        f = L().orc_postcorrection_bsm
        f.argtypes = L().orc_postcorrection_ung.argtypes
        rc = f(C.byref(ms), C.byref(tm), ot, N, seed, threads, is_type, uptr(counts), dptr(theta), theta.shape[0],
               dptr(modes), counts.size, arith, dptr(r.weights), dptr(r.posterior), dptr(r.alpha),
               dptr(r.alphahat), dptr(r.Vt), dptr(r.loglik))
        assert rc == 0
        return r
    rc = L().orc_postcorrection_ung(C.byref(ms), C.byref(tm), ot, N, seed, threads, is_type, uptr(counts),
                                    dptr(theta), theta.shape[0], dptr(modes), counts.size, arith,
                                    dptr(r.weights), dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat),
                                    dptr(r.Vt), dptr(r.loglik))
    assert rc == 0
    return r


def is_correction_bsf(model, mcmc_output, N, is_type=2, seed=1, threads=0, arith=ARITH_REPRO):
    """oracle of approx_mcmc::is_correction_bsf (src/approx_mcmc.cpp:447-518), m = 1 ung-family models"""
    from bssm_b200.api import n_draws, theta_map
    theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    approx = np.ascontiguousarray(mcmc_output["approx_loglik"], dtype=np.float64)
    ot = int(mcmc_output.get("output_type", 3))
    nd = n_draws(mcmc_output, is_type)
    n = model.n
    r = Result()
    r.weights, r.posterior, r.loglik = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    r.alpha = np.zeros((n + 1, 1, nd), order="F") if ot == 1 else None
    r.alphahat = np.zeros((n + 1, 1), order="F") if ot == 2 else None
    r.Vt = np.zeros((1, 1, n + 1), order="F") if ot == 2 else None
    ms, tm = model.c_struct(), theta_map(model)
    f = L().orc_is_correction_bsf_ung
    f.argtypes = [C.POINTER(UngModel), C.POINTER(ThetaMap), C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                  c_uint32_p, c_double_p, C.c_uint32, c_double_p, C.c_uint32, C.c_int] + [c_double_p] * 6
    rc = f(C.byref(ms), C.byref(tm), ot, N, seed, threads, is_type, uptr(counts), dptr(theta), theta.shape[0],
           dptr(approx), counts.size, arith, dptr(r.weights), dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat),
           dptr(r.Vt), dptr(r.loglik))
    assert rc == 0
    return r


def is_correction_bsf_sde(model, mcmc_output, N, Lf, is_type=2, seed=1, threads=0, arith=ARITH_REPRO):
    """oracle of approx_mcmc::is_correction_bsf<ssm_sde> (src/approx_mcmc.cpp:522-591)"""
    from bssm_b200.api import n_draws
    theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    approx = np.ascontiguousarray(mcmc_output["approx_loglik"], dtype=np.float64)
    ot = int(mcmc_output.get("output_type", 3))
    nd = n_draws(mcmc_output, is_type)
    n = model.n
    r = Result()
    r.weights, r.posterior, r.loglik = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    r.alpha = np.zeros((n + 1, 1, nd), order="F") if ot == 1 else None
    r.alphahat = np.zeros((n + 1, 1), order="F") if ot == 2 else None
    r.Vt = np.zeros((1, 1, n + 1), order="F") if ot == 2 else None
    ms = model.c_struct()
    f = L().orc_is_correction_bsf_sde
    f.argtypes = [C.POINTER(SdeModel), C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, c_uint32_p,
                  c_double_p, C.c_uint32, c_double_p, C.c_uint32, C.c_int] + [c_double_p] * 6
    rc = f(C.byref(ms), Lf, ot, N, seed, threads, is_type, uptr(counts), dptr(theta), theta.shape[0], dptr(approx),
           counts.size, arith, dptr(r.weights), dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat), dptr(r.Vt),
           dptr(r.loglik))
    assert rc == 0
    return r


def is_correction_bsf_nlg(model, mcmc_output, N, is_type=2, seed=1, threads=0, arith=ARITH_REPRO):
    """oracle of approx_mcmc::is_correction_bsf<ssm_nlg>; mcmc_output carries theta, counts, approx_loglik"""
    from bssm_b200.api import n_draws
    theta = np.asfortranarray(np.asarray(mcmc_output["theta"], dtype=np.float64).T)
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    approx = np.ascontiguousarray(mcmc_output["approx_loglik"], dtype=np.float64)
    ot = int(mcmc_output.get("output_type", 3))
    nd = n_draws(mcmc_output, is_type)
    n, m = model.n, model.m
    r = Result()
    r.weights, r.posterior, r.loglik = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    r.alpha = np.zeros((n + 1, m, nd), order="F") if ot == 1 else None
    r.alphahat = np.zeros((n + 1, m), order="F") if ot == 2 else None
    r.Vt = np.zeros((m, m, n + 1), order="F") if ot == 2 else None
    ms = model.c_struct()
    f = L().orc_is_correction_bsf_nlg
    f.argtypes = L().orc_postcorrection_nlg.argtypes
    rc = f(C.byref(ms), ot, N, seed, threads, is_type, uptr(counts), dptr(theta), theta.shape[0], dptr(approx),
           counts.size, arith, dptr(r.weights), dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat), dptr(r.Vt),
           dptr(r.loglik))
    assert rc == 0
    return r


def correction_parset(model, mcmc_output, N, is_type=2, seed=1, threads=0, arith=ARITH_REPRO, bsf=False):
    """oracle of is_correction_psi / is_correction_bsf for ssm_ung and ssm_mng models with an update_fn
    (src/approx_mcmc.cpp:685-921, :1038-1273); the parset is built like bssm_b200.post_correct builds it"""
    from bssm_b200.api import build_parset, n_draws
    from bssm_b200._lib import Parset
    mng = model.cls == "ssm_mng"
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    ps = build_parset(model, np.asarray(mcmc_output["theta"], dtype=np.float64))
    modes = approx = None
    if bsf:
        approx = np.ascontiguousarray(mcmc_output["approx_loglik"], dtype=np.float64)
    else:
        rows = model.n * (model.p if mng else 1)
        modes = np.asfortranarray(np.asarray(mcmc_output["modes"], dtype=np.float64).reshape(rows, -1, order="F"))
    ot = int(mcmc_output.get("output_type", 3))
    nd = n_draws(mcmc_output, is_type)
    n, m = model.n, model.m
    r = Result()
    r.weights, r.posterior, r.loglik = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    r.alpha = np.zeros((n + 1, m, nd), order="F") if ot == 1 else None
    r.alphahat = np.zeros((n + 1, m), order="F") if ot == 2 else None
    r.Vt = np.zeros((m, m, n + 1), order="F") if ot == 2 else None
    ms = model.c_struct()
    f = L().orc_correction_mng_parset if mng else L().orc_correction_ung_parset
    f.argtypes = [C.POINTER(MngModel if mng else UngModel), C.POINTER(Parset), C.c_int, C.c_uint32, C.c_uint32,
                  C.c_int, C.c_int, c_uint32_p, c_double_p, c_double_p, C.c_uint32, C.c_int] + [c_double_p] * 6
    rc = f(C.byref(ms), C.byref(ps), ot, N, seed, threads, is_type, uptr(counts), dptr(modes), dptr(approx),
           counts.size, arith, dptr(r.weights), dptr(r.posterior), dptr(r.alpha), dptr(r.alphahat), dptr(r.Vt),
           dptr(r.loglik))
    assert rc == 0
    return r


def correction_spdk_parset(model, mcmc_output, nsim, is_type=2, seed=1, threads=0):
    """oracle of is_correction_spdk (src/approx_mcmc.cpp:606-678), output_type 3: (weights, posterior)"""
    from bssm_b200.api import build_parset, n_draws
    from bssm_b200._lib import Parset
    counts = np.ascontiguousarray(mcmc_output["counts"], dtype=np.uint32)
    ps = build_parset(model, np.asarray(mcmc_output["theta"], dtype=np.float64))
    modes = np.asfortranarray(np.asarray(mcmc_output["modes"], dtype=np.float64).reshape(model.n, -1, order="F"))
    al = mcmc_output.get("approx_loglik")
    al = None if al is None else np.ascontiguousarray(al, dtype=np.float64)
    nd = n_draws(mcmc_output, is_type)
    w, post = np.zeros(nd), np.zeros(nd)
    ms = model.c_struct()
    f = L().orc_is_correction_spdk_ung_parset
    f.argtypes = [C.POINTER(UngModel), C.POINTER(Parset), C.c_uint32, C.c_uint32, C.c_int, C.c_int, c_uint32_p,
                  c_double_p, c_double_p, C.c_uint32, c_double_p, c_double_p]
    rc = f(C.byref(ms), C.byref(ps), nsim, seed, threads, is_type, uptr(counts), dptr(modes), dptr(al), counts.size,
           dptr(w), dptr(post))
    assert rc == 0
    return w, post


def nlg_bsf_filter(model, N, seed, filt=0, arith=ARITH_REPRO):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    r.loglik = L().orc_nlg_bsf_filter(C.byref(ms), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                                      uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def sde_bsf_filter(model, N, Lvl, seed, filt=0, arith=ARITH_REPRO):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    r.loglik = L().orc_sde_bsf_filter(C.byref(ms), N, Lvl, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                                      uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def approx_mode_m(model):
    """converged mode of approximate() for any m (n values)"""
    ms = model.c_struct()
    f = L().orc_ung_approx_mode_m
    f.restype = None
    f.argtypes = [C.POINTER(UngModel), c_double_p]
    mode = np.zeros(model.n)
    f(C.byref(ms), dptr(mode))
    return mode


def psi_filter_m(model, N, seed, filt=0, arith=ARITH_REPRO, mode=None):
    """general-m psi-APF (Gaussian approximation from include/bssm_b200/ung_gauss_approx.h)"""
    r = _alloc(model, N)
    n, m = model.n, model.m
    ms = model.c_struct()
    st = C.c_int32(-1)
    al = C.c_double(0)
    r.alphahat = np.zeros((m, n + 1), order="F")
    r.L = np.zeros((m, m, n + 1), order="F")
    r.Ct = np.zeros((m, m, n + 1), order="F")
    mode_a = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
    r.loglik = L().orc_ung_psi_filter_m(C.byref(ms), dptr(mode_a), N, seed, filt, arith, dptr(r.alpha),
                                        dptr(r.weights), uptr(r.indices), C.byref(al), C.byref(st),
                                        dptr(r.alphahat), dptr(r.L), dptr(r.Ct))
    r.approx_loglik = al.value
    r.stopped_at = st.value
    return r


def importance_sample(model, nsim, seed, use_antithetic=True, approx=False):
    """simulation smoother + SPDK weights: alpha (m, n+1, nsim), weights, (logLik, approx_logLik), alphahat"""
    lib = L()
    lib.orc_importance_sample_ung.argtypes = [C.POINTER(UngModel), C.c_uint32, C.c_int, C.c_uint32] + [c_double_p] * 6
    n, m = model.n, model.m
    ms = model.c_struct()
    alpha = np.zeros((m, n + 1, nsim), order="F")
    w = np.zeros(nsim)
    ll = np.zeros(2)
    ah = np.zeros((m, n + 1), order="F")
    yt, Ht = np.zeros(n), np.zeros(n)
    rc = lib.orc_importance_sample_ung(C.byref(ms), nsim, int(use_antithetic), seed, dptr(alpha), dptr(w), dptr(ll),
                                       dptr(ah), dptr(yt), dptr(Ht))
    assert rc == 0
    if approx:
        return alpha, w, ll, ah, yt, Ht
    return alpha, w, ll, ah


def sim_normal(seed, t, f):
    lib = L()
    lib.orc_sim_normal.restype = C.c_double
    lib.orc_sim_normal.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return lib.orc_sim_normal(seed, t, f)


def mng_bsf_filter(model, N, seed, filt=0, arith=ARITH_REPRO):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    r.loglik = L().orc_mng_bsf_filter(C.byref(ms), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                                      uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def mng_psi_filter(model, N, seed, filt=0, arith=ARITH_REPRO, mode=None):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    al = C.c_double(0)
    f = L().orc_mng_psi_filter
    f.restype = C.c_double
    f.argtypes = [C.POINTER(MngModel), c_double_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_double_p,
                  c_double_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    mode_a = None if mode is None else np.asfortranarray(mode, dtype=np.float64)
    r.loglik = f(C.byref(ms), dptr(mode_a), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                 uptr(r.indices), C.byref(al), C.byref(st))
    r.approx_loglik = al.value
    r.stopped_at = st.value
    return r


def mng_approx(model):
    """Gaussian approximation alone: (approx_loglik, alphahat m x (n+1), mode p x n, scales n)"""
    ms = model.c_struct()
    f = L().orc_mng_approx
    f.restype = C.c_double
    f.argtypes = [C.POINTER(MngModel), c_double_p, c_double_p, c_double_p]
    ah = np.zeros((model.m, model.n + 1), order="F")
    mode = np.zeros((model.p, model.n), order="F")
    sc = np.zeros(model.n)
    al = f(C.byref(ms), dptr(ah), dptr(mode), dptr(sc))
    return al, ah, mode, sc


def nlg_psi_filter(model, N, seed, filt=0, arith=ARITH_REPRO, mode=None, max_iter=100, conv_tol=1e-8, iekf_iter=0):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    al = C.c_double(0)
    r.mode = np.zeros((model.m, model.n), order="F")
    mode_a = None if mode is None else np.asfortranarray(mode, dtype=np.float64)
    if iekf_iter:
        f = L().orc_nlg_psi_filter_iekf
        f.restype = C.c_double
        f.argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                      C.c_int, c_double_p, c_double_p, c_uint32_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                      c_double_p]
        r.loglik = f(C.byref(ms), max_iter, conv_tol, iekf_iter, N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                     uptr(r.indices), C.byref(al), C.byref(st), dptr(r.mode))
        r.approx_loglik = al.value
        r.stopped_at = st.value
        return r
    r.loglik = L().orc_nlg_psi_filter(C.byref(ms), dptr(mode_a), max_iter, conv_tol, N, seed, filt, arith,
                                      dptr(r.alpha), dptr(r.weights), uptr(r.indices), C.byref(al), C.byref(st),
                                      dptr(r.mode))
    r.approx_loglik = al.value
    r.stopped_at = st.value
    return r


def nlg_ekf_filter(model, N, seed, filt=0, arith=ARITH_REPRO):
    """oracle of ssm_nlg::ekf_filter (EKF-based particle filter)"""
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    f = L().orc_nlg_ekf_filter
    f.restype = C.c_double
    f.argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_double_p, c_double_p,
                  c_uint32_p, C.POINTER(C.c_int32)]
    r.loglik = f(C.byref(ms), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights), uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def nlg_ekf_loglik(model, iekf_iter=0):
    """oracle of ssm_nlg::ekf: (log-likelihood, filtered means m x n)"""
    ms = model.c_struct()
    att = np.zeros((model.m, model.n), order="F")
    f = L().orc_nlg_ekf_loglik
    f.restype = C.c_double
    f.argtypes = [C.POINTER(NlgModel), C.c_uint32, c_double_p]
    ll = f(C.byref(ms), int(iekf_iter), dptr(att))
    return ll, att


def ulg_bsf_filter(model, N, seed, filt=0, arith=ARITH_REPRO):
    r = _alloc(model, N)
    ms = model.c_struct()
    st = C.c_int32(-1)
    r.loglik = L().orc_ulg_bsf_filter(C.byref(ms), N, seed, filt, arith, dptr(r.alpha), dptr(r.weights),
                                      uptr(r.indices), C.byref(st))
    r.stopped_at = st.value
    return r


def ulg_psi_filter(model, N, seed):
    alpha = np.zeros((model.m, model.n + 1, N), order="F")
    ms = model.c_struct()
    L().orc_ulg_psi_filter(C.byref(ms), N, seed, dptr(alpha))
    return alpha
