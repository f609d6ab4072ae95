"""ctypes binding of include/bssm_b200.h (the C ABI).  No torch types here."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSSM_B200_LIB: developer override used to A/B kernel build variants (tools/build_variants.sh)
LIB_PATH = os.environ.get("BSSM_B200_LIB") or os.path.join(_HERE, "csrc", "libbssm_b200.so")

c_double_p = C.POINTER(C.c_double)
c_uint32_p = C.POINTER(C.c_uint32)


class UngModel(C.Structure):
    """bssm_ung_model"""
    _fields_ = [
        ("n", C.c_int32), ("m", C.c_int32), ("k", C.c_int32),
        ("distribution", C.c_int32),
        ("Ztv", C.c_int32), ("Ttv", C.c_int32), ("Rtv", C.c_int32),
        ("Dtv", C.c_int32), ("Ctv", C.c_int32),
        ("y", c_double_p), ("Z", c_double_p), ("T", c_double_p), ("R", c_double_p),
        ("a1", c_double_p), ("P1", c_double_p), ("D", c_double_p), ("C", c_double_p),
        ("xbeta", c_double_p), ("u", c_double_p),
        ("phi", C.c_double),
        ("max_iter", C.c_int32), ("local_approx", C.c_int32),
        ("conv_tol", C.c_double),
        ("initial_mode", c_double_p),
        ("H", c_double_p), ("Htv", C.c_int32),
    ]


class FilterOut(C.Structure):
    """bssm_filter_out"""
    _fields_ = [
        ("alpha", c_double_p), ("weights", c_double_p), ("indices", c_uint32_p),
        ("at", c_double_p), ("att", c_double_p), ("Pt", c_double_p), ("Ptt", c_double_p),
        ("alphahat", c_double_p), ("Vt", c_double_p),
        ("loglik", C.c_double), ("approx_loglik", C.c_double),
        ("stopped_at", C.c_int32), ("device_ms", C.c_float),
    ]


class MngModel(C.Structure):
    """bssm_mng_model"""
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("k", C.c_int32), ("p", C.c_int32),
                ("Ztv", C.c_int32), ("Ttv", C.c_int32), ("Rtv", C.c_int32), ("Dtv", C.c_int32), ("Ctv", C.c_int32),
                ("distribution", C.POINTER(C.c_int32)),
                ("y", c_double_p), ("u", c_double_p), ("Z", c_double_p), ("T", c_double_p), ("R", c_double_p),
                ("a1", c_double_p), ("P1", c_double_p), ("D", c_double_p), ("C", c_double_p), ("phi", c_double_p),
                ("max_iter", C.c_int32), ("local_approx", C.c_int32), ("conv_tol", C.c_double),
                ("initial_mode", c_double_p)]


class NlgModel(C.Structure):
    """bssm_nlg_model"""
    _fields_ = [("model_id", C.c_int32), ("n", C.c_int32), ("p", C.c_int32), ("m", C.c_int32), ("k", C.c_int32),
                ("y", c_double_p), ("theta", c_double_p), ("n_theta", C.c_int32),
                ("known_params", c_double_p), ("n_known", C.c_int32)]


class SdeModel(C.Structure):
    """bssm_sde_model"""
    _fields_ = [("model_id", C.c_int32), ("n", C.c_int32), ("y", c_double_p), ("theta", c_double_p),
                ("n_theta", C.c_int32), ("x0", C.c_double), ("positive", C.c_int32)]


class ThetaMap(C.Structure):
    """bssm_theta_map"""
    _fields_ = [("model_type", C.c_int32), ("svm_type", C.c_int32), ("mu_est", C.c_int32), ("phi_est", C.c_int32),
                ("slope", C.c_int32), ("seasonal", C.c_int32), ("noise", C.c_int32),
                ("level_est", C.c_int32), ("slope_est", C.c_int32), ("seasonal_est", C.c_int32),
                ("n_xreg", C.c_int32), ("xreg", c_double_p)]


class PostcorrOut(C.Structure):
    """bssm_postcorr_out"""
    _fields_ = [("weights", c_double_p), ("posterior", c_double_p), ("alpha", c_double_p),
                ("alphahat", c_double_p), ("Vt", c_double_p), ("loglik", c_double_p),
                ("sum_w", C.c_double), ("device_ms", C.c_float), ("launches", C.c_int32)]


class Parset(C.Structure):
    """bssm_parset (parset_ung / parset_mng, src/parset_ng.h:11-62)"""
    _fields_ = [("phi", c_double_p), ("Z", c_double_p), ("T", c_double_p), ("R", c_double_p), ("a1", c_double_p),
                ("P1", c_double_p), ("D", c_double_p), ("C", c_double_p), ("beta", c_double_p),
                ("xreg", c_double_p), ("n_xreg", C.c_int32)]


_lib = None


class EngineError(RuntimeError):
    pass


def lib():
    """Loads libbssm_b200.so; raises loudly if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            "bssm_b200: %s is missing -- build it with `python -m bssm_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.bssm_b200_last_error.restype = C.c_char_p
    L.bssm_b200_version.restype = C.c_char_p
    L.bssm_b200_device_count.restype = C.c_int
    L.bssm_b200_set_device.argtypes = [C.c_int]
    L.bssm_b200_shutdown.restype = None
    L.bssm_b200_stratified_sample.argtypes = [c_double_p, C.c_uint32, c_double_p, C.c_uint32, c_uint32_p]
    L.bssm_b200_gaussian_psi_smoother.argtypes = [C.POINTER(UngModel), C.c_uint32, C.c_uint32, c_double_p]
    L.bssm_b200_is_correction_spdk_ung_parset.argtypes = [C.POINTER(UngModel), C.POINTER(Parset), C.c_uint32, C.c_uint32,
                                                          C.c_int, c_uint32_p, c_double_p, c_double_p, C.c_uint32,
                                                          c_double_p, c_double_p]
    L.bssm_b200_importance_sample_ung.argtypes = [C.POINTER(UngModel), C.c_uint32, C.c_int, C.c_uint32, c_double_p,
                                                  c_double_p, c_double_p]
    L.bssm_b200_math_probe.argtypes = [C.c_int, c_double_p, c_double_p, C.c_uint32]
    L.bssm_b200_rng_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, c_double_p, c_double_p]
    for name in ("bssm_b200_bsf", "bssm_b200_bsf_smoother", "bssm_b200_psi_smoother", "bssm_b200_psi_filter"):
        getattr(L, name).argtypes = [C.POINTER(UngModel), C.c_uint32, C.c_uint32, C.POINTER(FilterOut)]
    L.bssm_b200_job_create_ung.argtypes = [
        C.POINTER(UngModel), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
        C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    for name in ("bssm_b200_bsf_mng", "bssm_b200_bsf_smoother_mng", "bssm_b200_psi_smoother_mng",
                 "bssm_b200_psi_filter_mng"):
        getattr(L, name).argtypes = [C.POINTER(MngModel), C.c_uint32, C.c_uint32, C.POINTER(FilterOut)]
    for name in ("bssm_b200_job_create_mng", "bssm_b200_job_create_mng_psi"):
        getattr(L, name).argtypes = [C.POINTER(MngModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    for name in ("bssm_b200_bsf_nlg", "bssm_b200_bsf_smoother_nlg", "bssm_b200_ekpf_nlg",
                 "bssm_b200_ekpf_smoother_nlg"):
        getattr(L, name).argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.POINTER(FilterOut)]
    for name in ("bssm_b200_psi_smoother_nlg", "bssm_b200_psi_filter_nlg"):
        getattr(L, name).argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                     C.c_uint32, C.POINTER(FilterOut)]
    L.bssm_b200_job_create_nlg_psi.argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                               C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    for name in ("bssm_b200_bsf_sde", "bssm_b200_bsf_smoother_sde"):
        getattr(L, name).argtypes = [C.POINTER(SdeModel), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(FilterOut)]
    L.bssm_b200_loglik_sde.argtypes = [C.POINTER(SdeModel), C.c_uint32, C.c_uint32, C.c_uint32, c_double_p]
    for name in ("bssm_b200_job_create_nlg", "bssm_b200_job_create_nlg_ekpf"):
        getattr(L, name).argtypes = [C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bssm_b200_job_create_sde.argtypes = [C.POINTER(SdeModel), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bssm_b200_approximate_mode.argtypes = [C.POINTER(UngModel), c_double_p]
    pc_args = [C.POINTER(UngModel), C.POINTER(ThetaMap)]
    L.bssm_b200_postcorrection_nongaussian.argtypes = pc_args + [
        C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_uint32_p, c_double_p, C.c_uint32, c_double_p,
        C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostcorrOut)]
    L.bssm_b200_job_set_theta.argtypes = [C.c_void_p] + pc_args + [c_double_p, C.c_uint32]
    L.bssm_b200_job_set_theta_nlg.argtypes = [C.c_void_p, C.POINTER(NlgModel), c_double_p, C.c_uint32]
    L.bssm_b200_job_set_theta_sde.argtypes = [C.c_void_p, c_double_p, C.c_uint32]
    L.bssm_b200_is_correction_bsf.argtypes = pc_args + [
        C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_uint32_p, c_double_p, C.c_uint32, c_double_p,
        C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostcorrOut)]
    L.bssm_b200_is_correction_bsf_sde.argtypes = [
        C.POINTER(SdeModel), C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_uint32_p,
        c_double_p, C.c_uint32, c_double_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostcorrOut)]
    L.bssm_b200_suggest_n_nongaussian.argtypes = pc_args + [c_double_p, C.c_uint32, c_uint32_p, C.c_uint32,
                                                            C.c_uint32, C.c_uint32, c_double_p]
    L.bssm_b200_suggest_n_mng.argtypes = [C.POINTER(MngModel), c_uint32_p, C.c_uint32, C.c_uint32, C.c_uint32, c_double_p]
    L.bssm_b200_suggest_n_nonlinear.argtypes = [C.POINTER(NlgModel), c_double_p, C.c_uint32, c_uint32_p,
                                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                                c_double_p]
    L.bssm_b200_postcorrection_nonlinear.argtypes = [
        C.POINTER(NlgModel), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_uint32_p, c_double_p,
        C.c_uint32, c_double_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostcorrOut)]
    L.bssm_b200_is_correction_bsf_nlg.argtypes = L.bssm_b200_postcorrection_nonlinear.argtypes
    L.bssm_b200_sde_register.argtypes = [C.c_char_p, C.POINTER(C.c_int32)]
    L.bssm_b200_nlg_register.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    for name, mt in (("ung", UngModel), ("mng", MngModel)):
        for fn in ("bssm_b200_postcorrection_%s_parset", "bssm_b200_is_correction_bsf_%s_parset"):
            getattr(L, fn % name).argtypes = [
                C.POINTER(mt), C.POINTER(Parset), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, c_uint32_p,
                c_double_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(PostcorrOut)]
    L.bssm_b200_job_create_postcorr_nlg.argtypes = [
        C.POINTER(NlgModel), C.c_uint32, C.c_uint32, C.c_int, c_uint32_p, c_double_p, C.c_uint32, c_double_p,
        C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    L.bssm_b200_job_create_postcorr.argtypes = pc_args + [
        C.c_uint32, C.c_uint32, C.c_int, c_uint32_p, c_double_p, C.c_uint32, c_double_p, C.c_uint32,
        C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    L.bssm_b200_job_run.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.bssm_b200_job_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3)]
    L.bssm_b200_job_launches.argtypes = [C.c_void_p]
    L.bssm_b200_last_gather.restype = C.c_char_p
    L.bssm_b200_last_gather.argtypes = []
    L.bssm_b200_job_phase_cycles.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64 * 6)]
    L.bssm_b200_job_max_concurrent_filters.argtypes = [C.c_void_p]
    L.bssm_b200_job_fetch.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(FilterOut)]
    L.bssm_b200_job_logliks.argtypes = [C.c_void_p, c_double_p]
    L.bssm_b200_job_logliks_dev.argtypes = [C.c_void_p, C.c_void_p]
    L.bssm_b200_job_destroy.argtypes = [C.c_void_p]
    L.bssm_b200_job_destroy.restype = None
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise EngineError("bssm_b200 error %d: %s" % (rc, lib().bssm_b200_last_error().decode()))


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def uptr(a):
    return a.ctypes.data_as(c_uint32_p) if a is not None else None


def f64(a):
    """column-major contiguous binary64 copy, the layout R / Armadillo use"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))
