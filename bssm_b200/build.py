"""Builds the engine's shared library in-tree with nvcc for sm_100a.

The built ``bssm_b200/csrc/libbssm_b200.so`` is git-ignored but travels to the
GPU box with the repo snapshot.  ``-fmad=false`` (device) and
``-ffp-contract=off`` (host) are what make the kernels bit-identical to the CPU
oracle: every fused multiply-add in the code base is an explicit fma() of
include/bssm_b200/pf_math.h.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bssm_b200", "csrc")
LIB = os.path.join(CSRC, "libbssm_b200.so")
OBJ = os.path.join(CSRC, "_obj")
# one translation unit per model family: the kernel instantiations dominate build time
SOURCES = ["api.cu", "k_ung_psi1.cu", "k_ung_psim.cu", "k_ung_bsf_a.cu", "k_ung_bsf_b.cu", "k_ung_dyn.cu",
           "k_mng.cu", "k_mng_draws.cu", "k_nlg.cu", "k_nlg_lin.cu", "k_nlg_draws.cu", "k_sde.cu", "user_models.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-fopenmp",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def compile_cmd(src, obj, verbose=False, defines=()):
    cmd = ["nvcc"] + NVCC_FLAGS + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    return cmd


def link_cmd(objs, lib=None):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fopenmp",
            "-lgomp", "-ldl", "-lpthread", "-o", lib or LIB] + objs


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        deps += [os.path.join(base, f) for f in files if f.endswith(".h")]
    return deps


def _obj_of(src, objdir=None):
    return os.path.join(objdir or OBJ, src[:-3] + ".o")


def _stale_sources(objdir=None):
    """Sources whose object is missing or older than the source or any header."""
    newest_header = max(os.path.getmtime(h) for h in _headers())
    out = []
    for src in SOURCES:
        obj = _obj_of(src, objdir)
        if not os.path.exists(obj):
            out.append(src)
            continue
        t = os.path.getmtime(obj)
        if os.path.getmtime(os.path.join(CSRC, src)) > t or newest_header > t:
            out.append(src)
    return out


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _headers() + [os.path.join(CSRC, s) for s in SOURCES]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), lib=None, objdir=None):
    """Compile libbssm_b200.so if it is missing or older than its sources.

    Translation units are compiled in parallel (one nvcc process each) into
    bssm_b200/csrc/_obj and linked into the shared library.  `defines`, `lib`
    and `objdir` build an A/B variant elsewhere (tools/build_variants.py)."""
    variant = lib is not None
    if not force and not variant and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = objdir or OBJ
    os.makedirs(objdir, exist_ok=True)
    todo = list(SOURCES) if (force or variant) else _stale_sources(objdir)
    log = []

    def one(src):
        cmd = compile_cmd(src, _obj_of(src, objdir), verbose, defines)
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for cmd, res in ex.map(one, todo):
                log.append(res.stdout + res.stderr)
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = link_cmd([_obj_of(s, objdir) for s in SOURCES], lib)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write("".join(log))
    return lib or LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
