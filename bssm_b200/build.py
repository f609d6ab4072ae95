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
SOURCES = ["api.cu"]


def nvcc_cmd(verbose=False):
    cmd = [
        "nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
        "-fmad=false", "-std=c++17",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-fopenmp",
        "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
        "-lgomp", "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    return cmd


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        deps += [os.path.join(base, f) for f in files if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile libbssm_b200.so if it is missing or older than its sources."""
    if not force and not _stale():
        return LIB
    cmd = nvcc_cmd(verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


def build_oracle():
    """Compile the CPU oracle (test infrastructure) with its own Makefile."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True,
                   capture_output=True, text=True)
    return os.path.join(ROOT, "oracle", "_build", "liborc.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
