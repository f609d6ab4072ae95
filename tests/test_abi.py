"""The C-ABI library loads without a GPU, exports every symbol include/bssm_b200.h
declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bssm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bssm_b200_\w+)\s*\(", src)))


def test_header_symbols_exported(engine):
    L = engine.lib()
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), name


def test_version_and_last_error(engine):
    L = engine.lib()
    assert b"sm_100a" in L.bssm_b200_version()
    assert isinstance(L.bssm_b200_last_error(), bytes)


def test_no_cpu_fallback(engine):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import data
    model = data.svm_model(20)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.particle_smoother(model, 16)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.bootstrap_filter(model, 16)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.stratified_sample([0.5, 0.5], [0.1, 0.2])
    # the SPDK entry points likewise: the simulation smoother has no host path in the product
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.importance_sample(model, 16)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.logLik(model, 16, method="spdk")


def test_argument_checks(engine):
    import data
    model = data.svm_model(20)
    with pytest.raises(ValueError):
        engine.bootstrap_filter(model, 1)
    with pytest.raises(ValueError):
        engine.bootstrap_filter(model, 10, seed=-1)
    with pytest.raises(ValueError):
        engine.particle_smoother(model, 10, method="ekf")
    with pytest.raises(ValueError):
        engine.svm(np.zeros(5), 0.9, 0.1, sigma=1.0, mu=0.0)
    with pytest.raises(ValueError):
        engine.importance_sample(model, 1)
    with pytest.raises(ValueError):
        engine.importance_sample(data.lg_model(), 10)
    with pytest.raises(ValueError):  # sampling_method "spdk" needs an ssm_ung model with an update_fn
        engine.post_correct(model, data.svm_mcmc_output(model, 4), 10, sampling_method="spdk")


def test_model_constructors_match_reference_layout(engine):
    import data
    m = data.poisson_bsm_model()
    assert (m.n, m.m, m.k) == (100, 2, 2)
    assert np.array_equal(m.T[:, :, 0], [[1, 1], [0, 1]])
    assert np.array_equal(m.R[:, :, 0], [[0.1, 0], [0, 0.01]])
    assert np.array_equal(m.Z[:, 0], [1, 0])
    assert m.initial_mode[1] == np.log(0.1) and m.initial_mode[0] == np.log(2.0)
    s = data.svm_model(30, rho=0.9, sd_ar=0.3, sigma=2.0)
    assert s.P1[0, 0] == 0.3 ** 2 / (1 - 0.9 ** 2) and s.phi == 2.0
    np.testing.assert_allclose(s.initial_mode, np.log(np.maximum(1e-4, s.y ** 2)) - 2 * np.log(2.0))
    seas = engine.bsm_ng(np.arange(24.0), sd_level=0.1, sd_slope=0.01, sd_seasonal=0.2, period=4,
                         distribution="poisson")
    assert seas.m == 5 and seas.k == 3
    assert np.array_equal(seas.T[2, 2:, 0], [-1, -1, -1])
    assert seas.T[3, 2, 0] == 1 and seas.T[4, 3, 0] == 1


def test_header_is_plain_c_and_r_shim_type_checks():
    """include/bssm_b200.h must compile as C99 (it is the FFI boundary), and the R `.Call` shim a
    maintainer would build (r/bssm_b200_shim.c) must at least type-check against it: R's headers are
    not in this image, so declaration-only stand-ins (tests/mock_R/) take their place -- a syntax /
    signature check, not a build."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Wpedantic", "-Werror", "-x", "c",
                          os.path.join(root, "include", "bssm_b200.h")], capture_output=True, text=True)
    assert hdr.returncode == 0, hdr.stderr
    shim = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(root, "tests", "mock_R"),
                           "-I" + os.path.join(root, "include"), os.path.join(root, "r", "bssm_b200_shim.c")],
                          capture_output=True, text=True)
    assert shim.returncode == 0, shim.stderr
