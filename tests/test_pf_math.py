"""include/bssm_b200/pf_math.h on the host: Philox known answers (Random123
vectors) and accuracy of the deterministic libm replacements."""
import math

import numpy as np

import orc


def ulp_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    spacing = np.spacing(np.abs(want))
    return np.max(np.abs(got - want) / spacing)


def test_philox_known_answers():
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert orc.philox([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_exp_accuracy_and_edges():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-745, 709, 200000), rng.normal(0, 3, 200000), rng.uniform(-1e-3, 1e-3, 1000)])
    got = orc.vec("orc_exp_v", x)
    want = np.exp(x)
    kept = x >= -706.5  # below: results < 2^-1020 flush to zero (declared in pf_math.h; glibc returns < 1e-307)
    assert ulp_err(got[kept], want[kept]) <= 2.0
    assert np.all(got[~kept] == 0.0) and np.all(want[~kept] < 2e-307)
    L = orc.L()
    assert L.orc_exp(0.0) == 1.0
    assert L.orc_exp(-math.inf) == 0.0
    assert L.orc_exp(math.inf) == math.inf
    assert math.isnan(L.orc_exp(math.nan))
    assert L.orc_exp(-800.0) == 0.0 and L.orc_exp(800.0) == math.inf


def test_log_accuracy_and_edges():
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(0.5, 2.0, 200000),
                        1 + rng.normal(0, 1e-6, 1000), np.array([5e-324, 1e-310, 2.2250738585072014e-308])])
    got = orc.vec("orc_log_v", x)
    assert ulp_err(got, np.log(x)) <= 2.0
    L = orc.L()
    assert L.orc_log(1.0) == 0.0
    assert L.orc_log(0.0) == -math.inf
    assert L.orc_log(math.inf) == math.inf
    assert math.isnan(L.orc_log(-1.0)) and math.isnan(L.orc_log(math.nan))


def test_log1p_accuracy():
    rng = np.random.default_rng(3)
    x = np.exp(rng.uniform(-40, 40, 200000))
    got = orc.vec("orc_log1p_v", x)
    assert ulp_err(got, np.log1p(x)) <= 4.0


def test_sincos2pi_accuracy():
    rng = np.random.default_rng(4)
    u = np.concatenate([rng.uniform(0, 1, 400000), np.array([0.0, 0.125, 0.25, 0.5, 0.75, 1 - 2.0 ** -53])])
    s, c = orc.sincos2pi(u)
    # reference in extended precision: reduce exactly, then long double sin/cos
    ul = u.astype(np.longdouble)
    ws = np.sin(2 * np.pi * ul).astype(np.float64) if False else None
    # exact quadrant reduction then float64 libm on |a| <= pi/4 (libm error < 1 ulp there)
    x4 = u * 4
    q = np.floor(x4 + 0.5)
    a = (x4 - q) * (np.pi / 2)
    sa, ca = np.sin(a), np.cos(a)
    qi = q.astype(int) & 3
    want_s = np.choose(qi, [sa, ca, -sa, -ca])
    want_c = np.choose(qi, [ca, -sa, -ca, sa])
    assert np.max(np.abs(s - want_s)) <= 3e-16
    assert np.max(np.abs(c - want_c)) <= 3e-16
    assert np.max(np.abs(s * s + c * c - 1)) <= 5e-16
    assert s[-6] == 0.0 and c[-6] == 1.0  # u = 0
    assert s[-4] == 1.0 and c[-3] == -1.0  # u = .25, .5


def test_normals_moments_and_stream_layout():
    z = orc.normals(seed=7, filt=3, t=5, count=400000)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 1e-2
    assert abs(np.mean(z ** 3)) < 2e-2 and abs(np.mean(z ** 4) - 3) < 5e-2
    # counter-based: any sub-range can be regenerated, streams differ by (t, filter, seed)
    assert np.array_equal(orc.normals(7, 3, 5, 100), z[:100])
    assert not np.array_equal(orc.normals(7, 3, 6, 100), z[:100])
    assert not np.array_equal(orc.normals(7, 4, 5, 100), z[:100])
    assert not np.array_equal(orc.normals(8, 3, 5, 100), z[:100])
    r = orc.resample_uniforms(7, 3, 5, 100000)
    assert r.min() >= 0 and r.max() < 1 and abs(r.mean() - 0.5) < 5e-3


def test_sin_cos_accuracy():
    """pf_sin / pf_cos (the T_fn / T_gn of the nlg_sin_exp model): one multiplication by 1/(2 pi) as
    argument reduction, so the error grows like |x| * 2^-53 in the angle; state values are O(1)"""
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-20, 20, 200000), [0.0, -0.0, 1e-300, -1e-20, np.pi, -np.pi / 2]])
    s, c = orc.vec("orc_sin_v", x), orc.vec("orc_cos_v", x)
    assert np.max(np.abs(s - np.sin(x))) < 4e-15
    assert np.max(np.abs(c - np.cos(x))) < 4e-15
    small = np.abs(x) < 1
    assert np.max(np.abs(s[small] - np.sin(x[small]))) < 5e-16  # absolute: tiny |x| lose their sign to the reduction
