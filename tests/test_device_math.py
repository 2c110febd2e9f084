"""The device math (exp2 / sqrt / reciprocal / log2 of pncov_math.cuh) validated on the CPU:
the header compiles for the host, MUFU seeds are emulated with 22-bit approximations, and
the results are compared with mpmath at 120 bits.  The CUDA build runs the same code."""

import ctypes
import os
import subprocess
import tempfile

import mpmath as mp
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib():
    out = os.path.join(tempfile.mkdtemp(prefix="pncov_host_"), "libhostmath.so")
    subprocess.check_call([
        "g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
        "-I" + os.path.join(ROOT, "probnum_b200", "csrc"), "-o", out,
        os.path.join(ROOT, "tests", "host", "host_math.cpp"),
    ])
    return ctypes.CDLL(out)


def run(lib, name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    getattr(lib, name)(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(x.size))
    return out


def max_ulp(got, x, f):
    mp.mp.prec = 120
    worst = 0.0
    for g, xi in zip(got, x):
        t = f(mp.mpf(float(xi)))
        if t == 0:
            assert g == 0
            continue
        worst = max(worst, float(abs(mp.mpf(float(g)) - t) / mp.mpf(float(np.spacing(abs(float(t)))))))
    return worst


def test_exp2(hostlib):
    rng = np.random.default_rng(0)
    u = np.concatenate([-rng.uniform(0, 1021.9, 6000), -rng.uniform(0, 1, 2000),
                        -10.0 ** rng.uniform(-18, 0, 1000), [0.0, -0.0]])
    assert max_ulp(run(hostlib, "hm_exp2", u), u, lambda t: mp.mpf(2) ** t) < 1.05
    # underflow band, zero, inf, nan: identical to libm up to 1 ulp of a subnormal
    v = np.concatenate([-rng.uniform(1021, 1080, 4000), [-1022.0, -1074.0, -1075.0, -1076.0, -2000.0, -np.inf]])
    got, want = run(hostlib, "hm_exp2", v), np.exp2(v)
    assert np.all(np.abs(got - want) <= np.maximum(np.spacing(want), 5e-324))
    assert np.isnan(run(hostlib, "hm_exp2", np.array([np.nan]))[0])


def test_sqrt_and_reciprocal(hostlib):
    rng = np.random.default_rng(1)
    s = np.concatenate([rng.uniform(0, 100, 6000), 10.0 ** rng.uniform(-300, 300, 4000)])
    assert max_ulp(run(hostlib, "hm_sqrt", s), s, mp.sqrt) < 0.51
    edge = np.array([0.0, 5e-324, 1e-310, np.inf])
    np.testing.assert_array_equal(run(hostlib, "hm_sqrt", edge), np.sqrt(edge))
    b = np.concatenate([1 + rng.uniform(0, 100, 6000), 1 + 10.0 ** rng.uniform(-17, 290, 4000), [1.0]])
    assert max_ulp(run(hostlib, "hm_rcp", b), b, lambda t: 1 / t) < 0.51


def test_log2(hostlib):
    rng = np.random.default_rng(2)
    b = np.concatenate([1 + rng.uniform(0, 100, 4000), 1 + 10.0 ** rng.uniform(-17, 12, 4000), [1.0, 2.0, 1024.0]])
    got = run(hostlib, "hm_log2", b)
    mp.mp.prec = 120
    # absolute error is what matters for b^-alpha = 2^(-alpha log2 b): one ulp of max(L, 1)
    for g, x in zip(got, b):
        L = mp.log(mp.mpf(float(x)), 2)
        assert abs(mp.mpf(float(g)) - L) <= mp.mpf(2) ** -52 * max(L, 1)
    assert abs(run(hostlib, "hm_log2", np.array([1.0]))[0]) < 1e-18  # 2^(-alpha * this) still rounds to 1
