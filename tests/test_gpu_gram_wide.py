"""Dense Gram of ExpQuad / RatQuad on WIDE inputs (d > 16): the guarded norm-identity kernel on FP64
tensor cores (csrc/pncov_gram_mma.cu) against the pinned oracle (direct differences, the reference's
arithmetic: _covariance_function.py:757-781, _exponentiated_quadratic.py:89-100,
_rational_quadratic.py:71-84).

Tolerance: the reference's own ``assert_allclose(rtol=1e-12, atol=1e-12)``; where the cloud is narrow
in units of the lengthscale the entries are also checked purely relatively.  Bit-level properties the
reference has and the kernel must keep: ``K == K.T``, a unit diagonal, exactly 1.0 for coincident
points.
"""

import numpy as np
import pytest

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc
from probnum_b200 import _lib

pytestmark = pytest.mark.gpu


def make_kernel(spec: orc.KernelSpec):
    if spec.kind == orc.EXPQUAD:
        return pb.ExpQuad(spec.input_shape, lengthscales=spec.lengthscales)
    return pb.RatQuad(spec.input_shape, lengthscale=spec.lengthscales, alpha=spec.alpha)


def specs_for(d, rng):
    ell = 0.7 * np.sqrt(d)
    return [
        orc.KernelSpec("expquad", (d,), ell),
        orc.KernelSpec("expquad", (d,), ell * rng.uniform(0.6, 1.7, size=d)),  # ARD
        orc.KernelSpec("ratquad", (d,), ell, alpha=1.0),
        orc.KernelSpec("ratquad", (d,), ell, alpha=2.0),
        orc.KernelSpec("ratquad", (d,), 0.5 * ell, alpha=0.7),
    ]


@pytest.mark.parametrize("d", [17, 27, 32, 33, 100, 130])
def test_wide_gram_matches_oracle(d):
    rng = np.random.default_rng(100 + d)
    x0 = rng.normal(size=(413, d))
    x1 = rng.normal(size=(267, d))
    for spec in specs_for(d, rng):
        k = make_kernel(spec)
        for a, b in ((x0, x1), (x1, x0), (x0, None)):
            K = k.matrix(a, b)
            ref = orc.gram(spec, a, b, row_block=64)
            np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-12, err_msg=str(spec))
            # narrow cloud in units of the lengthscale: the identity is accurate relatively as well
            np.testing.assert_allclose(K, ref, rtol=1e-12, atol=0.0, err_msg=str(spec))
            if b is None:
                assert np.array_equal(K, K.T)
                assert np.all(np.diag(K) == 1.0)


@pytest.mark.parametrize("d", [17, 33, 100])
def test_wide_gram_uncentred_and_duplicated_points(d):
    """Clouds far from the origin (the identity's failure mode, removed by centring) and coincident /
    nearly coincident points (cancellation: those entries are redone by direct differences)."""
    rng = np.random.default_rng(7 + d)
    x = rng.normal(size=(600, d)) + 1.0e4
    x[10] = x[20]
    x[11] = x[21] * (1 + 1e-15)
    x[300] = x[301] + 1e-9 * rng.normal(size=d)  # close, but not a duplicate
    for spec in specs_for(d, rng):
        K = make_kernel(spec).matrix(x)
        ref = orc.gram(spec, x, row_block=64)
        np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-12, err_msg=str(spec))
        assert K[10, 20] == 1.0 and K[20, 10] == 1.0
        assert np.array_equal(K, K.T)
        assert np.all(np.diag(K) == 1.0)
        # the near-duplicate pair keeps the reference's RELATIVE accuracy in 1 - K
        np.testing.assert_allclose(1.0 - K[300, 301], 1.0 - ref[300, 301], rtol=1e-3, atol=1e-16)


@pytest.mark.parametrize("d", [20, 100])
def test_wide_gram_guard_trips_for_wide_clouds(d):
    """Two clusters far apart: |xt|^2 is large in units of the lengthscale, so entries within a cluster
    fail the per-entry guard and are redone, entries across underflow or are tiny."""
    rng = np.random.default_rng(17)
    x = rng.normal(size=(500, d))
    x[250:] += 3.0e3
    for spec in (orc.KernelSpec("expquad", (d,), 0.9 * np.sqrt(d)), orc.KernelSpec("ratquad", (d,), 0.9 * np.sqrt(d), alpha=1.5),
                 orc.KernelSpec("ratquad", (d,), 0.5, alpha=1.0)):
        K = make_kernel(spec).matrix(x)
        np.testing.assert_allclose(K, orc.gram(spec, x, row_block=64), rtol=1e-12, atol=1e-12, err_msg=str(spec))
        assert np.array_equal(K, K.T)


def test_wide_gram_underflow():
    """Short lengthscales: results below the fast window go through the robust epilogue (subnormals)
    or are exactly 0.0 like the reference's exp."""
    rng = np.random.default_rng(23)
    d = 40
    x = rng.normal(size=(300, d))
    for ell in (0.05, 0.2, 0.23, 0.3):
        spec = orc.KernelSpec("expquad", (d,), ell)
        K = make_kernel(spec).matrix(x)
        ref = orc.gram(spec, x, row_block=64)
        np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-12)
        assert np.all(K[ref == 0.0] == 0.0)
        nz = ref > 1e-290
        np.testing.assert_allclose(K[nz], ref[nz], rtol=1e-11, atol=0.0)
        sub = (ref > 0) & (ref < 2.3e-308)
        if sub.any():  # subnormal results: absolute spacing 4.9e-324
            assert np.all(np.abs(K[sub] - ref[sub]) <= 1e-10 * ref[sub] + 1e-322)


def test_wide_gram_shapes_and_edges():
    rng = np.random.default_rng(5)
    d = 24
    spec = orc.KernelSpec("expquad", (d,), 3.0)
    k = make_kernel(spec)
    for n0, n1 in ((1, 1), (1, 200), (129, 65), (128, 64), (127, 63), (513, 1)):
        a, b = rng.normal(size=(n0, d)), rng.normal(size=(n1, d))
        np.testing.assert_allclose(k.matrix(a, b), orc.gram(spec, a, b), rtol=1e-12, atol=1e-12)
    assert k.matrix(np.zeros((0, d)), rng.normal(size=(5, d))).shape == (0, 5)
    # non-finite input propagates like the reference
    a = rng.normal(size=(70, d))
    a[3, 2] = np.nan
    K = k.matrix(a)
    assert np.all(np.isnan(K[3, :])) and np.all(np.isnan(K[:, 3]))
    assert np.all(np.isfinite(np.delete(np.delete(K, 3, 0), 3, 1)))


def test_wide_gram_row_blocks_of_one_handle():
    """Launches that start at a row offset of the handle's point set (the prediction path and the host
    path's row blocks use them) agree bit for bit with the whole matrix."""
    rng = np.random.default_rng(9)
    d, n = 48, 1500
    x = rng.normal(size=(n, d))
    k = pb.RatQuad((d,), lengthscale=6.0, alpha=1.3)
    full = k.matrix(x)
    for lo, hi in ((0, 700), (700, 1500), (333, 334), (1000, 1129)):
        op = _lib.Operator(k._kernel_params(), np.ascontiguousarray(x[lo:hi]), x)
        blk = op.gram(np.empty((hi - lo, n)))
        op.close()
        # a cross handle centres on both sets together: same entries within the tolerance, and the
        # diagonal entries (coincident points across the two sets) are exactly 1
        np.testing.assert_allclose(blk, full[lo:hi], rtol=1e-12, atol=1e-12)
        assert np.all(blk[np.arange(hi - lo), np.arange(lo, hi)] == 1.0)


def test_wide_gram_switch_off(monkeypatch):
    """PNCOV_GRAM_IDENT=off keeps direct differences at every d (read once per handle)."""
    rng = np.random.default_rng(2)
    d = 33
    x = rng.normal(size=(300, d))
    spec = orc.KernelSpec("expquad", (d,), 4.0)
    monkeypatch.setenv("PNCOV_GRAM_IDENT", "off")
    K_direct = make_kernel(spec).matrix(x)
    monkeypatch.delenv("PNCOV_GRAM_IDENT")
    K_mma = make_kernel(spec).matrix(x)
    ref = orc.gram(spec, x)
    np.testing.assert_allclose(K_direct, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(K_mma, ref, rtol=1e-12, atol=1e-12)
