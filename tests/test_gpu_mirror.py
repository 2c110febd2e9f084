"""Host path of a whole symmetric Gram (pncov_gram with x1 = NULL, several row blocks): only the
blocks on and above the diagonal cross PCIe, host threads write the rest as mirror images
(csrc/pncov_host.cu).  The result must be the array a full copy gives, bit for bit."""

import numpy as np
import pytest

import probnum_b200 as pb
from probnum_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["matern-f64", "expquad-f32", "ratquad-ld"])
def test_mirrored_host_copy_equals_full_copy(case, monkeypatch):
    rng = np.random.default_rng(3)
    if case == "expquad-f32":
        n, d, k, dt = 17001, 2, pb.ExpQuad((2,), lengthscales=0.5), np.float32  # 1.16 GB of float32: two row blocks
    elif case == "matern-f64":
        n, d, k, dt = 12289, 3, pb.Matern((3,), nu=1.5), np.float64  # three row blocks, ragged last one
    else:
        n, d, k, dt = 12000, 5, pb.RatQuad((5,), alpha=1.3), np.float64
    x = rng.normal(size=(n, d)).astype(dt)
    ld = n + (7 if case == "ratquad-ld" else 0)  # a leading dimension larger than n
    monkeypatch.setenv("PNCOV_GRAM_MIRROR", "on")  # unset, the library picks the faster path per host
    out = np.full((n, ld), -7.0, dtype=dt)
    with _lib.Operator(k._kernel_params(), x, None) as op:
        assert op.gram_dtype == dt
        op.gram(out[:, :n])  # a view: the leading dimension is ld
    K = out[:, :n]
    assert np.array_equal(K, K.T)
    assert np.all(out[:, n:] == -7.0)  # nothing written outside the matrix
    monkeypatch.setenv("PNCOV_GRAM_MIRROR", "off")
    full = np.empty((n, n), dtype=dt)
    with _lib.Operator(k._kernel_params(), x, None) as op:
        op.gram(full)
    assert np.array_equal(K, full)
    # spot check against the mirror API on a few rows
    rows = np.array([0, 1, 4095, 4096, n - 1])
    np.testing.assert_allclose(K[rows], k.matrix(x[rows], x), rtol=1e-12 if dt == np.float64 else 1e-5, atol=1e-12 if dt == np.float64 else 1e-5)
