"""Host-side logic of the mirror API that needs no GPU: argument validation with the
reference's exception types and messages, shape bookkeeping, kernel parameters handed to the
C ABI, and the ``probnum.linops`` subset (exercised with a dense stand-in operator)."""

import fractions

import numpy as np
import pytest

import probnum_b200 as pb
from probnum_b200 import _lib, linops
from probnum_b200.sharded import shard_bounds


# -- parameter validation: tests/test_randprocs/test_covfuncs/test_exponentiated_quadratic.py:9-13,
#    test_matern.py:11-22, test_rational_quadratic.py:8-12 -----------------------------------------
@pytest.mark.parametrize("lengthscales", [0.0, -1.0, (0.0, 1.0), (-0.2, 2.0)])
def test_nonpositive_lengthscales_raise(lengthscales):
    with pytest.raises(ValueError):
        pb.ExpQuad(np.shape(lengthscales), lengthscales=lengthscales)
    with pytest.raises(ValueError):
        pb.Matern(np.shape(lengthscales), lengthscales=lengthscales)


@pytest.mark.parametrize("nu", [-1, -1.0, 0.0, 0])
def test_nonpositive_nu_raises(nu):
    with pytest.raises(ValueError):
        pb.Matern((), nu=nu)


@pytest.mark.parametrize("alpha", [-1, -1.0, 0.0, 0])
def test_nonpositive_alpha_raises(alpha):
    with pytest.raises(ValueError):
        pb.RatQuad((), alpha=alpha)


def test_lengthscales_must_broadcast_to_input_shape():
    with pytest.raises(ValueError):
        pb.ExpQuad((3,), lengthscales=[1.0, 2.0])
    assert pb.ExpQuad((3,), lengthscales=[1.0, 2.0, 3.0]).lengthscales.shape == (3,)
    with pytest.raises(ValueError):
        pb.ExpQuad((3,), lengthscales=[1.0, 2.0, 3.0]).lengthscale
    assert pb.ExpQuad((3,), lengthscales=2.0).lengthscale == 2.0


# -- shape errors: test_call.py:138-186, test_linop_matrix.py:157-191 ------------------------------
@pytest.mark.parametrize("shape", [(), (1,), (10,), (1, 10)])
def test_call_wrong_input_dimension_raises(shape):
    k = pb.Matern((3,), nu=1.5)
    x_bad = np.zeros(shape + (4,))
    x_ok = np.zeros((1, 3))
    with pytest.raises(ValueError, match="input_shape_0"):
        k(x_bad, None)
    with pytest.raises(ValueError, match="input_shape_0"):
        k(x_bad, x_ok)
    with pytest.raises(ValueError, match="input_shape_1"):
        k(x_ok, x_bad)


@pytest.mark.parametrize("x0_shape,x1_shape", [((2,), (8,)), ((2, 5), (3, 5)), ((4, 4), (4, 2))])
def test_call_unbroadcastable_raises(x0_shape, x1_shape):
    k = pb.ExpQuad((3,))
    with pytest.raises(ValueError, match="can not be broadcast"):
        k(np.zeros(x0_shape + (3,)), np.zeros(x1_shape + (3,)))


@pytest.mark.parametrize("method", ["matrix", "linop"])
def test_matrix_linop_wrong_shape_raises(method):
    k = pb.RatQuad((3,))
    with pytest.raises(ValueError, match="x0"):
        getattr(k, method)(np.zeros((5, 2)))
    with pytest.raises(ValueError, match="x1"):
        getattr(k, method)(np.zeros((5, 3)), np.zeros((4, 2)))
    with pytest.raises(ValueError, match="x0"):
        getattr(k, method)(np.zeros(()))


def test_diagonal_call_needs_no_device():
    """x1=None returns exact ones without touching the GPU, like the reference
    (_exponentiated_quadratic.py:90-94)."""
    for k in (pb.ExpQuad((3,)), pb.Matern((3,), nu=2.5), pb.RatQuad((3,))):
        out = k(np.zeros((7, 2, 3)), None)
        assert out.shape == (7, 2) and out.dtype == np.float64 and np.all(out == 1.0)
    assert pb.ExpQuad(())(np.float32(2.0), None).dtype == np.float32
    assert pb.Matern(())(np.zeros(3, dtype=np.float32), None).dtype == np.float64


def test_shapes_and_repr():
    k = pb.Matern(3, nu=2.5)
    assert k.input_shape == (3,) and k.input_ndim == 1 and k.input_size == 3
    assert k.output_shape_0 == () and k.output_size_1 == 1
    assert "Matern" in repr(k) and "input_shape_0=(3,)" in repr(k)
    assert pb.ExpQuad(()).input_size == 1
    assert k.p == 2 and k.is_half_integer and pb.Matern((), nu=0.7).p is None


def test_matern_coefficients_are_exact_rationals():
    F = fractions.Fraction
    assert pb.Matern.half_integer_coefficients(0) == (F(1),)
    assert pb.Matern.half_integer_coefficients(2) == (F(1), F(1), F(1, 3))
    assert pb.Matern.half_integer_coefficients(3) == (F(1), F(1), F(2, 5), F(1, 15))


def test_kernel_params_for_the_c_abi():
    p = pb.ExpQuad((4,), lengthscales=[1, 2, 3, 4])._kernel_params()
    assert (p.kind, p.d) == (_lib.EXPQUAD, 4) and p.lengthscales.tolist() == [1, 2, 3, 4]
    p = pb.Matern((), nu=1.5, lengthscales=0.3)._kernel_params()
    assert (p.kind, p.d, p.nu) == (_lib.MATERN, 1, 1.5) and p.lengthscales.tolist() == [0.3]
    p = pb.RatQuad((2,), lengthscale=-0.5, alpha=3)._kernel_params()
    assert (p.kind, p.alpha) == (_lib.RATQUAD, 3.0) and p.lengthscales.tolist() == [0.5]
    with pytest.raises(NotImplementedError, match="Bessel"):
        pb.Matern((2,), nu=0.7)._kernel_params()
    with pytest.raises(NotImplementedError):
        pb.Matern((2,), nu=20.5)._kernel_params()


def test_compute_and_result_dtypes():
    k = pb.ExpQuad((2,))
    f32, f64 = np.zeros((1, 2), np.float32), np.zeros((1, 2))
    assert k._compute_dtype(f32, f32) == np.float32 and k._compute_dtype(f32, f64) == np.float64
    assert k._compute_dtype(np.zeros((1, 2), int), None) == np.float64
    assert k._result_dtype(np.dtype(np.float32)) == np.float32
    assert pb.Matern((2,))._result_dtype(np.dtype(np.float32)) == np.float64  # NumPy-2 promotion


# -- linops subset ---------------------------------------------------------------------------------
class Dense(linops.LinearOperator):
    def __init__(self, A):
        self.A = A
        super().__init__(A.shape, A.dtype)

    def _matmul(self, x):
        return self.A @ x


def test_linop_matmul_shapes_and_errors():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(5, 3))
    L = Dense(A)
    np.testing.assert_allclose(L @ np.ones(3), A @ np.ones(3))
    x = rng.normal(size=(2, 3, 4))
    np.testing.assert_allclose(L @ x, A @ x)
    w = rng.normal(size=(5,))
    np.testing.assert_allclose(w @ L, w @ A)
    np.testing.assert_allclose(L.todense(), A)
    assert L.todense() is L.todense() and L.todense(cache=False) is L.todense()
    np.testing.assert_allclose(L.T.todense(), A.T)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        L @ np.ones(4)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        np.ones(4) @ L


def test_linop_arithmetic():
    rng = np.random.default_rng(1)
    A = rng.normal(size=(4, 4))
    L = Dense(A)
    v = rng.normal(size=4)
    S = L + 1e-3
    assert isinstance(S, linops.SumLinearOperator)
    np.testing.assert_allclose(S @ v, A @ v + 1e-3 * v)
    np.testing.assert_allclose((L + linops.Identity(4)) @ v, A @ v + v)
    np.testing.assert_allclose((2.5 * L) @ v, 2.5 * (A @ v))
    np.testing.assert_allclose((-L).todense(), -A)
    np.testing.assert_allclose((L - linops.Identity(4)).todense(), A - np.eye(4))
    np.testing.assert_allclose(S.diagonal(), np.diag(A) + 1e-3)
    np.testing.assert_allclose(S.trace(), np.trace(A) + 4e-3)
    np.testing.assert_allclose(S.T.todense(), A.T + 1e-3 * np.eye(4))
    np.testing.assert_allclose((L @ L) @ v, A @ (A @ v))
    with pytest.raises(ValueError):
        Dense(rng.normal(size=(4, 3))) + 1.0


def test_shard_bounds_partition_rows():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
