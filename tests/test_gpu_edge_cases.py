"""Edge cases of the device path: empty and single-point sets, ragged / non-contiguous /
integer / mixed-dtype inputs, very large d, non-finite coordinates, extreme hyper-parameters.
Reference behaviour is taken from the oracle (bit-identical to ProbNum, tests/test_oracle.py)."""

import numpy as np
import pytest

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc

pytestmark = pytest.mark.gpu

KERNELS = [
    (pb.ExpQuad((3,), lengthscales=0.9), orc.KernelSpec("expquad", (3,), 0.9)),
    (pb.Matern((3,), nu=0.5, lengthscales=1.1), orc.KernelSpec("matern", (3,), 1.1, nu=0.5)),
    (pb.Matern((3,), nu=2.5, lengthscales=[0.5, 1.0, 2.0]), orc.KernelSpec("matern", (3,), np.array([0.5, 1.0, 2.0]), nu=2.5)),
    (pb.RatQuad((3,), lengthscale=0.8, alpha=1.0), orc.KernelSpec("ratquad", (3,), 0.8, alpha=1.0)),
    (pb.RatQuad((3,), lengthscale=0.8, alpha=0.4), orc.KernelSpec("ratquad", (3,), 0.8, alpha=0.4)),
]
IDS = ["expquad", "matern12", "matern52-ard", "ratquad1", "ratquad-general"]


@pytest.mark.parametrize("k,spec", KERNELS, ids=IDS)
def test_empty_and_single_point_sets(k, spec):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(5, 3))
    e = np.empty((0, 3))
    assert k.matrix(e).shape == (0, 0)
    assert k.matrix(x, e).shape == (5, 0)
    assert k.matrix(e, x).shape == (0, 5)
    one = x[:1]
    np.testing.assert_array_equal(k.matrix(one), [[1.0]])
    np.testing.assert_allclose(k.matrix(one, x), orc.gram(spec, one, x), rtol=1e-12, atol=1e-12)
    L = k.linop(x, e)
    assert L.shape == (5, 0)
    np.testing.assert_array_equal(L @ np.empty((0, 2)), np.zeros((5, 2)))
    Lt = k.linop(e, x)
    assert (Lt @ np.ones(5)).shape == (0,)
    L1 = k.linop(one)
    np.testing.assert_array_equal(L1 @ np.array([2.5]), [2.5])
    assert k(e, None).shape == (0,) and k(e[:, None], x[None]).shape == (0, 5)


@pytest.mark.parametrize("k,spec", KERNELS, ids=IDS)
def test_input_layouts_and_dtypes(k, spec):
    rng = np.random.default_rng(1)
    base = rng.normal(size=(64, 6))
    x_strided = base[::2, ::2]              # non-contiguous view
    x_fortran = np.asfortranarray(rng.normal(size=(41, 3)))
    want = orc.gram(spec, x_strided, x_fortran)
    np.testing.assert_allclose(k.matrix(x_strided, x_fortran), want, rtol=1e-12, atol=1e-12)
    x_list = x_fortran.tolist()             # array-like, not ndarray
    np.testing.assert_allclose(k.matrix(x_strided, x_list), want, rtol=1e-12, atol=1e-12)
    xi = rng.integers(-3, 4, size=(17, 3))  # integers are promoted to float64 on the device path
    np.testing.assert_allclose(k.matrix(xi), orc.gram(spec, xi.astype(float)), rtol=1e-12, atol=1e-12)
    x32 = rng.normal(size=(9, 3)).astype(np.float32)
    mixed = k.matrix(x32, x_fortran)         # float32 x float64 -> float64 arithmetic
    assert mixed.dtype == np.float64
    np.testing.assert_allclose(mixed, orc.gram(spec, x32.astype(np.float64), x_fortran), rtol=1e-12, atol=1e-12)
    v = rng.normal(size=(41, 2))[:, ::-1]   # negative-stride operand
    np.testing.assert_allclose(k.linop(x_strided, x_fortran) @ v, want @ v, rtol=1e-10, atol=1e-12)


def test_very_large_input_dimension():
    rng = np.random.default_rng(2)
    d = 2500
    x0, x1 = rng.normal(size=(70, d)), rng.normal(size=(33, d))
    spec = orc.KernelSpec("expquad", (d,), 40.0)
    k = pb.ExpQuad((d,), lengthscales=40.0)
    np.testing.assert_allclose(k.matrix(x0, x1), orc.gram(spec, x0, x1), rtol=1e-12, atol=1e-12)
    with pytest.raises(NotImplementedError, match="d <= 1024"):
        k.linop(x0, x1) @ np.ones(33)
    d = 700
    x0, x1 = rng.normal(size=(70, d)), rng.normal(size=(33, d))
    spec = orc.KernelSpec("matern", (d,), 25.0, nu=1.5)
    L = pb.Matern((d,), nu=1.5, lengthscales=25.0).linop(x0, x1)
    v = rng.normal(size=33)
    np.testing.assert_allclose(L @ v, orc.matvec(spec, x0, x1, v), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("k,spec", KERNELS, ids=IDS)
def test_non_finite_and_extreme_coordinates(k, spec):
    """inf / nan / huge coordinates take the robust branch of the epilogue and must give what
    NumPy gives (nan where the reference has nan, 0 for infinitely distant points)."""
    x = np.array([[0.0, 0.0, 0.0], [1e200, 0.0, 0.0], [np.inf, 0.0, 0.0], [np.nan, 1.0, 2.0],
                  [1e-200, 0.0, 0.0], [3.0, -4.0, 12.0]])
    with np.errstate(all="ignore"):
        want = orc.gram(spec, x)
    got = k.matrix(x)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    fin = ~np.isnan(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12, atol=1e-12)


def test_extreme_hyperparameters():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(200, 2))
    for ls in (1e-6, 1e-2, 1e3, 1e8):
        for kk, spec in [(pb.ExpQuad((2,), lengthscales=ls), orc.KernelSpec("expquad", (2,), ls)),
                         (pb.Matern((2,), nu=1.5, lengthscales=ls), orc.KernelSpec("matern", (2,), ls, nu=1.5)),
                         (pb.RatQuad((2,), lengthscale=ls, alpha=3.0), orc.KernelSpec("ratquad", (2,), ls, alpha=3.0))]:
            with np.errstate(all="ignore"):
                want = orc.gram(spec, x)
            np.testing.assert_allclose(kk.matrix(x), want, rtol=1e-12, atol=1e-12, err_msg=f"{spec.kind} l={ls}")
    for alpha in (1e-3, 0.5, 2.0, 37.5, 400.0):
        spec = orc.KernelSpec("ratquad", (2,), 0.7, alpha=alpha)
        np.testing.assert_allclose(pb.RatQuad((2,), lengthscale=0.7, alpha=alpha).matrix(x), orc.gram(spec, x),
                                   rtol=1e-12, atol=1e-12, err_msg=f"alpha={alpha}")
    for nu in (0.5, 1.5, 2.5, 3.5, 5.5, 8.5):
        spec = orc.KernelSpec("matern", (2,), 0.9, nu=nu)
        np.testing.assert_allclose(pb.Matern((2,), nu=nu, lengthscales=0.9).matrix(x), orc.gram(spec, x),
                                   rtol=1e-12, atol=1e-12, err_msg=f"nu={nu}")
    with pytest.raises(NotImplementedError):
        pb.Matern((2,), nu=9.5).matrix(x)
    with pytest.raises(NotImplementedError):
        pb.Matern((2,), nu=0.7).matrix(x)


def test_float32_products_and_call():
    rng = np.random.default_rng(4)
    x0 = rng.normal(size=(300, 4)).astype(np.float32)
    x1 = rng.normal(size=(120, 4)).astype(np.float32)
    v = rng.normal(size=(120, 3)).astype(np.float32)
    for kk, spec in [(pb.ExpQuad((4,)), orc.KernelSpec("expquad", (4,), 1.0)),
                     (pb.Matern((4,), nu=1.5), orc.KernelSpec("matern", (4,), 1.0, nu=1.5)),
                     (pb.RatQuad((4,), alpha=2.0), orc.KernelSpec("ratquad", (4,), 1.0, alpha=2.0))]:
        L = kk.linop(x0, x1)
        assert L.dtype == np.float32
        y = L @ v
        K = orc.gram(spec, x0, x1).astype(np.float64)
        bound = np.abs(K) @ np.abs(v.astype(np.float64))
        assert np.all(np.abs(y - K @ v.astype(np.float64)) <= 1e-5 * bound)
        out = kk(x0[:50], x1[:50])            # elementwise pairs
        want = orc.evaluate(spec, x0[:50], x1[:50])
        assert out.dtype == want.dtype
        np.testing.assert_allclose(out, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("seed", range(6))
def test_randomized_shapes_and_kernels(seed):
    """Seeded sweep over shapes that straddle every dispatch boundary: 64-row tiles, 512-column
    strips, d = 8 / 16 transitions, odd leading dimensions, column splits of the matvec."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(8):
        d = int(rng.choice([1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31]))
        n0 = int(rng.choice([1, 63, 64, 65, 511, 512, 513, 1023, 1025, 1536, 2049])) + int(rng.integers(0, 3))
        n1 = int(rng.choice([1, 2, 127, 128, 129, 511, 512, 513, 1024, 1100])) + int(rng.integers(0, 3))
        kind = rng.choice(["expquad", "matern", "ratquad"])
        ard = kind != "ratquad" and rng.random() < 0.5 and d > 1
        ls = (rng.gamma(2.0, size=d) + 0.5) * np.sqrt(d) if ard else float(rng.uniform(0.5, 2.0) * np.sqrt(d))
        nu = float(rng.choice([0.5, 1.5, 2.5, 4.5]))
        alpha = float(rng.choice([1.0, 2.0, 0.6, 5.5]))
        spec = orc.KernelSpec(kind, (d,), ls, nu=nu, alpha=alpha)
        if kind == "expquad":
            k = pb.ExpQuad((d,), lengthscales=ls)
        elif kind == "matern":
            k = pb.Matern((d,), nu=nu, lengthscales=ls)
        else:
            k = pb.RatQuad((d,), lengthscale=ls, alpha=alpha)
        x0 = rng.normal(size=(n0, d)) + rng.choice([0.0, 25.0])
        same = rng.random() < 0.3
        x1 = None if same else rng.normal(size=(n1, d)) + rng.choice([0.0, 25.0])
        tag = f"{kind} d={d} n0={n0} n1={'same' if same else n1} ard={ard} nu={nu} alpha={alpha}"
        want = orc.gram(spec, x0, x1)
        np.testing.assert_allclose(k.matrix(x0, x1), want, rtol=1e-12, atol=1e-12, err_msg=tag)
        cols = want.shape[1]
        v = rng.normal(size=(cols, int(rng.choice([1, 2, 4, 5]))))
        L = k.linop(x0, x1)
        y = L @ v
        bound = np.abs(want) @ np.abs(v)
        assert np.all(np.abs(y - want @ v) <= 1e-12 * bound + 1e-300), tag
        L.close()
