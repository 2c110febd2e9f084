"""GPU parity of the SURVEY.md 8f N3 row: Linear, Polynomial, WhiteNoise, ProductMatern and
device-side covariance arithmetic (Scaled / Sum / Product), through the mirror API and the C ABI.
Tolerance: the reference's ``assert_allclose(rtol=1e-12, atol=1e-12)``
(tests/test_randprocs/test_covfuncs/test_call.py:123-135); 1e-5 for float32 points."""

import numpy as np
import pytest

import probnum_b200 as pb
from oracle import covfuncs_extra_oracle as xo
from probnum_b200 import _lib
from tests.conftest import load_golden_extra, mirror_from_spec

pytestmark = pytest.mark.gpu

G = load_golden_extra()
GRAM_CASES = [c for c in G.cases if c["type"] == "gram"]
CALL_CASES = [c for c in G.cases if c["type"] == "call"]


def _tol(x):
    return 1e-12 if x.dtype == np.float64 else 1e-5


@pytest.mark.parametrize("entry", G.kat, ids=[e["cite"].split("/")[-1] for e in G.kat])
def test_docstring_known_answers(entry):
    got = mirror_from_spec(entry["spec"]).matrix(G.get(entry["x"]))
    np.testing.assert_allclose(got, np.asarray(entry["doc"]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("chunk", range(6))
def test_gram_matches_reference_golden(chunk):
    for case in GRAM_CASES[chunk::6]:
        k = mirror_from_spec(case["spec"])
        x0, x1 = G.get(case["x0"]), G.get(case["x1"])
        want = G.get(case["K"])
        got = k.matrix(x0, x1)
        assert got.shape == want.shape and got.dtype == want.dtype, case["spec"]
        np.testing.assert_allclose(got, want, rtol=_tol(x0), atol=_tol(x0), err_msg=str(case["spec"]))


@pytest.mark.parametrize("chunk", range(3))
def test_call_matches_reference_golden(chunk):
    for case in CALL_CASES[chunk::3]:
        k = mirror_from_spec(case["spec"])
        want = G.get(case["out"])
        got = np.asarray(k(G.get(case["x0"]), G.get(case["x1"])))
        assert got.shape == want.shape, case["spec"]
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=str(case["spec"]))


SPECS_BIG = [
    dict(kind="linear", input_shape=[5], constant=0.5),
    dict(kind="polynomial", input_shape=[5], constant=1.0, exponent=4),
    dict(kind="whitenoise", input_shape=[5], sigma_sq=0.3),
    dict(kind="productmatern", input_shape=[5], lengthscales=[0.5, 0.8, 1.1, 1.4, 1.7], nus=[0.5, 1.5, 2.5, 3.5, 1.5]),
    dict(kind="linear", input_shape=[40], constant=-1.0),
    dict(kind="productmatern", input_shape=[19], lengthscales=2.0, nus=2.5),
    dict(kind="sum", children=[
        dict(kind="scaled", scalar=1.7, child=dict(kind="matern", input_shape=[5], lengthscales=0.8, nu=2.5)),
        dict(kind="product", children=[dict(kind="expquad", input_shape=[5], lengthscales=1.3),
                                       dict(kind="linear", input_shape=[5], constant=0.5)]),
        dict(kind="whitenoise", input_shape=[5], sigma_sq=0.3)]),
]


@pytest.mark.parametrize("spec", SPECS_BIG, ids=[s["kind"] + str(i) for i, s in enumerate(SPECS_BIG)])
@pytest.mark.parametrize("shape", [(1537, None), (700, 1111), (513, 2)])
def test_gram_matches_oracle_ragged(spec, shape):
    """Sizes that are not multiples of any tile edge, cross and symmetric, against the pinned oracle."""
    d = xo.input_shape(spec)[0]
    rng = np.random.default_rng(5)
    n0, n1 = shape
    x0 = rng.normal(size=(n0, d))
    x1 = None if n1 is None else rng.normal(size=(n1, d))
    if n1 is None:
        x0[11] = x0[402]
    got = mirror_from_spec(spec).matrix(x0, x1)
    np.testing.assert_allclose(got, xo.gram(spec, x0, x1), rtol=1e-12, atol=1e-12)


def test_arithmetic_row_blocked(monkeypatch):
    """Results larger than the device block are built in row blocks; same values."""
    from probnum_b200 import covfuncs

    spec = SPECS_BIG[-1]
    x0 = np.random.default_rng(6).normal(size=(900, 5))
    k = mirror_from_spec(spec)
    full = k.matrix(x0)
    monkeypatch.setattr(covfuncs, "_ARITH_BLOCK_BYTES", 8 * 900 * 100)
    blocked = k.matrix(x0)
    np.testing.assert_allclose(blocked, full, rtol=1e-15, atol=0)
    np.testing.assert_allclose(full, xo.gram(spec, x0), rtol=1e-12, atol=1e-12)


def test_arithmetic_stays_on_device():
    """One device->host copy per arithmetic Gram: the launch count is operands + combines."""
    x0 = np.random.default_rng(7).normal(size=(300, 3))
    k = 2.0 * pb.ExpQuad((3,)) + pb.Matern((3,), nu=1.5)
    before = _lib.launch_count()
    K = k.matrix(x0)
    assert _lib.launch_count() - before >= 3
    want = 2.0 * pb.ExpQuad((3,)).matrix(x0) + pb.Matern((3,), nu=1.5).matrix(x0)
    np.testing.assert_allclose(K, want, rtol=1e-15, atol=0)


def test_combine_abi():
    """pncov_combine_device on raw buffers (odd length: the scalar tail)."""
    import ctypes

    n = 1001
    a = np.random.default_rng(8).normal(size=n)
    b = np.random.default_rng(9).normal(size=n)
    lib = _lib.load()
    with _lib.DeviceBuffer(8 * n) as da, _lib.DeviceBuffer(8 * n) as db:
        import torch

        ta = torch.from_numpy(a).cuda()
        tb = torch.from_numpy(b).cuda()
        _lib._check(lib.pncov_combine_device(0, ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()), n,
                                              _lib.COMBINE_LINEAR, 0.5, -2.0, None))
        torch.cuda.synchronize()
        np.testing.assert_allclose(ta.cpu().numpy(), 0.5 * a - 2.0 * b, rtol=1e-15)
        _lib._check(lib.pncov_combine_device(0, ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()), n,
                                              _lib.COMBINE_MULTIPLY, 0.0, 0.0, None))
        torch.cuda.synchronize()
        np.testing.assert_allclose(ta.cpu().numpy(), (0.5 * a - 2.0 * b) * b, rtol=1e-15)
        assert da.ptr and db.ptr


def test_dense_only_families_have_no_matrix_free_product():
    x0 = np.random.default_rng(10).normal(size=(50, 3))
    with pytest.raises(NotImplementedError):
        with _lib.Operator(pb.Linear((3,))._kernel_params(), x0) as op:
            op.matmul(np.ones((50, 1)))
    # no CPU stand-in for the product either: the mirror says so
    with pytest.raises(NotImplementedError):
        pb.Linear((3,), constant=1.0).linop(x0)
    with pytest.raises(NotImplementedError):
        (pb.ExpQuad((3,)) * pb.Matern((3,))).linop(x0)
    # Scaled / Sum nodes compose the matrix-free operators of isotropic operands
    L = (2.0 * pb.ExpQuad((3,)) + pb.Matern((3,), nu=1.5)).linop(x0)
    v = np.random.default_rng(11).normal(size=(50, 2))
    want = (2.0 * pb.ExpQuad((3,)).matrix(x0) + pb.Matern((3,), nu=1.5).matrix(x0)) @ v
    np.testing.assert_allclose(L @ v, want, rtol=1e-12, atol=1e-13)


def test_error_behaviour():
    with pytest.raises(ValueError):
        pb.WhiteNoise((3,), sigma_sq=-1.0)
    with pytest.raises(ValueError):
        pb.ExpQuad((3,)) + pb.ExpQuad((2,))
    with pytest.raises(NotImplementedError):
        pb.Polynomial((3,), exponent=2.5).matrix(np.ones((4, 3)))
    with pytest.raises(ValueError):
        pb.Linear((3,)).matrix(np.ones((4, 2)))


def test_edge_shapes_of_the_dense_only_families():
    """Empty point sets, a single point, scalar inputs, wide d (several coordinate chunks)."""
    rng = np.random.default_rng(12)
    for k, spec in [
        (pb.Linear((3,), constant=0.5), dict(kind="linear", input_shape=[3], constant=0.5)),
        (pb.WhiteNoise((3,), sigma_sq=0.2), dict(kind="whitenoise", input_shape=[3], sigma_sq=0.2)),
        (pb.ProductMatern((3,), lengthscales=1.0, nus=1.5), dict(kind="productmatern", input_shape=[3], lengthscales=1.0, nus=1.5)),
    ]:
        assert k.matrix(np.empty((0, 3))).shape == (0, 0)
        assert k.matrix(np.empty((0, 3)), rng.normal(size=(4, 3))).shape == (0, 4)
        assert k.matrix(rng.normal(size=(4, 3)), np.empty((0, 3))).shape == (4, 0)
        one = rng.normal(size=(1, 3))
        np.testing.assert_allclose(k.matrix(one), xo.gram(spec, one), rtol=1e-12, atol=1e-12)
    ks = pb.Polynomial((), constant=1.0, exponent=2)
    x = rng.normal(size=9)
    np.testing.assert_allclose(ks.matrix(x), (x[:, None] * x[None, :] + 1.0) ** 2, rtol=1e-12, atol=1e-12)
    # d = 70: five coordinate chunks in gram_extra_kernel
    spec = dict(kind="productmatern", input_shape=[70], lengthscales=3.0, nus=0.5)
    x0, x1 = rng.normal(size=(130, 70)), rng.normal(size=(77, 70))
    np.testing.assert_allclose(pb.ProductMatern((70,), lengthscales=3.0, nus=0.5).matrix(x0, x1), xo.gram(spec, x0, x1),
                               rtol=1e-12, atol=1e-12)
    # arithmetic with an empty operand set keeps shapes and dtypes
    assert (2.0 * pb.ExpQuad((3,)) + pb.Linear((3,))).matrix(np.empty((0, 3))).shape == (0, 0)


def test_linear_gram_on_dmma_is_symmetric_and_matches_the_oracle():
    """d > 16: the inner products run as an FP64 GEMM on DMMA (gram_dot_mma_kernel); C_ij and C_ji
    accumulate the same products in the same order, so the symmetric Gram is symmetric bit for bit."""
    rng = np.random.default_rng(14)
    for d, n in ((17, 700), (40, 1300), (100, 513)):
        x = rng.normal(size=(n, d))
        y = rng.normal(size=(n // 3 + 5, d))
        for k, spec in ((pb.Linear((d,), constant=0.25), dict(kind="linear", input_shape=[d], constant=0.25)),
                        (pb.Polynomial((d,), constant=1.0, exponent=2), dict(kind="polynomial", input_shape=[d], constant=1.0, exponent=2))):
            K = k.matrix(x)
            np.testing.assert_array_equal(K, K.T)
            scale = np.abs(x) @ np.abs(x).T + 1.0  # inner products cancel: compare on the scale of sum |x_i y_i|
            ref = xo.gram(spec, x)
            tol = 1e-12 * (scale if spec["kind"] == "linear" else 2 * np.abs(ref) ** 0.5 * scale + scale ** 2 * 1e-3)
            assert np.all(np.abs(K - ref) <= tol + 1e-12)
            Kc = k.matrix(x, y)
            np.testing.assert_allclose(Kc, xo.gram(spec, x, y), rtol=1e-11, atol=1e-11)


def test_update_chain_matches_numpy():
    """pncov_chain_* (SURVEY 8f N2 ii): sums of rank-2 / rank-1 updates applied from device-resident
    vector pairs; growth past the initial capacity; odd n; several columns."""
    from probnum_b200 import _lib

    rng = np.random.default_rng(21)
    for n, pairs in ((1001, 5), (4096, 70)):
        chain = _lib.UpdateChain(n, capacity=4 if pairs < 10 else 64)
        U, V = rng.normal(size=(n, pairs)), rng.normal(size=(n, pairs))
        for i in range(pairs):
            assert chain.append(U[:, i], V[:, i]) == i + 1
        x = rng.normal(size=(n, 3))
        for k in (0, 1, pairs // 2, pairs):
            Uk, Vk = U[:, :k], V[:, :k]
            refs = {0: Uk @ (Vk.T @ x) + Vk @ (Uk.T @ x), 1: Vk @ (Uk.T @ x), 2: Uk @ (Vk.T @ x)}
            bound = np.abs(Uk) @ (np.abs(Vk).T @ np.abs(x)) + np.abs(Vk) @ (np.abs(Uk).T @ np.abs(x))
            for mode, ref in refs.items():
                y = chain.apply(k, mode, x)
                assert np.all(np.abs(y - ref) <= 1e-12 * bound + 1e-300), (n, k, mode)
        with pytest.raises(ValueError):
            chain.apply(pairs + 1, 0, x)
        with pytest.raises(ValueError):
            chain.apply(1, 0, x[:-1])
        chain.close()
