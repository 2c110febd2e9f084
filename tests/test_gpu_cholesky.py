"""Hand-written FP64 Cholesky against NumPy / SciPy on the host."""

import numpy as np
import pytest
import scipy.linalg

import probnum_b200 as pb
from probnum_b200 import _lib

pytestmark = pytest.mark.gpu


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    B = rng.normal(size=(n, n + 8))
    return B @ B.T / n + 0.1 * np.eye(n)


@pytest.mark.parametrize("n", [1, 5, 128, 129, 300, 512, 1000, 2048])
def test_factor_and_solve_match_lapack(n):
    A = _spd(n, n)
    ch = _lib.Cholesky(A)
    L = ch.factor()
    want = np.linalg.cholesky(A)
    np.testing.assert_allclose(L, want, rtol=1e-11, atol=1e-13)
    rng = np.random.default_rng(1)
    for k in (1, 2, 3, 11):
        b = rng.normal(size=(n, k))
        x = ch.solve(b)
        ref = scipy.linalg.cho_solve((want, True), b)
        np.testing.assert_allclose(x, ref, rtol=1e-9, atol=1e-12)
        assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(A) * np.linalg.norm(x)
    v = rng.normal(size=n)
    np.testing.assert_allclose(ch.solve(v), scipy.linalg.cho_solve((want, True), v), rtol=1e-9, atol=1e-12)
    ch.close()


def test_chained_substitution_with_more_block_rows_than_sms():
    """pncov_chol_solve for 1-3 columns is one launch per direction whose CTAs wait for each other
    (chol_chain_solve_kernel): 150 block rows do not fit the 148 SMs at once, so the last CTAs start only when the
    first ones have left.  Checked by the residual on the device and by bit-reproducibility of a second solve."""
    import torch

    n = 150 * 128 - 37
    g = torch.Generator(device="cuda").manual_seed(11)
    R = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=g)
    A = (R + R.T) * (0.5 / n ** 0.5)
    del R
    A.diagonal().add_(2.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
    ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
    rng = np.random.default_rng(4)
    for k in (1, 3):
        b = rng.normal(size=(n, k))
        x = ch.solve(b)
        r = (A @ torch.from_numpy(x).cuda()).cpu().numpy() - b
        assert np.linalg.norm(r) <= 1e-12 * n * np.linalg.norm(b)
        np.testing.assert_array_equal(ch.solve(b), x)
    ch.close()


@pytest.mark.parametrize("n", [1, 33, 512, 700, 1500])
def test_upper_triangle_input_like_scipy_cho_factor(n):
    """scipy.linalg.cho_factor(a) (lower=False, _belief_update.py:79) reads the upper triangle only: with
    ``upper=True`` an array whose lower triangle is garbage factors like its symmetric completion, and only the blocks
    on and above the diagonal are uploaded."""
    A = _spd(n, n + 1)
    want = np.linalg.cholesky(A)
    junk = np.triu(A) + np.tril(np.full((n, n), np.nan), -1)
    ch = _lib.Cholesky(junk, upper=True)
    np.testing.assert_allclose(ch.factor(), want, rtol=1e-11, atol=1e-13)
    b = np.random.default_rng(2).normal(size=(n, 2))
    ref = scipy.linalg.cho_solve(scipy.linalg.cho_factor(junk, check_finite=False), b, check_finite=False)
    np.testing.assert_allclose(ch.solve(b), ref, rtol=1e-9, atol=1e-12)
    ch.close()
    # the default keeps reading the lower triangle
    ch = _lib.Cholesky(junk.T)
    np.testing.assert_allclose(ch.factor(), want, rtol=1e-11, atol=1e-13)
    ch.close()


def test_gram_plus_jitter_like_the_belief_update():
    """_belief_update.py:61-79: cho_factor(gram + jitter * I), then cho_solve (:81-103)."""
    rng = np.random.default_rng(3)
    nodes = rng.normal(size=(1500, 4))
    K = pb.ExpQuad((4,)).matrix(nodes)
    jitter = 1e-8
    ch = _lib.Cholesky(K, jitter=jitter)
    z = np.sin(nodes).sum(axis=1)
    got = ch.solve(z)
    cho = scipy.linalg.cho_factor(K + jitter * np.eye(1500))
    want = scipy.linalg.cho_solve(cho, z)
    # the system is ill-conditioned (cond ~ 1e8 / jitter-limited): compare residuals and the
    # quantities the belief update forms from the solution
    Kj = K + jitter * np.eye(1500)
    assert np.linalg.norm(Kj @ got - z) <= 10 * np.linalg.norm(Kj @ want - z) + 1e-9
    assert abs(z @ got - z @ want) <= 1e-7 * abs(z @ want)
    ch.close()


def test_not_positive_definite_raises_like_scipy():
    A = _spd(300, 7)
    A[200, 200] = -5.0
    with pytest.raises(np.linalg.LinAlgError):
        _lib.Cholesky(A)
    with pytest.raises(np.linalg.LinAlgError):
        scipy.linalg.cho_factor(A, lower=True)


def test_device_input_and_timing():
    import torch

    n = 1024
    A = _spd(n, 11)
    Ad = torch.from_numpy(A).cuda()
    ch = _lib.Cholesky(Ad.data_ptr(), on_device=True, n=n)
    np.testing.assert_allclose(ch.factor(), np.linalg.cholesky(A), rtol=1e-11, atol=1e-13)
    assert ch.last_factor_ms() > 0
    ch.close()


def test_factor_from_operator_keeps_the_gram_on_the_device():
    """pncov_chol_factor_gram: the Gram matrix goes straight into the factorisation buffer."""
    from probnum_b200 import covfuncs

    rng = np.random.default_rng(5)
    for n in (640, 1000):  # a multiple of 128 and a padded size
        x = rng.normal(size=(n, 3))
        k = pb.Matern((3,), nu=2.5)
        K = k.matrix(x)
        z = rng.normal(size=n)
        ref = _lib.Cholesky(K, jitter=1e-6)
        with _lib.Operator(k._kernel_params(), x) as op:
            ch = _lib.Cholesky(op, jitter=1e-6)
        np.testing.assert_allclose(ch.factor(), ref.factor(), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(ch.solve(z), ref.solve(z), rtol=1e-9, atol=1e-12)
        # provenance: the host array returned by matrix() is recognised while it is unchanged
        origin = covfuncs.gram_provenance(K)
        assert origin is not None and np.array_equal(origin[1], x)
        K[0, 0] += 1.0
        assert covfuncs.gram_provenance(K) is None
        assert covfuncs.gram_provenance(K.copy()) is None
        ch.close()
        ref.close()


@pytest.mark.parametrize("n,m,d", [(640, 128, 3), (1000, 333, 4), (300, 1, 2), (129, 2100, 3)])
def test_solve_gram_matches_host_linear_algebra(n, m, d):
    """pncov_chol_solve_gram: W = (L L^T)^-1 K(nodes, x) by the blocked DMMA
    substitution with transposed right-hand sides; column sums of W * K and vecs @ W against
    scipy's cho_solve on the oracle's Gram block.  The solves amplify rounding by cond(K + jitter I)
    (~1e4 here), hence 1e-9 rather than the Gram's 1e-12."""
    from oracle import covfuncs_oracle as orc

    rng = np.random.default_rng(n + m)
    nodes, x = rng.normal(size=(n, d)), rng.normal(size=(m, d))
    k = pb.Matern((d,), nu=1.5, lengthscales=0.7)
    spec = orc.KernelSpec("matern", (d,), 0.7, nu=1.5)
    jitter = 1e-3
    K = orc.gram(spec, nodes) + jitter * np.eye(n)
    G = orc.gram(spec, nodes, x)
    W = scipy.linalg.cho_solve(scipy.linalg.cho_factor(K, lower=True), G)
    vecs = rng.normal(size=(3, n))
    with _lib.Operator(k._kernel_params(), nodes) as sq:
        ch = _lib.Cholesky(sq, jitter=jitter)
    before = _lib.launch_count()
    with _lib.Operator(k._kernel_params(), nodes, x) as op:
        quad, dots = ch.solve_gram(op, vecs)
        quad0, dots0 = ch.solve_gram(op)  # no vectors
        quad_f, dots_f = ch.solve_gram(op, vecs, forward_only=True)  # |L^-1 g|^2 and (L^-1 v).(L^-1 g)
        quad_f0, _ = ch.solve_gram(op, forward_only=True)
        # the generic host-to-host solve gives the same W
        np.testing.assert_allclose(ch.solve(G), W, rtol=1e-9, atol=1e-11)
        with pytest.raises(ValueError):
            ch.solve_gram(op, rng.normal(size=(5, n)))
    assert _lib.launch_count() - before > 4 * ((n + 127) // 128) - 2
    assert ch.last_solve_gram_ms > 0.0
    scale = np.sum(np.abs(W * G), axis=0)
    assert np.all(np.abs(quad - np.sum(W * G, axis=0)) <= 1e-9 * scale + 1e-13)
    np.testing.assert_array_equal(quad0, quad)
    assert dots0.shape == (0, m)
    ref = vecs @ W
    assert np.all(np.abs(dots - ref) <= 1e-9 * (np.abs(vecs) @ np.abs(W)) + 1e-13)
    assert np.all(np.abs(quad_f - np.sum(W * G, axis=0)) <= 1e-9 * scale + 1e-13)
    assert np.all(np.abs(dots_f - ref) <= 1e-9 * (np.abs(vecs) @ np.abs(W)) + 1e-13)
    np.testing.assert_array_equal(quad_f0, quad_f)
    with _lib.Operator(k._kernel_params(), x, nodes) as wrong:
        if m != n:
            with pytest.raises(ValueError):
                ch.solve_gram(wrong)
    ch.close()


def test_solve_gram_edge_cases_and_errors():
    """float32 operators (rejected: the factor is float64), a dense-only family
    as the cross Gram, and the raw C ABI's argument checks."""
    import ctypes

    rng = np.random.default_rng(3)
    n, d = 200, 3
    nodes = rng.normal(size=(n, d))
    k = pb.ExpQuad((d,), lengthscales=0.5)
    with _lib.Operator(k._kernel_params(), nodes) as sq:
        ch = _lib.Cholesky(sq, jitter=1e-6)
    x32 = rng.normal(size=(10, d)).astype(np.float32)
    with _lib.Operator(k._kernel_params(), nodes.astype(np.float32), x32) as op32:
        with pytest.raises(ValueError, match="float64"):
            ch.solve_gram(op32)
    # ProductMatern cross Gram against an ExpQuad factor: any family the Gram launcher knows works
    pm = pb.ProductMatern((d,), lengthscales=1.3, nus=1.5)
    x = rng.normal(size=(37, d))
    with _lib.Operator(pm._kernel_params(), nodes, x) as op:
        quad, _ = ch.solve_gram(op)
    G = pm.matrix(nodes, x)
    W = ch.solve(G)
    np.testing.assert_allclose(quad, np.sum(W * G, axis=0), rtol=1e-9, atol=1e-12)
    # raw ABI: NULL output, bad mode, too many vectors
    lib = _lib.load()
    with _lib.Operator(k._kernel_params(), nodes, x) as op:
        q = np.empty(37)
        assert lib.pncov_chol_solve_gram(ch._h, op._h, 0, None, n, None, 37, None, 0, None) == -1
        assert lib.pncov_chol_solve_gram(ch._h, op._h, 0, None, n, None, 37, q.ctypes.data_as(ctypes.c_void_p), 7, None) == -1
        assert b"mode" in lib.pncov_last_error()
        assert lib.pncov_chol_solve_gram(ch._h, op._h, 5, None, n, None, 37, q.ctypes.data_as(ctypes.c_void_p), 0, None) == -1
    ch.close()
