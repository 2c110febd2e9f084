"""Parity of the CUDA path against the pinned oracle and the reference's golden vectors.

Everything here runs on a B200 (``-m gpu``) and goes through the public mirror API
(``probnum_b200.covfuncs``), i.e. through the C ABI of libpncov.so.  The tolerance is
the reference's own: ``assert_allclose(rtol=1e-12, atol=1e-12)``
(tests/test_randprocs/test_covfuncs/test_call.py:123-135, test_linop_matrix.py:111-154)
for float64, 1e-5 for float32; matrix-free products use the componentwise bound
``|dy_i| <= tol * (|K| |v|)_i`` of SURVEY.md 8d.
"""

import zlib

import numpy as np
import pytest

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc
from tests.conftest import load_golden, spec_from_dict

pytestmark = pytest.mark.gpu

G = load_golden()
GRAM_CASES = [c for c in G.cases if c["type"] == "gram"]
CALL_CASES = [c for c in G.cases if c["type"] == "call"]


def make_kernel(spec: orc.KernelSpec):
    if spec.kind == orc.EXPQUAD:
        return pb.ExpQuad(spec.input_shape, lengthscales=spec.lengthscales)
    if spec.kind == orc.MATERN:
        return pb.Matern(spec.input_shape, nu=spec.nu, lengthscales=spec.lengthscales)
    return pb.RatQuad(spec.input_shape, lengthscale=spec.lengthscales, alpha=spec.alpha)


def tol_for(dtype):
    return 1e-12 if np.dtype(dtype) == np.float64 else 1e-5


def assert_matvec_close(got, want, bound, tol):
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert np.all(err <= tol * bound + 1e-300), float(np.max(err / (bound + 1e-300)))


# ---------------------------------------------------------------------------------------------
# golden vectors recorded from the real reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("entry", G.kat, ids=[e["cite"].split("/")[-1] for e in G.kat])
def test_docstring_known_answers(entry):
    spec = spec_from_dict(entry["spec"])
    got = make_kernel(spec).matrix(G.get(entry["x"]))
    np.testing.assert_allclose(got, np.asarray(entry["doc"]), rtol=5e-8)
    np.testing.assert_allclose(got, G.get(entry["ref"]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("chunk", range(8))
def test_gram_matches_reference_golden(chunk):
    for case in GRAM_CASES[chunk::8]:
        spec = spec_from_dict(case["spec"])
        k = make_kernel(spec)
        x0, x1, want = G.get(case["x0"]), G.get(case["x1"]), G.get(case["K"])
        got = k.matrix(x0, x1)
        tol = tol_for(case["dtype"])
        assert got.dtype == want.dtype, case["spec"]
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol, err_msg=str(case["spec"]))


@pytest.mark.parametrize("chunk", range(8))
def test_linop_matches_reference_golden(chunk):
    for case in GRAM_CASES[chunk::8]:
        spec = spec_from_dict(case["spec"])
        k = make_kernel(spec)
        x0, x1, v = G.get(case["x0"]), G.get(case["x1"]), G.get(case["v"])
        tol = tol_for(case["dtype"])
        L = k.linop(x0, x1)
        assert L.shape == G.get(case["K"]).shape
        assert L.dtype == np.dtype(case["dtype"])
        bound = np.abs(G.get(case["K"]).astype(np.float64)) @ np.abs(v.astype(np.float64))
        assert_matvec_close(L @ v, G.get(case["Kv"]), bound, tol)
        assert_matvec_close(L @ v[:, 0], G.get(case["Kv"])[:, 0], bound[:, 0], tol)
        # todense == matrix (test_linop_matrix.py:111-123)
        np.testing.assert_allclose(L.todense(), G.get(case["K"]), rtol=tol, atol=tol)
        L.close()


@pytest.mark.parametrize("chunk", range(4))
def test_call_matches_reference_golden(chunk):
    for case in CALL_CASES[chunk::4]:
        spec = spec_from_dict(case["spec"])
        got = make_kernel(spec)(G.get(case["x0"]), G.get(case["x1"]))
        want = G.get(case["out"])
        assert isinstance(got, (np.ndarray, np.floating))
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# seeded comparisons with the oracle at sizes the oracle finishes in seconds
# ---------------------------------------------------------------------------------------------
SPECS = [
    orc.KernelSpec("expquad", (3,), 1.0),
    orc.KernelSpec("expquad", (8,), 1.5),
    orc.KernelSpec("matern", (3,), 0.7, nu=0.5),
    orc.KernelSpec("matern", (3,), 1.0, nu=1.5),
    orc.KernelSpec("matern", (3,), 1.0, nu=2.5),
    orc.KernelSpec("matern", (4,), 1.3, nu=3.5),
    orc.KernelSpec("ratquad", (3,), 1.0, alpha=1.0),
    orc.KernelSpec("ratquad", (3,), 0.5, alpha=2.0),
    orc.KernelSpec("ratquad", (5,), 1.7, alpha=0.37),
]


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: f"{s.kind}-d{s.input_shape[0]}-nu{s.nu}-a{s.alpha}")
@pytest.mark.parametrize("shape", [(1000, 777), (513, 1025), (2048, None)])
def test_gram_vs_oracle(spec, shape):
    rng = np.random.default_rng(zlib.crc32(repr((spec, shape)).encode()))
    n0, n1 = shape
    x0 = rng.normal(size=(n0,) + spec.input_shape)
    x1 = None if n1 is None else rng.normal(size=(n1,) + spec.input_shape)
    got = make_kernel(spec).matrix(x0, x1)
    np.testing.assert_allclose(got, orc.gram(spec, x0, x1), rtol=1e-12, atol=1e-12)
    if n1 is None:
        np.testing.assert_array_equal(np.diag(got), np.ones(n0))
        np.testing.assert_array_equal(got, got.T)


@pytest.mark.parametrize("d", [1, 2, 5, 7, 10, 16, 17, 33, 100])
@pytest.mark.parametrize("kind", ["expquad", "matern", "ratquad"])
def test_gram_dimensions_and_ard(kind, d):
    rng = np.random.default_rng(d)
    ard = rng.gamma(1.0, size=d) + 0.5
    ls = 0.9 * np.sqrt(d) if kind == "ratquad" else ard * np.sqrt(d)
    spec = orc.KernelSpec(kind, (d,), ls, nu=2.5, alpha=1.5)
    x0 = rng.normal(size=(301, d))
    x1 = rng.normal(size=(155, d))
    k = make_kernel(spec)
    np.testing.assert_allclose(k.matrix(x0, x1), orc.gram(spec, x0, x1), rtol=1e-12, atol=1e-12)
    v = rng.normal(size=(155, 3))
    L = k.linop(x0, x1)
    assert_matvec_close(L @ v, orc.matvec(spec, x0, x1, v), orc.abs_matvec(spec, x0, x1, v), 1e-12)
    vt = rng.normal(size=(301,))
    assert_matvec_close(L.T @ vt, orc.matvec(spec, x1, x0, vt), orc.abs_matvec(spec, x1, x0, vt), 1e-12)
    L.close()


@pytest.mark.parametrize("spec", SPECS[:7:2], ids=lambda s: s.kind + str(s.nu))
def test_uncentred_and_duplicated_points(spec):
    """Direct differences keep full accuracy for clouds far from the origin and for
    coincident points (hard parts): the failure mode of the norm identity."""
    rng = np.random.default_rng(7)
    x = rng.normal(size=(400,) + spec.input_shape) + 1.0e4
    x[10] = x[20]
    x[11] = x[21] * (1 + 1e-15)
    K = make_kernel(spec).matrix(x)
    np.testing.assert_allclose(K, orc.gram(spec, x), rtol=1e-12, atol=1e-12)
    assert K[10, 20] == 1.0 and K[20, 10] == 1.0


def test_scalar_input_shape_and_batches():
    """input_shape = () and multi-dimensional batch shapes are flattened in C order
    (_covariance_function.py:676-693)."""
    rng = np.random.default_rng(3)
    spec = orc.KernelSpec("matern", (), 0.4, nu=1.5)
    k = make_kernel(spec)
    x0 = rng.normal(size=(6, 5))
    x1 = rng.normal(size=(4, 3, 2))
    K = k.matrix(x0, x1)
    assert K.shape == (30, 24)
    np.testing.assert_allclose(K, orc.gram(spec, x0, x1), rtol=1e-12, atol=1e-12)
    spec3 = orc.KernelSpec("expquad", (3,), 1.1)
    k3 = make_kernel(spec3)
    y0 = rng.normal(size=(2, 5, 3))
    np.testing.assert_allclose(k3.matrix(y0), orc.gram(spec3, y0), rtol=1e-12, atol=1e-12)


def test_underflow_and_far_field():
    """exp underflow to subnormals and zero takes the slow branch of pn_exp2."""
    rng = np.random.default_rng(11)
    spec = orc.KernelSpec("expquad", (2,), 0.02)
    x = rng.uniform(-1.0, 1.0, size=(600, 2))
    K = make_kernel(spec).matrix(x)
    ref = orc.gram(spec, x)
    np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-12)
    nz = ref > 1e-290
    np.testing.assert_allclose(K[nz], ref[nz], rtol=2e-12, atol=0.0)  # relative, even in the tail
    assert np.all(K[ref == 0.0] == 0.0)
    sub = (ref > 0) & (ref < 2.3e-308)
    assert sub.any()
    np.testing.assert_allclose(K[sub], ref[sub], rtol=1e-9, atol=5e-324)


def test_float32_dtypes_follow_the_reference():
    rng = np.random.default_rng(5)
    x0 = rng.normal(size=(200, 3)).astype(np.float32)
    x1 = rng.normal(size=(90, 3)).astype(np.float32)
    for spec in SPECS[:7:2]:
        got = make_kernel(spec).matrix(x0, x1)
        want = orc.gram(spec, x0, x1)
        assert got.dtype == want.dtype
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_stacked_and_left_operands():
    """(B, N1, K) operands and v @ L (linops/_linear_operator.py:1206-1254)."""
    rng = np.random.default_rng(9)
    spec = orc.KernelSpec("expquad", (4,), 1.2)
    k = make_kernel(spec)
    x0, x1 = rng.normal(size=(130, 4)), rng.normal(size=(70, 4))
    K = orc.gram(spec, x0, x1)
    L = k.linop(x0, x1)
    xs = rng.normal(size=(3, 70, 2))
    np.testing.assert_allclose(L @ xs, K @ xs, rtol=1e-11, atol=1e-12)
    w = rng.normal(size=(130,))
    np.testing.assert_allclose(w @ L, w @ K, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(L.T.todense(), K.T, rtol=1e-12, atol=1e-12)
    assert L.T.shape == (70, 130)
    # K + sigma^2 I as used by problinsolve callers
    Ls = k.linop(x0)
    A = Ls + 1e-3
    z = rng.normal(size=130)
    Ks = orc.gram(spec, x0)
    np.testing.assert_allclose(A @ z, Ks @ z + 1e-3 * z, rtol=1e-11, atol=1e-12)
    np.testing.assert_array_equal(Ls.diagonal(), np.ones(130))
    assert Ls.trace() == 130.0
    np.testing.assert_allclose(L.diagonal(), np.diag(K), rtol=1e-12, atol=1e-12)


def test_linop_equals_matrix_times_identity():
    """Mirror of test_keops_equals_no_keops (test_linop_matrix.py:141-154): the
    matrix-free product applied to I reproduces the dense path."""
    rng = np.random.default_rng(13)
    for spec in SPECS:
        k = make_kernel(spec)
        x0 = rng.normal(size=(97,) + spec.input_shape)
        x1 = rng.normal(size=(61,) + spec.input_shape)
        L = k.linop(x0, x1)
        np.testing.assert_allclose(L @ np.eye(61), k.matrix(x0, x1), rtol=1e-12, atol=1e-12)
        L.close()


def test_matvec_deterministic_and_split_columns():
    """Large enough to use several column splits; two runs must agree bit for bit."""
    rng = np.random.default_rng(17)
    spec = orc.KernelSpec("expquad", (4,), 0.6)
    k = make_kernel(spec)
    x = rng.normal(size=(3000, 4))
    v = rng.normal(size=(3000, 5))
    L = k.linop(x)
    y1 = L @ v
    y2 = L @ v
    np.testing.assert_array_equal(y1, y2)
    assert_matvec_close(y1, orc.matvec(spec, x, None, v), orc.abs_matvec(spec, x, None, v), 1e-12)
    pos = np.abs(v)
    np.testing.assert_allclose(L @ pos, orc.matvec(spec, x, None, pos), rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties + sampled rows against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec", [SPECS[0], SPECS[4], SPECS[6]], ids=["expquad", "matern52", "ratquad"])
def test_config2_gram_32768(spec):
    rng = np.random.default_rng(42)
    n = 32768
    x = rng.normal(size=(n, 3))
    K = make_kernel(spec).matrix(x)
    assert K.shape == (n, n) and K.dtype == np.float64
    rows = rng.choice(n, size=48, replace=False)
    np.testing.assert_allclose(K[rows], orc.gram(spec, x[rows], x), rtol=1e-12, atol=1e-12)
    assert np.all(np.diagonal(K) == 1.0)
    blk = slice(1000, 3000)
    np.testing.assert_array_equal(K[blk, 20000:22000], K[20000:22000, blk].T)  # symmetry
    assert np.all((K[rows] > 0) & (K[rows] <= 1.0))


def test_config4_matvec_65536():
    rng = np.random.default_rng(42)
    n = 65536
    spec = orc.KernelSpec("expquad", (4,), 1.0)
    x = rng.normal(size=(n, 4))
    v = rng.normal(size=n)
    L = make_kernel(spec).linop(x)
    y = L @ v
    rows = rng.choice(n, size=64, replace=False)
    Kr = orc.gram(spec, x[rows], x)
    assert_matvec_close(y[rows], Kr @ v, np.abs(Kr) @ np.abs(v), 1e-12)
    # linearity: K (a v + w) = a K v + K w
    w = rng.uniform(size=n)
    lhs = L @ (2.5 * v + w)
    rhs = 2.5 * y + L @ w
    np.testing.assert_allclose(lhs, rhs, rtol=1e-10, atol=1e-9)
    L.close()


# ---------------------------------------------------------------------------------------------
# symmetric mirror stores and the device-pointer entry points of the C ABI
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [512, 1024, 1537, 2048 + 64, 3000])
@pytest.mark.parametrize("spec", [SPECS[0], SPECS[4], SPECS[8]], ids=["expquad", "matern52", "ratquad-general"])
def test_symmetric_gram_is_bitwise_the_general_gram(spec, n):
    """x1=None takes the mirror-store kernel; passing the same points as x1 takes the general
    one.  K[i][j] and K[j][i] come from the same instruction sequence, so both must agree
    bit for bit, and with the oracle to 1e-12."""
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n,) + spec.input_shape)
    k = make_kernel(spec)
    Ks = k.matrix(x)
    Kg = k.matrix(x, x.copy())
    np.testing.assert_array_equal(Ks, Kg)
    np.testing.assert_array_equal(Ks, Ks.T)
    rows = rng.choice(n, size=40, replace=False)
    np.testing.assert_allclose(Ks[rows], orc.gram(spec, x[rows], x), rtol=1e-12, atol=1e-12)
    assert np.all(np.diagonal(Ks) == 1.0)


def test_device_pointer_entry_points():
    """pncov_gram_device / pncov_matmul_device on torch-owned memory and stream."""
    torch = pytest.importorskip("torch")
    from probnum_b200 import _lib

    rng = np.random.default_rng(23)
    n0, n1, d = 1100, 1500, 4
    spec = orc.KernelSpec("matern", (d,), 0.9, nu=1.5)
    x0, x1 = rng.normal(size=(n0, d)), rng.normal(size=(n1, d))
    dev = torch.device("cuda", 0)
    tx0, tx1 = torch.from_numpy(x0).to(dev), torch.from_numpy(x1).to(dev)
    op = _lib.Operator(make_kernel(spec)._kernel_params(), tx0.data_ptr(), tx1.data_ptr(), on_device=True,
                       n0=n0, n1=n1, dtype=np.float64, device=0)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        out = torch.full((n0, n1 + 6), -7.0, dtype=torch.float64, device=dev)  # padded leading dimension
        op.gram_device(out.data_ptr(), n1 + 6, stream=side.cuda_stream)
        v = torch.from_numpy(rng.normal(size=(n1, 3))).to(dev)
        y = torch.empty((n0, 3), dtype=torch.float64, device=dev)
        op.matmul_device(v.data_ptr(), 3, 3, y.data_ptr(), 3, stream=side.cuda_stream)
        outT = torch.empty((n1, n0), dtype=torch.float64, device=dev)
        op.gram_device(outT.data_ptr(), n0, transpose=True, stream=side.cuda_stream)
    side.synchronize()
    K = orc.gram(spec, x0, x1)
    got = out.cpu().numpy()
    np.testing.assert_allclose(got[:, :n1], K, rtol=1e-12, atol=1e-12)
    assert np.all(got[:, n1:] == -7.0)  # padding columns untouched
    np.testing.assert_allclose(outT.cpu().numpy(), K.T, rtol=1e-12, atol=1e-12)
    vh = v.cpu().numpy()
    assert_matvec_close(y.cpu().numpy(), K @ vh, np.abs(K) @ np.abs(vh), 1e-12)
    assert op.last_kernel_ms() > 0
    op.close()


def test_matvec_path_selection_and_parity_of_both_paths():
    """ExpQuad / RatQuad products use the norm-identity + DMMA kernel only when the host-side
    error bound allows it; a cloud that is wide in units of the lengthscale must fall back to
    direct differences.  Both paths have to meet the same componentwise bound."""
    from probnum_b200 import _lib

    rng = np.random.default_rng(31)
    n, d = 2500, 8
    x = rng.normal(size=(n, d)) + 300.0  # far from the origin: the centring has to cope
    v = rng.normal(size=(n, 2))
    for spec, want_path in [
        # x1 = None, two right-hand sides, n >= 2048: the symmetric multi-column form of the identity kernel
        (orc.KernelSpec("expquad", (d,), 1.0), "norm-identity-dmma-symmetric"),
        (orc.KernelSpec("ratquad", (d,), 1.0, alpha=1.7), "norm-identity-dmma-symmetric"),
        (orc.KernelSpec("expquad", (d,), 0.02), "direct-differences"),   # R^2 ~ 1e5 scaled units
        (orc.KernelSpec("matern", (d,), 1.0, nu=1.5), "direct-differences"),
    ]:
        k = make_kernel(spec)
        op = _lib.Operator(k._kernel_params(), x)
        assert op.matmul_path() == "undecided"
        y = op.matmul(v)
        assert op.matmul_path() == want_path, spec
        yt = op.matmul(v, transpose=True)
        ref = orc.matvec(spec, x, None, v)
        bound = orc.abs_matvec(spec, x, None, v)
        assert_matvec_close(y, ref, bound, 1e-12)
        assert_matvec_close(yt, ref, bound, 1e-12)
        op.close()


def test_dmma_matvec_rectangular_small_d_and_columns():
    rng = np.random.default_rng(37)
    for d, k_cols in [(1, 1), (3, 4), (4, 1), (5, 3), (8, 8)]:
        spec = orc.KernelSpec("expquad", (d,), 0.8 * np.sqrt(d))
        x0 = rng.normal(size=(700, d))
        x1 = rng.normal(size=(1301, d))
        v = rng.uniform(0.0, 1.0, size=(1301, k_cols))
        L = make_kernel(spec).linop(x0, x1)
        y = L @ v
        assert L._device_op().matmul_path() == "norm-identity-dmma"
        np.testing.assert_allclose(y, orc.matvec(spec, x0, x1, v), rtol=1e-12)  # v >= 0: plain rtol
        w = rng.normal(size=(700, k_cols))
        assert_matvec_close(L.T @ w, orc.matvec(spec, x1, x0, w), orc.abs_matvec(spec, x1, x0, w), 1e-12)
        L.close()


def test_repeated_product_is_served_from_the_memo():
    """Same operand twice in a row (what SymmetricMatrixBasedSolver does): identical result,
    no second launch; a different operand invalidates the memo."""
    from probnum_b200 import _lib

    rng = np.random.default_rng(41)
    x = rng.normal(size=(900, 4))
    L = pb.Matern((4,), nu=1.5).linop(x)
    v = rng.normal(size=(900, 1))
    y1 = L @ v
    before = _lib.launch_count()
    y2 = L @ v.copy()
    assert _lib.launch_count() == before
    np.testing.assert_array_equal(y1, y2)
    y2[0] = 123.0  # the caller owns the returned array
    np.testing.assert_array_equal(L @ v, y1)
    w = v + 1.0
    y3 = L @ w
    assert _lib.launch_count() > before
    spec = orc.KernelSpec("matern", (4,), 1.0, nu=1.5)
    assert_matvec_close(y3, orc.matvec(spec, x, None, w), orc.abs_matvec(spec, x, None, w), 1e-12)


def test_config3_matvec_2pow20():
    """BASELINE configs[2]: matrix-free ExpQuad K.v at N = 1,048,576, d = 8.  Sampled rows
    against the oracle with the componentwise bound, plus a checksum identity: 1^T (K v) =
    (K 1)^T v because K is symmetric."""
    rng = np.random.default_rng(42)
    n, d = 1 << 20, 8
    spec = orc.KernelSpec("expquad", (d,), 1.0)
    x = rng.normal(size=(n, d))
    v = rng.normal(size=n)
    L = make_kernel(spec).linop(x)
    y = L @ v
    assert L._device_op().matmul_path() == "norm-identity-dmma-symmetric"
    rows = rng.choice(n, size=24, replace=False)
    Kr = orc.gram(spec, x[rows], x, row_block=8)
    assert_matvec_close(y[rows], Kr @ v, np.abs(Kr) @ np.abs(v), 1e-12)
    ones = np.ones(n)
    k1 = L @ ones
    lhs, rhs = float(np.sum(y)), float(k1 @ v)
    scale = float(k1 @ np.abs(v))
    assert abs(lhs - rhs) <= 1e-11 * scale
    assert np.all(k1 >= 1.0)  # every row contains the unit diagonal
    L.close()


@pytest.mark.parametrize("k_cols", [5, 8, 9, 16])
def test_wide_right_hand_sides(k_cols):
    """More than four right-hand sides run 8 columns per pass in the norm-identity kernel
    (and 4 per pass in the direct kernel)."""
    rng = np.random.default_rng(k_cols)
    x0, x1 = rng.normal(size=(1000, 6)), rng.normal(size=(1400, 6))
    v = rng.normal(size=(1400, k_cols))
    for spec in (orc.KernelSpec("expquad", (6,), 1.3), orc.KernelSpec("matern", (6,), 1.3, nu=2.5)):
        L = make_kernel(spec).linop(x0, x1)
        assert_matvec_close(L @ v, orc.matvec(spec, x0, x1, v), orc.abs_matvec(spec, x0, x1, v), 1e-12)
        L.close()


# ---------------------------------------------------------------------------------------------
# symmetric halving of the matrix-free product (x1 is x0, one right-hand side)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,d,n,nu,ls", [
    ("expquad", 8, 5000, 0, 0.9), ("expquad", 3, 2177, 0, 0.9), ("ratquad", 4, 4096, 0, 0.9), ("ratquad", 8, 6001, 0, 0.9),
    ("matern", 8, 5000, 2.5, 0.9), ("matern", 3, 2305, 0.5, 0.9), ("matern", 5, 4096, 3.5, [0.5, 0.7, 0.9, 1.1, 1.3]),
    ("expquad", 6, 3000, 0, 0.01),  # wide cloud in lengthscale units: identity inadmissible -> direct differences
    ("matern", 12, 2200, 1.5, 3.0), ("matern", 40, 2100, 2.5, 6.0),  # 8 < d <= 96: row-staged symmetric kernel
])
def test_symmetric_matvec_matches_oracle(kind, d, n, nu, ls):
    rng = np.random.default_rng(n)
    spec = orc.KernelSpec(kind, (d,), np.asarray(ls) if np.ndim(ls) else ls, nu=nu or 1.5, alpha=1.5)
    k = make_kernel(spec)
    x = rng.normal(size=(n, d))
    x[17] = x[n - 5]  # a duplicated point far off the diagonal: sqrt(0) inside a mirrored tile
    L = k.linop(x)
    direct = kind == "matern" or np.ndim(ls) == 0 and ls < 0.1
    family = "direct-differences" if direct else "norm-identity-dmma"
    for v in (rng.normal(size=(n, 1)), rng.uniform(size=(n, 1))):
        y = L._matmul_device(v)
        assert L._device_op().matmul_path() == family + "-symmetric"
        assert_matvec_close(y, orc.matvec(spec, x, None, v), orc.abs_matvec(spec, x, None, v), 1e-12)
    # several columns: the symmetric multi-column form where the DMMA identity kernel serves (d <= 8), the ordered
    # kernel otherwise; either must agree with the one-column results
    V = rng.normal(size=(n, 3))
    Y = L._matmul_device(V)
    assert L._device_op().matmul_path() == (family + "-symmetric" if not direct and d <= 8 else family)
    y0 = L._matmul_device(V[:, :1])
    bound = orc.abs_matvec(spec, x, None, V[:, :1])
    assert np.all(np.abs(Y[:, :1] - y0) <= 1e-13 * bound)
    L.close()


def test_symmetric_matvec_parts_sum_to_the_product():
    """pncov_matmul_sym_part_device: the multi-GPU decomposition, here all parts on one device."""
    import torch

    from probnum_b200 import _lib

    n, d = 3333, 8
    rng = np.random.default_rng(3)
    x = rng.normal(size=(n, d))
    v = rng.normal(size=(n, 1))
    spec = orc.KernelSpec("expquad", (d,), 1.0)
    op = _lib.Operator(make_kernel(spec)._kernel_params(), x)
    vd = torch.from_numpy(v).cuda()
    total = torch.zeros((n, 1), dtype=torch.float64, device="cuda")
    part = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    for p in range(3):
        op.matmul_sym_part_device(vd.data_ptr(), part.data_ptr(), p, 3)
        torch.cuda.synchronize()
        total += part
    assert_matvec_close(total.cpu().numpy(), orc.matvec(spec, x, None, v), orc.abs_matvec(spec, x, None, v), 1e-12)
    op.close()
    # Matern takes the direct-difference form of the same decomposition
    specm = orc.KernelSpec("matern", (d,), 1.0, nu=1.5)
    opm = _lib.Operator(make_kernel(specm)._kernel_params(), x)
    total.zero_()
    for p in range(2):
        opm.matmul_sym_part_device(vd.data_ptr(), part.data_ptr(), p, 2)
        torch.cuda.synchronize()
        total += part
    assert_matvec_close(total.cpu().numpy(), orc.matvec(specm, x, None, v), orc.abs_matvec(specm, x, None, v), 1e-12)
    opm.close()
    # d > 96 has no symmetric kernel: the library says so and the caller shards rows instead
    x100 = rng.normal(size=(n, 100))
    op100 = _lib.Operator(pb.ExpQuad((100,), lengthscales=10.0)._kernel_params(), x100)
    with pytest.raises(NotImplementedError):
        op100.matmul_sym_part_device(vd.data_ptr(), part.data_ptr(), 0, 2)
    op100.close()
    # ... while 8 < d <= 32 has the wide DMMA form of the decomposition
    x20 = rng.normal(size=(n, 20))
    spec20 = orc.KernelSpec("expquad", (20,), 4.0)
    op20 = _lib.Operator(make_kernel(spec20)._kernel_params(), x20)
    total.zero_()
    for p in range(3):
        op20.matmul_sym_part_device(vd.data_ptr(), part.data_ptr(), p, 3)
        torch.cuda.synchronize()
        total += part
    assert_matvec_close(total.cpu().numpy(), orc.matvec(spec20, x20, None, v), orc.abs_matvec(spec20, x20, None, v), 1e-12)
    op20.close()


def test_symmetric_matvec_can_be_switched_off(monkeypatch):
    monkeypatch.setenv("PNCOV_MATVEC_SYM", "off")
    rng = np.random.default_rng(9)
    x = rng.normal(size=(2500, 5))
    v = rng.normal(size=(2500, 1))
    L = pb.ExpQuad((5,)).linop(x)
    y1 = L._matmul_device(v)
    assert L._device_op().matmul_path() == "norm-identity-dmma"
    np.testing.assert_array_equal(L._matmul_device(v), y1)  # the ordered kernel is bit-reproducible
    L.close()


def test_use_keops_false_selects_the_conservative_device_kernels():
    """The reference's tests switch its accelerated operator path off with `linop._use_keops = False`
    (test_linop_matrix.py:38-43).  Here the same attribute selects direct differences and a fixed
    summation order -- still on the device -- and the results agree with the default kernels."""
    rng = np.random.default_rng(21)
    x = rng.normal(size=(2600, 6))
    v = rng.normal(size=(2600, 1))
    spec = orc.KernelSpec("expquad", (6,), 1.1)
    L = make_kernel(spec).linop(x)
    y_fast = L._matmul_device(v)
    assert L._device_op().matmul_path() == "norm-identity-dmma-symmetric"
    L._use_keops = False
    y_cons = L._matmul_device(v)
    assert L._device_op().matmul_path() == "direct-differences"
    np.testing.assert_array_equal(L._matmul_device(v), y_cons)  # order-preserving: bit-reproducible
    bound = orc.abs_matvec(spec, x, None, v)
    assert_matvec_close(y_cons, orc.matvec(spec, x, None, v), bound, 1e-12)
    assert np.all(np.abs(y_fast - y_cons) <= 1e-12 * bound)
    L._use_keops = True
    L._matmul_device(v)
    assert L._device_op().matmul_path() == "norm-identity-dmma-symmetric"
    L.close()


@pytest.mark.parametrize("kind,d,n", [("expquad", 9, 2100), ("expquad", 16, 3000), ("expquad", 17, 2500), ("expquad", 32, 2304),
                                      ("ratquad", 12, 2600), ("ratquad", 29, 2200), ("expquad", 33, 2100), ("expquad", 64, 2300),
                                      ("ratquad", 50, 2200)])
def test_wide_dmma_matvec_8_lt_d_le_64(kind, d, n):
    """8 < d <= 64: the norm-identity kernels with K-chunked DMMA contraction (one right-hand side),
    symmetric and rectangular, against the oracle; several columns fall back to the row-staged kernel."""
    rng = np.random.default_rng(d)
    ls = 0.9 * np.sqrt(d)
    spec = orc.KernelSpec(kind, (d,), ls, alpha=1.3)
    k = make_kernel(spec)
    x = rng.normal(size=(n, d))
    x1 = rng.normal(size=(n // 2 + 3, d))
    v = rng.normal(size=(n, 1))
    L = k.linop(x)
    y = L._matmul_device(v)
    assert L._device_op().matmul_path() == "norm-identity-dmma-symmetric"
    assert_matvec_close(y, orc.matvec(spec, x, None, v), orc.abs_matvec(spec, x, None, v), 1e-12)
    Y = L._matmul_device(np.concatenate([v, 2 * v], axis=1))
    assert L._device_op().matmul_path() == "direct-differences"
    assert np.all(np.abs(Y[:, :1] - y) <= 1e-12 * orc.abs_matvec(spec, x, None, v))
    L.close()
    Lr = k.linop(x, x1)
    w = rng.normal(size=(x1.shape[0], 1))
    yr = Lr._matmul_device(w)
    assert Lr._device_op().matmul_path() == "norm-identity-dmma"
    assert_matvec_close(yr, orc.matvec(spec, x, x1, w), orc.abs_matvec(spec, x, x1, w), 1e-12)
    yt = Lr.T @ v
    assert_matvec_close(yt, orc.matvec(spec, x1, x, v), orc.abs_matvec(spec, x1, x, v), 1e-12)
    Lr.close()
