"""The symmetric-halving products accumulate the partial sums of their CTAs order-independently
(csrc/pncov_common.cuh: pn_bin_add -- exact integer planes on a per-component power-of-two grid), so
the DEFAULT K @ v is bit-for-bit reproducible like the reference's NumPy product
(_covariance_linear_operator.py:78-81), for every kernel family of the symmetric path and for every
split of the row blocks over ranks."""

import numpy as np
import pytest
import torch

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc
from probnum_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [
    ("expquad", 8, dict(lengthscales=1.0), "norm-identity-dmma-symmetric"),      # factorised ExpQuad on DMMA
    ("ratquad", 5, dict(lengthscale=1.3, alpha=1.7), "norm-identity-dmma-symmetric"),
    ("matern", 3, dict(nu=2.5, lengthscales=0.7), "direct-differences-symmetric"),
    ("expquad", 20, dict(lengthscales=4.0), "norm-identity-dmma-symmetric"),     # wide DMMA operands
    ("matern", 40, dict(nu=1.5, lengthscales=5.0), "direct-differences-symmetric"),  # row-staged kernel
]


def kernel(kind, d, kw):
    return {"expquad": pb.ExpQuad, "matern": pb.Matern, "ratquad": pb.RatQuad}[kind]((d,), **kw)


def spec_of(kind, d, kw):
    return orc.KernelSpec(kind, (d,), kw.get("lengthscales", kw.get("lengthscale", 1.0)), nu=kw.get("nu", 1.5),
                          alpha=kw.get("alpha", 1.0))


@pytest.mark.parametrize("kind,d,kw,path", CASES, ids=[f"{c[0]}-d{c[1]}" for c in CASES])
def test_default_product_is_bit_reproducible_and_rank_count_independent(kind, d, kw, path):
    n = 6000
    rng = np.random.default_rng(17)
    x = rng.normal(size=(n, d))
    v = rng.normal(size=(n, 1)) * np.exp(rng.normal(size=(n, 1)) * 3.0)  # wide dynamic range in v
    k = kernel(kind, d, kw)
    op = _lib.Operator(k._kernel_params(), x)
    y0 = op.matmul(v)
    assert op.matmul_path() == path
    for _ in range(4):
        assert np.array_equal(op.matmul(v), y0)
    spec = spec_of(kind, d, kw)
    bound = orc.abs_matvec(spec, x, None, v)
    assert np.all(np.abs(y0 - orc.matvec(spec, x, None, v)) <= 1e-12 * bound)
    # the row blocks dealt to 2, 3 and 8 "ranks" (all on this device): integer planes summed, then finished
    vd = torch.from_numpy(v).cuda()
    for nparts in (2, 3, 8):
        total = torch.zeros(3 * n, dtype=torch.int64, device="cuda")
        bins = torch.empty(3 * n, dtype=torch.int64, device="cuda")
        for p in range(nparts):
            op.matmul_sym_part_bins_device(vd.data_ptr(), bins.data_ptr(), p, nparts)
            torch.cuda.synchronize()
            total[: 2 * n] += bins[: 2 * n]
            total[2 * n:].view(torch.float64).add_(bins[2 * n:].view(torch.float64))
        y = torch.empty((n, 1), dtype=torch.float64, device="cuda")
        op.matmul_sym_finish_device(total.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(y.cpu().numpy(), y0), nparts
    op.close()


def test_non_finite_right_hand_sides_propagate_like_numpy():
    n, d = 4096, 4
    rng = np.random.default_rng(5)
    x = rng.normal(size=(n, d))
    k = pb.Matern((d,), nu=1.5)
    K = k.matrix(x)
    op = _lib.Operator(k._kernel_params(), x)
    for bad in (np.inf, -np.inf, np.nan):
        v = rng.normal(size=(n, 1))
        v[100, 0] = bad
        y = op.matmul(v)
        with np.errstate(invalid="ignore"):
            want = K @ v
        assert np.array_equal(np.isnan(y), np.isnan(want))
        assert np.array_equal(np.isinf(y), np.isinf(want))
        assert np.array_equal(np.sign(y[np.isinf(y)]), np.sign(want[np.isinf(want)]))
    v = rng.normal(size=(n, 1))
    v[7, 0], v[9, 0] = np.inf, -np.inf
    assert np.all(np.isnan(op.matmul(v)))  # Inf - Inf in every row, as in NumPy
    assert np.array_equal(op.matmul(np.zeros((n, 1))), np.zeros((n, 1)))
    op.close()


def test_tiny_and_huge_right_hand_sides():
    n, d = 3000, 8
    rng = np.random.default_rng(6)
    x = rng.normal(size=(n, d))
    spec = orc.KernelSpec("expquad", (d,), 1.0)
    op = _lib.Operator(pb.ExpQuad((d,))._kernel_params(), x)
    for scale in (1e-290, 1e-150, 1.0, 1e150, 1e290):
        v = rng.normal(size=(n, 1)) * scale
        y = op.matmul(v)
        bound = orc.abs_matvec(spec, x, None, v)
        assert np.all(np.abs(y - orc.matvec(spec, x, None, v)) <= 1e-12 * bound + 1e-320), scale
        assert np.array_equal(op.matmul(v), y)
    op.close()


@pytest.mark.parametrize("kind,d,kw", [("expquad", 8, dict(lengthscales=1.0)), ("expquad", 3, dict(lengthscales=0.6)),
                                       ("ratquad", 6, dict(lengthscale=1.3, alpha=1.7))])
@pytest.mark.parametrize("k", [2, 3, 4, 5, 8, 11, 16])
def test_symmetric_products_with_several_right_hand_sides(kind, d, kw, k):
    """x1 = None and 2 .. 16 right-hand sides: symmetric passes of eight columns with both accumulations on the
    tensor cores (matvec_mma_symd_kernel); same parity bound, bit-reproducible, consistent with the single-column product."""
    n = 5000
    rng = np.random.default_rng(23 + k)
    x = rng.normal(size=(n, d))
    V = rng.uniform(size=(n, k)) * np.exp(rng.normal(size=(1, k)) * 2.0)
    op = _lib.Operator(kernel(kind, d, kw)._kernel_params(), x)
    Y = op.matmul(V)
    assert op.matmul_path() == "norm-identity-dmma-symmetric"
    spec = spec_of(kind, d, kw)
    bound = orc.abs_matvec(spec, x, None, V)
    assert np.all(np.abs(Y - orc.matvec(spec, x, None, V)) <= 1e-12 * bound)
    assert np.array_equal(op.matmul(V), Y)
    y0 = op.matmul(V[:, :1])
    assert np.all(np.abs(Y[:, :1] - y0) <= 1e-13 * bound[:, :1])
    op.close()


@pytest.mark.parametrize("k", [2, 8])
def test_multi_column_parts_finish_to_the_bits_of_one_part(k):
    """pncov_matmulk_sym_part_bins_device: the multi-GPU decomposition for 2 .. 8 columns, here all parts on one
    device; the summed integer planes finish to exactly the one-part product."""
    n, d = 4500, 8
    rng = np.random.default_rng(41)
    x = rng.normal(size=(n, d))
    V = rng.normal(size=(n, k))
    op = _lib.Operator(pb.ExpQuad((d,))._kernel_params(), x)
    Y0 = op.matmul(V)
    Vd = torch.from_numpy(V).cuda()
    for nparts in (2, 5):
        total = torch.zeros(24 * n, dtype=torch.int64, device="cuda")
        bins = torch.empty(24 * n, dtype=torch.int64, device="cuda")
        for p in range(nparts):
            op.matmulk_sym_part_bins_device(Vd.data_ptr(), k, k, bins.data_ptr(), p, nparts)
            torch.cuda.synchronize()
            total[: 16 * n] += bins[: 16 * n]
            total[16 * n:].view(torch.float64).add_(bins[16 * n:].view(torch.float64))
        Y = torch.empty((n, k), dtype=torch.float64, device="cuda")
        op.matmulk_sym_finish_device(total.data_ptr(), k, Y.data_ptr(), k)
        torch.cuda.synchronize()
        assert np.array_equal(Y.cpu().numpy(), Y0), nparts
    # Matern has no multi-column symmetric form: the library says so and the caller shards rows instead
    opm = _lib.Operator(pb.Matern((d,), nu=1.5)._kernel_params(), x)
    with pytest.raises(NotImplementedError):
        opm.matmulk_sym_part_bins_device(Vd.data_ptr(), k, k, bins.data_ptr(), 0, 2)
    op.close()
    opm.close()
