"""CPU oracle for the covariance-function hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy, the arithmetic that ProbNum's reference
implementation performs for the isotropic covariance functions ExpQuad, Matern
(half-integer nu) and RatQuad, for the dense Gram matrix and for the
matrix-vector product.  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing in ``probnum_b200`` imports it; the product
path raises when the CUDA library is missing instead of falling back to this.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
(``/root/reference/src/probnum``) in the build container and stores its outputs as
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this module bit-for-bit
against those files and against the doctest known-answer values of the reference.

All ``file:line`` citations are relative to ``/root/reference/src/probnum``.

The order of floating-point operations deliberately follows the reference (same
ufunc sequence, same reduction axis layout), so the outputs are bit-identical to
the reference's on NumPy >= 2, including its dtype behaviour for float32 inputs.
"""

from __future__ import annotations

import fractions
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

EXPQUAD = "expquad"
MATERN = "matern"
RATQUAD = "ratquad"


@dataclass(frozen=True)
class KernelSpec:
    """Hyper-parameters of one covariance function.

    kind          "expquad" | "matern" | "ratquad"
    input_shape   () or (d,)
    lengthscales  scalar or array broadcastable to input_shape (RatQuad: scalar only)
    nu            Matern smoothness (half-integer only in this oracle)
    alpha         RatQuad scale mixture
    """

    kind: str
    input_shape: Tuple[int, ...] = ()
    lengthscales: object = 1.0
    nu: float = 1.5
    alpha: float = 1.0

    @property
    def input_ndim(self) -> int:
        return len(self.input_shape)


def matern_p(nu: float) -> Optional[int]:
    """Degree p with nu = p + 1/2, else None.  randprocs/covfuncs/_matern.py:136-151."""
    shifted = float(nu) - 0.5
    return int(shifted) if shifted == int(shifted) else None


def matern_coefficients(p: int) -> np.ndarray:
    """Exact rational polynomial coefficients c_0..c_p converted to float64.

    Recursion c_{p-i} = 2(i+1) / ((p+i+1)(p-i)) * c_{p-(i+1)} with c_0 = 1,
    randprocs/covfuncs/_matern.py:225-267.  Index i holds the degree-i coefficient.
    """
    acc = fractions.Fraction(1)
    out = [acc]
    for i in reversed(range(p)):
        acc = acc * 2 * (i + 1) / (p + i + 1) / (p - i)
        out.append(acc)
    return np.asarray(out, dtype=np.float64)


def scale_factors(spec: KernelSpec) -> Optional[np.ndarray]:
    """Per-dimension factor applied to x0 - x1 before squaring.

    ExpQuad: sqrt(1/2)/l   (_exponentiated_quadratic.py:85-87)
    Matern : sqrt(2 nu)/l  (_matern.py:171-173)
    RatQuad: none          (_rational_quadratic.py:78-84 passes no scale_factors)
    """
    if spec.kind == EXPQUAD:
        return np.sqrt(0.5) / np.asarray(spec.lengthscales, dtype=np.double)
    if spec.kind == MATERN:
        nu = np.float64(spec.nu)
        return np.sqrt(2 * nu) / np.asarray(spec.lengthscales, dtype=np.double)
    if spec.kind == RATQUAD:
        return None
    raise ValueError(f"unknown kind {spec.kind!r}")


def squared_distances(
    spec: KernelSpec, x0: np.ndarray, x1: np.ndarray, scale: Optional[np.ndarray]
) -> np.ndarray:
    """Direct-difference squared distance, broadcasting over batch dimensions.

    _covariance_function.py:757-781: difference, in-place scaling, in-place square,
    sum over the trailing input dimensions (nothing to sum for scalar inputs).
    """
    diff = x0 - x1
    if scale is not None:
        diff *= scale
    diff *= diff
    return np.sum(diff, axis=tuple(range(-spec.input_ndim, 0)))


def _batch_shape(spec: KernelSpec, x: np.ndarray) -> Tuple[int, ...]:
    return x.shape[: x.ndim - spec.input_ndim]


def evaluate(spec: KernelSpec, x0, x1) -> np.ndarray:
    """k(x0, x1) with NumPy broadcasting over batch shapes; x1=None means x1 = x0.

    Mirrors CovarianceFunction.__call__ -> subclass._evaluate
    (_covariance_function.py:280-357).
    """
    x0 = np.asarray(x0)
    if x1 is not None:
        x1 = np.asarray(x1)

    if spec.kind == EXPQUAD:
        # _exponentiated_quadratic.py:89-100
        if x1 is None:
            return np.ones_like(x0, shape=_batch_shape(spec, x0))
        return np.exp(-squared_distances(spec, x0, x1, scale_factors(spec)))

    if spec.kind == RATQUAD:
        # _rational_quadratic.py:71-84
        if x1 is None:
            return np.ones_like(x0, shape=_batch_shape(spec, x0))
        ell = np.float64(spec.lengthscales)
        alpha = np.float64(spec.alpha)
        base = 1.0 + squared_distances(spec, x0, x1, None) / (2.0 * alpha * ell**2)
        return base**-alpha

    if spec.kind == MATERN:
        # _matern.py:175-193 with _euclidean_distances (_covariance_function.py:783-802)
        p = matern_p(spec.nu)
        if p is None:
            raise NotImplementedError("oracle covers half-integer nu only")
        if x1 is None:
            r = np.zeros_like(x0, shape=_batch_shape(spec, x0))
        else:
            r = np.sqrt(squared_distances(spec, x0, x1, scale_factors(spec)))
        c = matern_coefficients(p)
        acc = c[p]
        for i in reversed(range(p)):
            acc *= r
            acc += c[i]
        acc *= np.exp(-r)
        return acc

    raise ValueError(f"unknown kind {spec.kind!r}")


def _flatten_points(spec: KernelSpec, x) -> np.ndarray:
    """(batch..., *input_shape) -> (N, *input_shape); _covariance_function.py:676-693."""
    x = np.asarray(x)
    return x.reshape((-1,) + tuple(spec.input_shape), order="C")


def gram(spec: KernelSpec, x0, x1=None, row_block: int = 1024) -> np.ndarray:
    """Dense Gram matrix K[i, j] = k(x0_i, x1_j), shape (N0, N1).

    Same arithmetic as CovarianceFunction._evaluate_matrix
    (_covariance_function.py:553-582: ``self(x0[:, None], x1[None, :])``), but
    evaluated on row blocks so the (rows, N1, d) temporary stays bounded.  Row
    blocking does not change any entry: every entry is an independent sequence of
    elementwise ufuncs and a reduction over its own d-vector.
    """
    a = _flatten_points(spec, x0)
    b = a if x1 is None else _flatten_points(spec, x1)
    n0, n1 = a.shape[0], b.shape[0]
    bcols = b[None, ...]
    first = evaluate(spec, a[:1, None, ...], bcols) if n0 else None
    out = np.empty((n0, n1), dtype=first.dtype if n0 else np.result_type(a, b))
    for lo in range(0, n0, row_block):
        hi = min(lo + row_block, n0)
        out[lo:hi] = evaluate(spec, a[lo:hi, None, ...], bcols)
    return out


def matvec(spec: KernelSpec, x0, x1, v, row_block: int = 1024) -> np.ndarray:
    """y = K(x0, x1) @ v without holding K: row blocks of :func:`gram` times v.

    The reference's non-KeOps operator does ``self.todense() @ x``
    (_covariance_linear_operator.py:78-81); per row block this is the same product.
    """
    a = _flatten_points(spec, x0)
    b = a if x1 is None else _flatten_points(spec, x1)
    v = np.asarray(v)
    out = np.empty((a.shape[0],) + v.shape[1:], dtype=np.result_type(a, b, v))
    for lo in range(0, a.shape[0], row_block):
        hi = min(lo + row_block, a.shape[0])
        out[lo:hi] = gram(spec, a[lo:hi], b, row_block=row_block) @ v
    return out


def abs_matvec(spec: KernelSpec, x0, x1, v, row_block: int = 1024) -> np.ndarray:
    """(|K| |v|)_i, the scale of the componentwise K.v error bound."""
    return matvec(spec, x0, x1, np.abs(np.asarray(v)), row_block=row_block)
