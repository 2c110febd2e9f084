"""CPU oracle for the remaining covariance functions and covariance arithmetic.
TEST INFRASTRUCTURE ONLY -- same rules as ``covfuncs_oracle.py``: nothing in ``probnum_b200``
imports this module.

Restates, in NumPy and in the reference's order of operations, ``_evaluate`` of

    Linear         randprocs/covfuncs/_linear.py:49-50   (+ _euclidean_inner_products,
                   _covariance_function.py:695-707)
    Polynomial     _polynomial.py:55-56
    WhiteNoise     _white_noise.py:41-55
    ProductMatern  _product_matern.py:103-122 (one Matern(input_shape=()) per dimension)
    Scaled / Sum / Product covariance functions   _arithmetic_fallbacks.py:55-56, :101-104, :163-166

Parity status: PINNED.  ``tests/golden/make_golden_extra.py`` imports the real reference and
stores its outputs in ``tests/golden/covfuncs_extra_golden.{npz,json}``; ``tests/test_oracle_extra.py``
checks this module bit-for-bit against them.

A kernel is described by a plain dict (JSON-able), e.g.
``{"kind": "sum", "children": [{"kind": "scaled", "scalar": 2.0, "child": {...}}, {...}]}``;
leaves of kind expquad / matern / ratquad use the fields of ``covfuncs_oracle.KernelSpec``.
"""

from __future__ import annotations

import functools
import operator

import numpy as np

from . import covfuncs_oracle as base

ISOTROPIC = ("expquad", "matern", "ratquad")


def input_shape(spec) -> tuple:
    kind = spec["kind"]
    if kind == "scaled":
        return input_shape(spec["child"])
    if kind in ("sum", "product"):
        return input_shape(spec["children"][0])
    return tuple(spec["input_shape"])


def _iso_spec(spec) -> base.KernelSpec:
    ls = spec.get("lengthscales", 1.0)
    return base.KernelSpec(kind=spec["kind"], input_shape=tuple(spec["input_shape"]),
                           lengthscales=np.asarray(ls) if np.ndim(ls) else float(ls),
                           nu=float(spec.get("nu", 1.5)), alpha=float(spec.get("alpha", 1.0)))


def _inner_products(nd: int, x0, x1):
    """_covariance_function.py:695-707."""
    prods = x0**2 if x1 is None else x0 * x1
    if nd == 0:
        return prods
    return np.sum(prods, axis=tuple(range(-nd, 0)))


def evaluate(spec, x0, x1):
    """k(x0, x1) with NumPy broadcasting over batch shapes (x1 None: x1 = x0).

    Hyper-parameters are 0-d arrays of inferred dtype, exactly what the reference's
    ``utils.as_numpy_scalar`` (utils/argutils.py) stores: ``x ** array(3)`` takes NumPy's array
    power loop where ``x ** np.float64(3)`` would take the scalar one (1 ulp apart)."""
    x0 = np.asarray(x0)
    if x1 is not None:
        x1 = np.asarray(x1)
    kind = spec["kind"]
    if kind in ISOTROPIC:
        return base.evaluate(_iso_spec(spec), x0, x1)
    if kind == "scaled":  # _arithmetic_fallbacks.py:55-56
        return np.asarray(spec["scalar"]) * evaluate(spec["child"], x0, x1)
    if kind == "sum":  # :101-104
        return functools.reduce(operator.add, (evaluate(c, x0, x1) for c in spec["children"]))
    if kind == "product":  # :163-166
        return functools.reduce(operator.mul, (evaluate(c, x0, x1) for c in spec["children"]))
    shape = tuple(spec["input_shape"])
    nd = len(shape)
    if kind == "linear":  # _linear.py:49-50
        return _inner_products(nd, x0, x1) + np.asarray(spec["constant"])
    if kind == "polynomial":  # _polynomial.py:55-56
        return (_inner_products(nd, x0, x1) + np.asarray(spec["constant"])) ** np.asarray(spec["exponent"])
    if kind == "whitenoise":  # _white_noise.py:41-55
        sigma_sq = np.asarray(spec["sigma_sq"])
        if x1 is None:
            return np.full_like(x0, sigma_sq, shape=x0.shape[: x0.ndim - nd])
        if shape == ():
            return sigma_sq * (x0 == x1)
        return sigma_sq * np.all(x0 == x1, axis=-1)
    if kind == "productmatern":  # _product_matern.py:103-122
        ls = np.broadcast_to(np.asarray(spec["lengthscales"], dtype=np.double), (max(1, int(np.prod(shape))),))
        nus = np.broadcast_to(np.asarray(spec["nus"], dtype=np.double), ls.shape)
        uni = [base.KernelSpec("matern", (), float(ls[i]), nu=float(nus[i])) for i in range(ls.size)]
        if shape == ():
            return base.evaluate(uni[0], x0, x1)
        k = 1.0
        for dim in range(shape[0]):
            k = k * base.evaluate(uni[dim], x0[..., dim], None if x1 is None else x1[..., dim])
        return k
    raise ValueError(f"unknown kind {kind!r}")


def gram(spec, x0, x1=None, row_block: int = 1024) -> np.ndarray:
    """Dense Gram matrix (N0, N1): ``self(x0[:, None], x1[None, :])`` of
    CovarianceFunction._evaluate_matrix (_covariance_function.py:553-582) on row blocks."""
    shape = input_shape(spec)
    a = np.asarray(x0).reshape((-1,) + shape, order="C")
    b = a if x1 is None else np.asarray(x1).reshape((-1,) + shape, order="C")
    n0, n1 = a.shape[0], b.shape[0]
    bcols = b[None, ...]
    if n0 == 0:
        return np.empty((0, n1), dtype=np.result_type(a, b, np.float64))
    first = evaluate(spec, a[:1, None, ...], bcols)
    out = np.empty((n0, n1), dtype=np.asarray(first).dtype)
    for lo in range(0, n0, row_block):
        hi = min(lo + row_block, n0)
        out[lo:hi] = evaluate(spec, a[lo:hi, None, ...], bcols)
    return out
