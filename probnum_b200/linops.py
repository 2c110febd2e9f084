"""Minimal mirror of the ``probnum.linops`` interface that the covariance path touches.

Only the contract that ``CovarianceLinearOperator`` and its consumers rely on is
reproduced (/root/reference/src/probnum/linops/_linear_operator.py): ``shape``,
``dtype``, ``@`` with (N,), (N, K) and stacked (..., N, K) operands (:1170-1227),
``v @ L`` through the transpose (:1229-1254), ``.T`` (:938-951), ``todense(cache)``
(:385-411), ``diagonal()`` / ``trace()`` (:741-790) and the lazy ``+`` / scalar ``*``
arithmetic that ``problinsolve`` style callers use to form ``K + sigma^2 I``
(linops/_arithmetic.py:445-457, _arithmetic_fallbacks.py:20-161).  Everything here is
host-side plumbing around the device operator; no kernel arithmetic happens in it.
"""

from __future__ import annotations

import abc
import numbers
from typing import Optional, Tuple

import numpy as np


class LinearOperator(abc.ABC):
    """Abstract (M, N) linear operator; subclasses implement ``_matmul``."""

    # make ``ndarray @ operator`` / ``ndarray + operator`` defer to the reflected methods
    __array_ufunc__ = None

    def __init__(self, shape: Tuple[int, int], dtype):
        if len(shape) != 2:
            raise ValueError(f"shape must be 2-d, got {shape}")
        self._shape = (int(shape[0]), int(shape[1]))
        self._dtype = np.dtype(dtype)
        self._todense_cache: Optional[np.ndarray] = None
        self._diagonal_cache: Optional[np.ndarray] = None
        self._trace_cache = None
        self.is_symmetric: Optional[bool] = None

    # -- basic properties ---------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, int]:
        return self._shape

    @property
    def ndim(self) -> int:
        return 2

    @property
    def size(self) -> int:
        return self._shape[0] * self._shape[1]

    @property
    def dtype(self) -> np.dtype:
        return self._dtype

    @property
    def is_square(self) -> bool:
        return self._shape[0] == self._shape[1]

    def __repr__(self) -> str:
        return f"<{self.__class__.__name__} with shape={self.shape} and dtype={self.dtype}>"

    # -- products -----------------------------------------------------------------------------
    @abc.abstractmethod
    def _matmul(self, x: np.ndarray) -> np.ndarray:
        """``self @ x`` for ``x`` of shape (..., N, K); returns (..., M, K)."""

    def __matmul__(self, other):
        if isinstance(other, np.ndarray):
            M, N = self.shape
            x = other
            if x.ndim == 1 and x.shape == (N,):
                y = self._matmul(x[:, np.newaxis])[:, 0]
                assert y.shape == (M,)
            elif x.ndim > 1 and x.shape[-2] == N:
                y = self._matmul(x)
                assert y.shape == x.shape[:-2] + (M, x.shape[-1])
            else:
                raise ValueError(
                    f"Dimension mismatch. Expected operand of shape ({N},) or "
                    f"(..., {N}, K), but got {x.shape}."
                )
            return y
        if isinstance(other, LinearOperator):
            return ProductLinearOperator(self, other)
        return NotImplemented

    def __rmatmul__(self, other):
        if isinstance(other, np.ndarray):
            M, N = self.shape
            x = other
            if x.ndim >= 1 and x.shape[-1] == M:
                y = self.T(x, axis=-1)
            else:
                raise ValueError(
                    f"Dimension mismatch. Expected operand of shape (..., {M}), but got {x.shape}."
                )
            assert y.shape == x.shape[:-1] + (N,)
            return y
        return NotImplemented

    def __call__(self, x: np.ndarray, axis: Optional[int] = None) -> np.ndarray:
        """Apply the operator along ``axis`` of ``x`` (default: like ``@``)."""
        if axis is None:
            return self @ x
        x = np.asarray(x)
        axis = axis if axis >= 0 else axis + x.ndim
        if x.shape[axis] != self.shape[1]:
            raise ValueError(f"Dimension mismatch along axis {axis}.")
        moved = np.moveaxis(x, axis, -1)[..., np.newaxis]  # (..., N, 1)
        y = self._matmul(np.ascontiguousarray(moved))[..., 0]
        return np.moveaxis(y, -1, axis)

    # -- dense / structure ----------------------------------------------------------------------
    def _todense(self) -> np.ndarray:
        return self @ np.eye(self.shape[1], dtype=self.dtype)

    def todense(self, cache: bool = True) -> np.ndarray:
        if self._todense_cache is None:
            dense = self._todense()
            if not cache:
                return dense
            self._todense_cache = dense
        return self._todense_cache

    def _transpose(self) -> "LinearOperator":
        return TransposedLinearOperator(self)

    @property
    def T(self) -> "LinearOperator":
        if self.is_symmetric:
            return self
        t = self._transpose()
        t.is_symmetric = self.is_symmetric
        return t

    def transpose(self) -> "LinearOperator":
        return self.T

    def _diagonal(self) -> np.ndarray:
        n = min(self.shape)
        diag = np.zeros(n, dtype=self.dtype)
        vec = np.zeros(self.shape[1], dtype=self.dtype)
        for i in range(n):
            vec[i] = 1
            diag[i] = (self @ vec)[i]
            vec[i] = 0
        return diag

    def diagonal(self) -> np.ndarray:
        if self._diagonal_cache is None:
            self._diagonal_cache = self._diagonal()
            self._diagonal_cache.setflags(write=False)
        return self._diagonal_cache

    def _trace(self):
        return np.sum(self.diagonal())

    def trace(self):
        if self._trace_cache is None:
            if not self.is_square:
                raise np.linalg.LinAlgError("The trace is only defined for square operators.")
            self._trace_cache = self._trace()
        return self._trace_cache

    # -- lazy arithmetic --------------------------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, numbers.Number) or (isinstance(other, np.ndarray) and other.ndim == 0):
            # the reference turns a scalar summand into Scaling(c) = c * I (_arithmetic.py:445-457)
            if not self.is_square:
                raise ValueError("scalar + operator needs a square operator")
            other = Scaling(other, self.shape, dtype=self.dtype)
        if isinstance(other, np.ndarray):
            other = Matrix(other)
        if isinstance(other, LinearOperator):
            return SumLinearOperator(self, other)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-1.0) * other if isinstance(other, LinearOperator) else self + (-other)

    def __mul__(self, other):
        if isinstance(other, numbers.Number) or (isinstance(other, np.ndarray) and other.ndim == 0):
            return ScaledLinearOperator(self, other)
        return NotImplemented

    __rmul__ = __mul__

    def __neg__(self):
        return ScaledLinearOperator(self, -1.0)


class TransposedLinearOperator(LinearOperator):
    def __init__(self, linop: LinearOperator):
        self._linop = linop
        super().__init__((linop.shape[1], linop.shape[0]), linop.dtype)

    def _matmul(self, x):
        return self._linop.todense().T @ x

    def _todense(self):
        return self._linop.todense(cache=False).T.copy(order="C")

    def _transpose(self):
        return self._linop


class Matrix(LinearOperator):
    """Dense matrix as an operator (linops/_linear_operator.py:1602)."""

    def __init__(self, A):
        self.A = np.asarray(A)
        super().__init__(self.A.shape, self.A.dtype)

    def _matmul(self, x):
        return self.A @ x

    def _todense(self):
        return self.A

    def _transpose(self):
        return Matrix(self.A.T)

    def _diagonal(self):
        return np.diagonal(self.A).copy()


class Scaling(LinearOperator):
    """``c * I`` (linops/_scaling.py:16)."""

    def __init__(self, factor, shape, dtype=np.double):
        if isinstance(shape, numbers.Integral):
            shape = (shape, shape)
        if shape[0] != shape[1]:
            raise ValueError("Scaling needs a square shape")
        self.scalar = np.asarray(factor, dtype=np.result_type(factor, dtype))[()]
        super().__init__(shape, np.result_type(self.scalar, dtype))
        self.is_symmetric = True

    def _matmul(self, x):
        return self.scalar * x

    def _todense(self):
        return self.scalar * np.eye(self.shape[0], dtype=self.dtype)

    def _transpose(self):
        return self

    def _diagonal(self):
        return np.full(self.shape[0], self.scalar, dtype=self.dtype)


class Identity(Scaling):
    """Identity operator (linops/_linear_operator.py:1677)."""

    def __init__(self, shape, dtype=np.double):
        super().__init__(1.0, shape, dtype=dtype)

    def _matmul(self, x):
        return x.astype(np.result_type(self.dtype, x.dtype), copy=False)


class ScaledLinearOperator(LinearOperator):
    """``c * A`` (linops/_arithmetic_fallbacks.py:20-86)."""

    def __init__(self, linop: LinearOperator, scalar):
        self._linop = linop
        self._scalar = np.asarray(scalar)[()]
        super().__init__(linop.shape, np.result_type(linop.dtype, self._scalar))
        self.is_symmetric = linop.is_symmetric

    def _matmul(self, x):
        return self._scalar * (self._linop @ x)

    def _todense(self):
        return self._scalar * self._linop.todense(cache=False)

    def _transpose(self):
        return ScaledLinearOperator(self._linop.T, self._scalar)

    def _diagonal(self):
        return self._scalar * self._linop.diagonal()


class SumLinearOperator(LinearOperator):
    """``A + B + ...`` (linops/_arithmetic_fallbacks.py:89-175)."""

    def __init__(self, *summands: LinearOperator):
        flat = []
        for s in summands:
            flat.extend(s._summands if isinstance(s, SumLinearOperator) else [s])
        if any(s.shape != flat[0].shape for s in flat):
            raise ValueError("All summands must have the same shape.")
        self._summands = tuple(flat)
        super().__init__(flat[0].shape, np.result_type(*[s.dtype for s in flat]))
        if all(s.is_symmetric for s in flat):
            self.is_symmetric = True

    def _matmul(self, x):
        total = self._summands[0] @ x
        for s in self._summands[1:]:
            total = total + (s @ x)
        return total

    def _todense(self):
        total = self._summands[0].todense(cache=False)
        for s in self._summands[1:]:
            total = total + s.todense(cache=False)
        return total

    def _transpose(self):
        return SumLinearOperator(*[s.T for s in self._summands])

    def _diagonal(self):
        total = self._summands[0].diagonal()
        for s in self._summands[1:]:
            total = total + s.diagonal()
        return np.array(total)


class ProductLinearOperator(LinearOperator):
    """``A @ B`` (linops/_arithmetic_fallbacks.py:178-258)."""

    def __init__(self, *factors: LinearOperator):
        for a, b in zip(factors[:-1], factors[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError("Shape mismatch in operator product.")
        self._factors = tuple(factors)
        super().__init__((factors[0].shape[0], factors[-1].shape[1]),
                         np.result_type(*[f.dtype for f in factors]))

    def _matmul(self, x):
        for f in reversed(self._factors):
            x = f @ x
        return x

    def _transpose(self):
        return ProductLinearOperator(*[f.T for f in reversed(self._factors)])
