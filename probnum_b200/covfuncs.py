"""Host-side mirror of ``probnum.randprocs.covfuncs`` for the accelerated kernels.

Same class names, constructor arguments, methods, shapes and error behaviour as the
reference (/root/reference/src/probnum/randprocs/covfuncs/), but every evaluation
runs in libpncov on a B200:

==================================  =====================================================
reference                           here
==================================  =====================================================
``CovarianceFunction.__call__``     shape checks identical (:280-357, :606-674); the
                                    Gram-shaped broadcast ``(N0,1,d) x (1,N1,d)`` goes to
                                    the Gram kernel, every other broadcast pattern to the
                                    elementwise ``pncov_pairs`` kernel
``CovarianceFunction.matrix``       ``pncov_gram``  (:359-413, :553-582)
``CovarianceFunction.linop``        :class:`CovarianceLinearOperator` over a resident
                                    device operator (:415-488, :584-604)
``CovarianceLinearOperator``        ``_matmul`` -> ``pncov_matmul`` (matrix-free),
                                    ``_todense`` -> ``pncov_gram``, ``_transpose`` swaps the
                                    point sets (_covariance_linear_operator.py:22-90)
==================================  =====================================================

There is no NumPy evaluation path in this module: without libpncov.so and a GPU the
calls raise.
"""

from __future__ import annotations

import abc
import collections
import fractions
import functools
import operator
import weakref
from typing import Optional, Tuple

import numpy as np

from . import _lib, hostmem, linops

ShapeType = Tuple[int, ...]

# Provenance of recent symmetric float64 Gram matrices returned by ``matrix(X)``: lets a consumer that
# is handed the HOST array (the Bayesian-quadrature belief update passes it to its Cholesky) rebuild
# the matrix on the device in a millisecond instead of uploading N^2 doubles.  Keyed by id(), guarded
# by a weak reference and a sampled checksum, a handful of entries.
_GRAM_PROVENANCE: "collections.OrderedDict" = collections.OrderedDict()
_PROVENANCE_SAMPLES = 257


def _sample_positions(n: int) -> np.ndarray:
    return (np.arange(_PROVENANCE_SAMPLES, dtype=np.int64) * 2654435761) % (n * n)


def _remember_gram(K: np.ndarray, params, x0: np.ndarray) -> None:
    n = K.shape[0]
    if K.ndim != 2 or K.shape[1] != n or K.dtype != np.float64 or n < 512:
        return
    try:
        ref = weakref.ref(K)
    except TypeError:
        return
    _GRAM_PROVENANCE[id(K)] = (ref, params, np.array(x0, dtype=np.float64, copy=True),
                               K.reshape(-1)[_sample_positions(n)].copy())
    while len(_GRAM_PROVENANCE) > 4:
        _GRAM_PROVENANCE.popitem(last=False)


def gram_provenance(K: np.ndarray):
    """``(kernel_params, points)`` if ``K`` is an unmodified symmetric Gram matrix returned by this
    module (same object, sampled entries unchanged), else ``None``."""
    entry = _GRAM_PROVENANCE.get(id(K))
    if entry is None:
        return None
    ref, params, x0, samples = entry
    if ref() is not K or not np.array_equal(K.reshape(-1)[_sample_positions(K.shape[0])], samples):
        _GRAM_PROVENANCE.pop(id(K), None)
        return None
    return params, x0


def _as_shape(x) -> ShapeType:
    """probnum.utils.as_shape (utils/argutils.py:13): int or iterable of ints -> tuple."""
    if isinstance(x, (int, np.integer)):
        return (int(x),)
    try:
        shape = tuple(int(i) for i in x)
    except TypeError as err:
        raise TypeError(f"The given shape {x} must be an integer or an iterable of integers.") from err
    return shape


def _as_numpy_scalar(x, dtype=np.double):
    """probnum.utils.as_numpy_scalar (utils/argutils.py:54)."""
    if np.ndim(x) != 0:
        raise ValueError("The given input is not a scalar.")
    return np.asarray(x, dtype=dtype)[()]


class DeviceMatrix:
    """A row-major float64 matrix in the memory of one GPU (``CovarianceFunction.matrix_device``)."""

    def __init__(self, buffer: "_lib.DeviceBuffer", shape):
        self.buffer, self.shape, self.dtype = buffer, tuple(shape), np.dtype(np.float64)

    @property
    def ptr(self) -> int:
        return self.buffer.ptr

    @property
    def device(self) -> int:
        return self.buffer.device

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f8", "data": (self.buffer.ptr, False), "version": 3, "strides": None}

    def to_host(self) -> np.ndarray:
        out = hostmem.empty(self.shape, np.float64)
        if out.size:
            self.buffer.to_host(out)
        return out

    def close(self) -> None:
        self.buffer.close()


class CovarianceFunction(abc.ABC):
    """(Cross-)covariance function with scalar output, evaluated on the GPU.

    Mirrors the shape bookkeeping and argument validation of the reference base class
    (_covariance_function.py:138-269, :606-693) for the case used by the isotropic
    kernels: ``input_shape_0 == input_shape_1 == input_shape`` and ``output_shape = ()``.
    """

    def __init__(self, *, input_shape):
        self._input_shape = _as_shape(input_shape)
        if len(self._input_shape) > 1:
            raise ValueError("The accelerated covariance functions take scalar or vector inputs.")

    # -- shapes ------------------------------------------------------------------------------
    @property
    def input_shape(self) -> ShapeType:
        return self._input_shape

    input_shape_0 = input_shape
    input_shape_1 = input_shape

    @property
    def input_ndim(self) -> int:
        return len(self._input_shape)

    input_ndim_0 = input_ndim
    input_ndim_1 = input_ndim

    @functools.cached_property
    def input_size(self) -> int:
        return functools.reduce(operator.mul, self._input_shape, 1)

    output_shape_0: ShapeType = ()
    output_shape_1: ShapeType = ()
    output_ndim_0 = 0
    output_ndim_1 = 0
    output_size_0 = 1
    output_size_1 = 1

    def __repr__(self) -> str:
        return (
            f"<{self.__class__.__name__} with input_shape_0={self.input_shape_0},"
            f" input_shape_1={self.input_shape_1}, output_shape_0={self.output_shape_0}, and"
            f" output_shape_1={self.output_shape_1}>"
        )

    # -- public API ----------------------------------------------------------------------------
    def __call__(self, x0, x1) -> np.ndarray:
        """``k(x0, x1)`` vectorised over broadcast batch shapes (:280-357)."""
        x0 = np.asarray(x0)
        if x1 is not None:
            x1 = np.asarray(x1)
        batch_shape = self._check_shapes(x0.shape, x1.shape if x1 is not None else None)
        out = self._evaluate(x0, x1)
        assert out.shape == batch_shape
        return out

    def matrix(self, x0, x1=None) -> np.ndarray:
        """Gram matrix of shape (N0, N1) over flattened batches (:359-413)."""
        x0 = self._preprocess_linop_input(x0, argnum=0)
        if x1 is not None:
            x1 = self._preprocess_linop_input(x1, argnum=1)
        K = self._evaluate_matrix(x0, x1)
        assert isinstance(K, np.ndarray)
        assert K.shape == (x0.shape[0], x1.shape[0] if x1 is not None else x0.shape[0])
        return K

    def matrix_device(self, x0, x1=None, device: Optional[int] = None) -> "DeviceMatrix":
        """The float64 Gram matrix of ``matrix(x0, x1)`` left in HBM (not part of the reference's API): for
        consumers on the device -- a factorisation, further arithmetic, a torch / CuPy computation through
        ``__cuda_array_interface__`` -- that would otherwise pay two PCIe crossings of N0 x N1 doubles.
        ``.to_host()`` gives the array ``matrix`` returns."""
        x0 = self._preprocess_linop_input(x0, argnum=0)
        if x1 is not None:
            x1 = self._preprocess_linop_input(x1, argnum=1)
        n0 = x0.shape[0]
        n1 = n0 if x1 is None else x1.shape[0]
        buf = _lib.DeviceBuffer(8 * max(n0 * n1, 1), device=device)
        if n0 and n1:
            x0f = np.asarray(x0, dtype=np.float64)
            self._gram_to_device(x0f, None if x1 is None else np.asarray(x1, dtype=np.float64), buf)
        return DeviceMatrix(buf, (n0, n1))

    def linop(self, x0, x1=None) -> linops.LinearOperator:
        """Matrix-free operator for the Gram matrix (:415-488)."""
        x0 = self._preprocess_linop_input(x0, argnum=0)
        if x1 is not None:
            x1 = self._preprocess_linop_input(x1, argnum=1)
        op = self._evaluate_linop(x0, x1)
        assert isinstance(op, linops.LinearOperator)
        assert op.shape == (x0.shape[0], x1.shape[0] if x1 is not None else x0.shape[0])
        return op

    # -- validation (same messages and exception types as the reference) ----------------------------
    def _check_shapes(self, x0_shape: ShapeType, x1_shape: Optional[ShapeType] = None) -> ShapeType:
        err_msg = (
            "The shape of the input array `x{argnum}` must match "
            "`input_shape_{argnum}`, i.e. `{input_shape}`, along its trailing "
            "dimensions, but an array with shape `{shape}` was given."
        )
        nd = self.input_ndim
        if x0_shape[len(x0_shape) - nd:] != self.input_shape:
            raise ValueError(err_msg.format(argnum=0, input_shape=self.input_shape, shape=x0_shape))
        batch = x0_shape[: len(x0_shape) - nd]
        if x1_shape is not None:
            if x1_shape[len(x1_shape) - nd:] != self.input_shape:
                raise ValueError(err_msg.format(argnum=1, input_shape=self.input_shape, shape=x1_shape))
            try:
                batch = np.broadcast_shapes(batch, x1_shape[: len(x1_shape) - nd])
            except ValueError as ve:
                raise ValueError(
                    f"The input arrays `x0` and `x1` with shapes {x0_shape} and "
                    f"{x1_shape} can not be broadcast to a common shape."
                ) from ve
        return tuple(batch)

    def _preprocess_linop_input(self, x, argnum: int) -> np.ndarray:
        x = np.asarray(x)
        nd = self.input_ndim
        if not (x.ndim >= nd and x.shape[(x.ndim - nd):] == self.input_shape):
            raise ValueError(
                f"The shape of `x{argnum}` must must match `input_shape_{argnum}`, "
                f"i.e. `{self.input_shape}`, of the covariance function along its "
                f"trailing dimensions, but an array with shape `{x.shape}` was given."
            )
        return x.reshape((-1,) + self.input_shape, order="C")

    # -- device dispatch ---------------------------------------------------------------------------
    @abc.abstractmethod
    def _kernel_params(self) -> _lib.KernelParams:
        """Description of this covariance function for the C ABI."""

    def _points(self, x: np.ndarray, dtype) -> np.ndarray:
        """(N,) + input_shape -> C-contiguous (N, d) in the compute dtype."""
        return np.ascontiguousarray(x.reshape(x.shape[0], max(self.input_size, 1)), dtype=dtype)

    # families without a float32 device path (Linear, Polynomial, WhiteNoise, ProductMatern and
    # arithmetic nodes) promote float32 points to float64, which is also the dtype the
    # reference's NumPy-2 arithmetic returns for them
    _float64_only = False

    def _compute_dtype(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> np.dtype:
        dt = x0.dtype if x1 is None else np.result_type(x0.dtype, x1.dtype)
        if dt == np.float32 and not self._float64_only:
            return np.dtype(np.float32)
        return np.dtype(np.float64)

    def _result_dtype(self, compute_dtype: np.dtype) -> np.dtype:
        """Dtype the reference returns under NumPy 2 for these inputs."""
        if compute_dtype == np.float32 and isinstance(self, ExpQuad):
            return np.dtype(np.float32)
        return np.dtype(np.float64)

    # -- arithmetic (_covariance_function.py:709-741, _arithmetic.py) ---------------------------------
    __array_ufunc__ = None  # NumPy scalars defer to __rmul__ / __radd__ (:713-717)

    def __add__(self, other):
        return SumCovarianceFunction(self, other)

    def __radd__(self, other):
        return SumCovarianceFunction(other, self)

    def __mul__(self, other):
        return _mul(self, other)

    def __rmul__(self, other):
        return _mul(other, self)

    def _diag_self(self, x0: np.ndarray) -> np.ndarray:
        """k(x, x) over a batch, the ``x1 is None`` branch of ``_evaluate``.  k(x, x) = 1 for the
        three isotropic families: the reference returns exact ones as well
        (_exponentiated_quadratic.py:90-94, _rational_quadratic.py:72-76; Matern:
        c_0 * exp(-0) = 1, _matern.py:175-191)."""
        dt = self._result_dtype(self._compute_dtype(x0, None))
        return np.ones(x0.shape[: x0.ndim - self.input_ndim], dtype=dt)

    def _gram_to_device(self, x0: np.ndarray, x1: Optional[np.ndarray], out: "_lib.DeviceBuffer") -> None:
        """Gram matrix (n0, n1), float64, row-major, into device memory (covariance arithmetic)."""
        n1 = x0.shape[0] if x1 is None else x1.shape[0]
        dt = np.dtype(np.float64)
        with _lib.Operator(self._kernel_params(), self._points(x0, dt),
                           None if x1 is None else self._points(x1, dt), device=out.device) as op:
            op.gram_device(out.ptr, n1)

    def _evaluate(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> np.ndarray:
        nd = self.input_ndim
        if x1 is None:
            return self._diag_self(x0)
        b0 = x0.shape[: x0.ndim - nd]
        b1 = x1.shape[: x1.ndim - nd]
        batch = np.broadcast_shapes(b0, b1)
        dt = self._compute_dtype(x0, x1)
        # Gram-shaped broadcast: x0 batch (N0, 1), x1 batch (1, N1)
        if len(b0) == 2 and len(b1) == 2 and b0[1] == 1 and b1[0] == 1 and b0[0] * b1[1] > 1:
            X0 = x0.reshape((b0[0],) + self.input_shape)
            X1 = x1.reshape((b1[1],) + self.input_shape)
            return self._gram(X0, X1, dt)
        size = int(np.prod(batch, dtype=np.int64))
        if size == 0:
            return np.empty(batch, dtype=self._result_dtype(dt))
        full = tuple(batch) + self.input_shape
        p0 = np.broadcast_to(x0, full).reshape(size, -1)
        p1 = np.broadcast_to(x1, full).reshape(size, -1)
        out = _lib.pairs(self._kernel_params(), np.ascontiguousarray(p0, dtype=dt),
                         np.ascontiguousarray(p1, dtype=dt))
        return out.reshape(batch)

    def _gram(self, x0: np.ndarray, x1: Optional[np.ndarray], dt: np.dtype) -> np.ndarray:
        n0 = x0.shape[0]
        n1 = n0 if x1 is None else x1.shape[0]
        out_dt = self._result_dtype(dt)
        if n0 == 0 or n1 == 0:
            return np.empty((n0, n1), dtype=out_dt)
        pts0 = self._points(x0, dt)
        with _lib.Operator(self._kernel_params(), pts0, None if x1 is None else self._points(x1, dt)) as op:
            K = hostmem.empty((n0, n1), op.gram_dtype)
            op.gram(K)
        if x1 is None:
            _remember_gram(K, self._kernel_params(), pts0)
        return K

    def _evaluate_matrix(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> np.ndarray:
        """Override point #1 of the drop-in (:553-582)."""
        return self._gram(x0, x1, self._compute_dtype(x0, x1))

    def _evaluate_linop(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> linops.LinearOperator:
        """Override point #2 of the drop-in (:584-604)."""
        return CovarianceLinearOperator(self, x0, x1)


class CovarianceLinearOperator(linops.LinearOperator):
    """Pairwise covariances of two point sets as a matrix-free operator on the GPU.

    Mirror of _covariance_linear_operator.py:22-90.  The points are uploaded once, when
    the first product is requested, and stay resident; ``L @ v`` evaluates the kernel in
    registers and never builds the matrix (the reference needs KeOps for that, and
    silently caches the dense matrix without it, :78-81).
    """

    def __init__(self, covfunc: CovarianceFunction, x0: np.ndarray, x1: Optional[np.ndarray],
                 *, _transposed_of: Optional["CovarianceLinearOperator"] = None):
        self._covfunc = covfunc
        self._x0 = x0
        self._x1 = x1
        self._parent = _transposed_of
        self._op: Optional[_lib.Operator] = None
        self._last_product = None  # (operand, result) of the most recent narrow product
        # The reference's tests switch its accelerated path off per operator with `_use_keops = False`
        # (test_linop_matrix.py:38-43).  Same attribute here: False selects the conservative device
        # kernels (direct differences like the reference, fixed summation order), never a CPU path.
        self._use_keops = True
        n0 = x0.shape[0]
        n1 = n0 if x1 is None else x1.shape[0]
        shape = (n0, n1) if _transposed_of is None else (n1, n0)
        dtype = np.promote_types(x0.dtype, x1.dtype) if x1 is not None else x0.dtype
        if dtype not in (np.float32, np.float64):
            dtype = np.dtype(np.float64)
        super().__init__(shape, dtype)

    # the device operator is shared between an operator and its transpose
    def _device_op(self) -> _lib.Operator:
        root = self._parent if self._parent is not None else self
        if root._op is None:
            cf = root._covfunc
            dt = cf._compute_dtype(root._x0, root._x1)
            root._op = _lib.Operator(cf._kernel_params(), cf._points(root._x0, dt),
                                     None if root._x1 is None else cf._points(root._x1, dt))
            root._op_conservative = False
        want = not (self._use_keops and root._use_keops)
        if want != root._op_conservative:
            for opt in (_lib.OPT_IDENTITY, _lib.OPT_SYMMETRIC, _lib.OPT_FACTORISED):
                root._op.set_option(opt, False if want else None)
            root._op_conservative = want
            root._memo_epoch = getattr(root, "_memo_epoch", 0) + 1  # invalidates the memos of root and transpose
        return root._op

    @property
    def _is_transposed(self) -> bool:
        return self._parent is not None

    def _matmul(self, x: np.ndarray) -> np.ndarray:
        # The matrix-based solver applies A to the same vector twice per iteration
        # (linalg/solvers/matrixbased.py:872 and :885).  Every product kernel is deterministic (the
        # symmetric-halving ones accumulate in exact integer planes), so serving the repeat from the
        # last (operand, result) pair gives the bits a second launch would give and halves the device
        # work of problinsolve.  The memo belongs to one kernel selection: the
        # device operator is fetched first (it applies `_use_keops` and bumps the epoch on a change).
        self._device_op()
        root = self._parent if self._parent is not None else self
        epoch = getattr(root, "_memo_epoch", 0)
        narrow = x.ndim == 2 and x.shape[-1] <= 8
        if narrow and self._last_product is not None:
            last_x, last_y, last_epoch = self._last_product
            if (last_epoch == epoch and last_x.shape == x.shape and last_x.dtype == x.dtype
                    and np.array_equal(last_x, x)):
                return last_y.copy()
        y = self._matmul_device(x)
        if narrow:
            self._last_product = (np.array(x, copy=True), y.copy(), epoch)
        return y

    def _matmul_device(self, x: np.ndarray) -> np.ndarray:
        M, N = self.shape
        op = self._device_op()
        lead = x.shape[:-2]
        k = x.shape[-1]
        if M == 0 or k == 0:
            return np.zeros(lead + (M, k), dtype=np.result_type(self.dtype, x.dtype))
        if N == 0:
            return np.zeros(lead + (M, k), dtype=np.result_type(self.dtype, x.dtype))
        if lead:
            # stacked operands (..., N, K): fold the stack into columns
            B = int(np.prod(lead, dtype=np.int64))
            v = np.moveaxis(x.reshape(B, N, k), 0, 1).reshape(N, B * k)
            y = op.matmul(v, transpose=self._is_transposed)
            y = np.moveaxis(y.reshape(M, B, k), 1, 0).reshape(lead + (M, k))
        else:
            y = op.matmul(x, transpose=self._is_transposed)
        return y.astype(np.result_type(self.dtype, x.dtype), copy=False)

    def _todense(self) -> np.ndarray:
        op = self._device_op()
        K = hostmem.empty(self.shape, op.gram_dtype)
        if K.size:
            op.gram(K, transpose=self._is_transposed)
        return K

    def _transpose(self) -> linops.LinearOperator:
        if self._parent is not None:
            return self._parent
        return CovarianceLinearOperator(self._covfunc, self._x0, self._x1, _transposed_of=self)

    def _diagonal(self) -> np.ndarray:
        # k(x0_i, x1_i): O(N d) through the elementwise kernel instead of N matvecs
        n = min(self.shape)
        if self._x1 is None:
            return np.ones(n, dtype=self._covfunc._result_dtype(np.dtype(self.dtype)))
        return np.array(self._covfunc(self._x0[:n], self._x1[:n]))

    def close(self) -> None:
        root = self._parent if self._parent is not None else self
        if root._op is not None:
            root._op.close()
            root._op = None


def _lengthscales_for_device(lengthscales: np.ndarray, input_shape: ShapeType) -> np.ndarray:
    if lengthscales.ndim == 0:
        return lengthscales.reshape(1)
    return np.ascontiguousarray(np.broadcast_to(lengthscales, input_shape), dtype=np.double).reshape(-1)


class ExpQuad(CovarianceFunction):
    r"""Exponentiated quadratic / RBF kernel
    :math:`k(x_0,x_1)=\exp(-\frac12\sum_i((x_{0,i}-x_{1,i})/l_i)^2)`
    (_exponentiated_quadratic.py:20-111)."""

    def __init__(self, input_shape, *, lengthscales=1.0):
        super().__init__(input_shape=input_shape)
        lengthscales = np.asarray(lengthscales if lengthscales is not None else 1.0, dtype=np.double)
        if np.any(lengthscales <= 0):
            raise ValueError(f"Lengthscales l={lengthscales} must be positive.")
        np.broadcast_to(lengthscales, self.input_shape)  # must broadcast to the input dimension
        self._lengthscales = lengthscales

    @property
    def lengthscales(self) -> np.ndarray:
        return self._lengthscales

    @property
    def lengthscale(self) -> np.ndarray:
        if self._lengthscales.shape == ():
            return self._lengthscales
        raise ValueError("There is more than one lengthscale.")

    def _kernel_params(self) -> _lib.KernelParams:
        return _lib.KernelParams(_lib.EXPQUAD, max(self.input_size, 1),
                                 _lengthscales_for_device(self._lengthscales, self.input_shape))


class Matern(CovarianceFunction):
    r"""Matern kernel with half-integer smoothness :math:`\nu=p+\frac12`
    (_matern.py:22-267).  The Bessel-function branch for other :math:`\nu`
    (_matern.py:195, :270-279) is outside the accelerated scope and raises."""

    def __init__(self, input_shape, nu=1.5, *, lengthscales=None):
        super().__init__(input_shape=input_shape)
        nu = _as_numpy_scalar(nu)
        if nu <= 0:
            raise ValueError(f"Hyperparameter nu={nu} must be positive.")
        self._nu = nu
        lengthscales = np.asarray(lengthscales if lengthscales is not None else 1.0, dtype=np.double)
        if np.any(lengthscales <= 0):
            raise ValueError(f"All lengthscales l={lengthscales} must be positive.")
        np.broadcast_to(lengthscales, self.input_shape)
        self._lengthscales = lengthscales

    @property
    def nu(self):
        return self._nu

    @functools.cached_property
    def p(self) -> Optional[int]:
        nu_minus_half = self._nu - 0.5
        if nu_minus_half == int(nu_minus_half):
            return int(nu_minus_half)
        return None

    @property
    def is_half_integer(self) -> bool:
        return self.p is not None

    @property
    def lengthscales(self) -> np.ndarray:
        return self._lengthscales

    @property
    def lengthscale(self) -> np.ndarray:
        if self._lengthscales.shape == ():
            return self._lengthscales
        raise ValueError("There is more than one lengthscale.")

    @staticmethod
    @functools.lru_cache(maxsize=None)
    def half_integer_coefficients(p: int) -> Tuple[fractions.Fraction, ...]:
        """Exact polynomial coefficients, degree i at index i (_matern.py:225-261)."""
        coeffs = [fractions.Fraction(1)]
        for i in reversed(range(p)):
            coeffs.append(coeffs[-1] * fractions.Fraction(2 * (i + 1), (p + i + 1) * (p - i)))
        return tuple(coeffs)

    def _kernel_params(self) -> _lib.KernelParams:
        if not self.is_half_integer:
            raise NotImplementedError(
                f"Matern with nu={self._nu} needs the Bessel-function branch, which the B200 path "
                "does not implement (half-integer nu only)."
            )
        if self.p > _lib.MAX_MATERN_P:
            raise NotImplementedError(f"Matern degree p={self.p} exceeds {_lib.MAX_MATERN_P}.")
        return _lib.KernelParams(_lib.MATERN, max(self.input_size, 1),
                                 _lengthscales_for_device(self._lengthscales, self.input_shape),
                                 nu=float(self._nu))


class RatQuad(CovarianceFunction):
    r"""Rational quadratic kernel
    :math:`k(x_0,x_1)=(1+\lVert x_0-x_1\rVert^2/(2\alpha l^2))^{-\alpha}`
    (_rational_quadratic.py:18-84)."""

    def __init__(self, input_shape, lengthscale=1.0, alpha=1.0):
        self.lengthscale = _as_numpy_scalar(lengthscale)
        self.alpha = _as_numpy_scalar(alpha)
        if not self.alpha > 0:
            raise ValueError(f"Scale mixture alpha={self.alpha} must be positive.")
        super().__init__(input_shape=input_shape)

    def _kernel_params(self) -> _lib.KernelParams:
        if self.lengthscale == 0:
            raise ValueError("Lengthscale l=0 makes the rational quadratic kernel undefined.")
        # only l^2 enters the formula (_rational_quadratic.py:78-84)
        return _lib.KernelParams(_lib.RATQUAD, max(self.input_size, 1), np.reshape(abs(self.lengthscale), 1),
                                 alpha=float(self.alpha))


# =================================================================================================
# the remaining covariance functions and covariance arithmetic
# =================================================================================================
class _DenseOnly(CovarianceFunction):
    """Families served by the dense Gram and elementwise kernels only (pncov_extra.cu)."""

    _float64_only = True

    def _diag_self(self, x0: np.ndarray) -> np.ndarray:
        # k(x_b, x_b) through the elementwise kernel
        batch = x0.shape[: x0.ndim - self.input_ndim]
        size = int(np.prod(batch, dtype=np.int64))
        if size == 0:
            return np.empty(batch, dtype=np.float64)
        p = np.ascontiguousarray(x0.reshape(size, -1), dtype=np.float64)
        return _lib.pairs(self._kernel_params(), p, p).reshape(batch)

    def _evaluate_linop(self, x0, x1):
        # The reference has no matrix-free form for these families either: its operator caches the
        # dense matrix and multiplies with NumPy (_covariance_linear_operator.py:78-81).  There is no
        # device product for them here and no CPU stand-in: fail loudly.  Under patch.enable() the
        # reference's own dense-cached operator is kept and only its matrix comes from the device.
        raise NotImplementedError(
            f"{type(self).__name__} has no matrix-free product on the device; use .matrix() "
            "(the dense Gram is built on the GPU)")


class Linear(_DenseOnly):
    r"""Linear kernel :math:`k(x_0,x_1)=x_0^\top x_1+c` (_linear.py:14-50)."""

    def __init__(self, input_shape, constant=0.0):
        self.constant = _as_numpy_scalar(constant)
        super().__init__(input_shape=input_shape)

    def _kernel_params(self) -> _lib.KernelParams:
        return _lib.KernelParams(_lib.LINEAR, max(self.input_size, 1), None, alpha=float(self.constant))


class Polynomial(_DenseOnly):
    r"""Polynomial kernel :math:`k(x_0,x_1)=(x_0^\top x_1+c)^q` (_polynomial.py:14-56).  The device
    path takes integer exponents 0..64 (repeated multiplication); others raise."""

    def __init__(self, input_shape, constant=0.0, exponent=1.0):
        self.constant = _as_numpy_scalar(constant)
        self.exponent = _as_numpy_scalar(exponent)
        super().__init__(input_shape=input_shape)

    def _kernel_params(self) -> _lib.KernelParams:
        q = float(self.exponent)
        if not (0 <= q <= 64 and q == int(q)):
            raise NotImplementedError(f"Polynomial exponent {q} is outside the device path (integers 0..64).")
        return _lib.KernelParams(_lib.POLYNOMIAL, max(self.input_size, 1), None, nu=q, alpha=float(self.constant))


class WhiteNoise(_DenseOnly):
    r"""White noise :math:`k(x_0,x_1)=\sigma^2\delta(x_0,x_1)` (_white_noise.py:14-55)."""

    def __init__(self, input_shape, sigma_sq=1.0):
        if sigma_sq < 0:
            raise ValueError(f"Noise level sigma_sq={sigma_sq} must be non-negative.")
        self._sigma_sq = _as_numpy_scalar(sigma_sq)
        super().__init__(input_shape=input_shape)

    @property
    def sigma_sq(self):
        return self._sigma_sq

    def _kernel_params(self) -> _lib.KernelParams:
        return _lib.KernelParams(_lib.WHITENOISE, max(self.input_size, 1), None, alpha=float(self._sigma_sq))


class ProductMatern(_DenseOnly):
    r"""Product of one-dimensional Matern kernels
    :math:`k(x_0,x_1)=\prod_i k_i(x_{0,i},x_{1,i})` (_product_matern.py:17-122), half-integer
    smoothness in every dimension."""

    def __init__(self, input_shape, lengthscales, nus):
        input_shape = _as_shape(input_shape)
        if input_shape == () and not (np.isscalar(lengthscales) and np.isscalar(nus)):
            raise ValueError(f"'lengthscales' and 'nus' must be scalar if 'input_shape' is {input_shape}.")
        input_dim = 1 if input_shape == () else input_shape[0]

        def expand(x):
            x = np.asarray(x, dtype=np.double)
            return np.full((input_dim,), x[()]) if x.shape == () else x

        self.lengthscales = expand(lengthscales)
        self.nus = expand(nus)
        if self.lengthscales.shape != (input_dim,) or self.nus.shape != (input_dim,):
            raise ValueError("'lengthscales' and 'nus' must be scalars or have one entry per input dimension.")
        # the reference builds one Matern per dimension, whose constructor validates these
        self.univariate_materns = [Matern((), nu=self.nus[i], lengthscales=self.lengthscales[i])
                                   for i in range(input_dim)]
        super().__init__(input_shape=input_shape)

    def _diag_self(self, x0: np.ndarray) -> np.ndarray:
        return np.ones(x0.shape[: x0.ndim - self.input_ndim], dtype=np.float64)  # product of ones

    def _kernel_params(self) -> _lib.KernelParams:
        for m in self.univariate_materns:
            if not m.is_half_integer or m.p > _lib.MAX_MATERN_P:
                raise NotImplementedError("ProductMatern needs half-integer nu in every dimension on the device path.")
        return _lib.KernelParams(_lib.PRODUCT_MATERN, max(self.input_size, 1),
                                 np.concatenate([self.lengthscales, self.nus]))


# -- arithmetic (_arithmetic.py, _arithmetic_fallbacks.py) ----------------------------------------------
_ARITH_BLOCK_BYTES = 4 << 30  # device-side row blocks of arithmetic results
_FUSED_ARITHMETIC = True      # one-launch evaluation where the expression allows it (tests switch it off to compare)


class _Unfusable(Exception):
    pass


def _compile_program(node):
    """Translate an arithmetic expression into the program of ``pncov_set_program`` -- step for step the
    operations `_combine_into` below applies to whole matrices, so both paths give the same bits -- or
    return None when the expression does not fit: leaves must be isotropic ExpQuad / Matern / RatQuad or
    WhiteNoise with d <= 16, at most 6 leaf evaluations, and the tree must evaluate with two values per entry
    (a sum of [scaled] leaves or [scaled] products of leaves, optionally scaled as a whole)."""
    leaves, steps = [], []

    def leaf_index(k):
        if isinstance(k, _ArithmeticNode) or max(k.input_size, 1) > 16:
            raise _Unfusable
        if isinstance(k, WhiteNoise):
            desc = (_lib.WHITENOISE, 1.0, 0.5, float(k.sigma_sq))
        elif isinstance(k, (ExpQuad, Matern, RatQuad)):
            try:
                kp = k._kernel_params()
            except NotImplementedError:
                raise _Unfusable from None
            if kp.lengthscales.size != 1:
                raise _Unfusable  # per-dimension lengthscales: S differs between the leaves
            desc = (kp.kind, float(kp.lengthscales[0]), kp.nu, kp.alpha)
        else:
            raise _Unfusable
        if desc not in leaves:
            leaves.append(desc)
        return leaves.index(desc)

    def into_val(k):  # `val` = k, without touching `acc`
        if isinstance(k, ScaledCovarianceFunction):
            into_val(k._covfunc)
            steps.append((_lib.F_SCALEVAL, 0, float(k._scalar), 0.0))
        elif isinstance(k, ProductCovarianceFunction):
            if any(isinstance(f, _ArithmeticNode) for f in k._factors):
                raise _Unfusable
            steps.append((_lib.F_LOAD, leaf_index(k._factors[0]), 0.0, 0.0))
            for f in k._factors[1:]:
                steps.append((_lib.F_MULLEAF, leaf_index(f), 0.0, 0.0))
        elif isinstance(k, _ArithmeticNode):
            raise _Unfusable
        else:
            steps.append((_lib.F_LOAD, leaf_index(k), 0.0, 0.0))

    def into_acc(k):  # `acc` = k; mirrors _combine_into of each node type
        if isinstance(k, ScaledCovarianceFunction):
            into_acc(k._covfunc)
            steps.append((_lib.F_SCALEACC, 0, float(k._scalar), 0.0))
        elif isinstance(k, SumCovarianceFunction):
            into_acc(k._summands[0])
            for s_ in k._summands[1:]:
                if isinstance(s_, ScaledCovarianceFunction):
                    into_val(s_._covfunc)
                    steps.append((_lib.F_LIN, 0, 1.0, float(s_._scalar)))
                else:
                    into_val(s_)
                    steps.append((_lib.F_LIN, 0, 1.0, 1.0))
        elif isinstance(k, ProductCovarianceFunction):
            into_acc(k._factors[0])
            for f in k._factors[1:]:
                into_val(f)
                steps.append((_lib.F_MUL, 0, 0.0, 0.0))
        else:
            into_val(k)
            steps.append((_lib.F_SET, 0, 0.0, 0.0))

    try:
        into_acc(node)
    except _Unfusable:
        return None
    evaluations = sum(1 for st in steps if st[0] in (_lib.F_LOAD, _lib.F_MULLEAF))
    if evaluations > _lib.FUSED_MAX_LEAVES or len(steps) > _lib.FUSED_MAX_STEPS:
        return None
    # between two leaf evaluations the kernel's canonical slot takes [SCALEVAL] [SET | LIN | MUL] [SCALEACC], each at
    # most once and in this order (e.g. a scalar times a scalar times a kernel does not fit)
    stage = 0
    for st in steps:
        if st[0] in (_lib.F_LOAD, _lib.F_MULLEAF):
            stage = 0
            continue
        this = 1 if st[0] == _lib.F_SCALEVAL else (3 if st[0] == _lib.F_SCALEACC else 2)
        if this <= stage:
            return None
        stage = this
    # the handle only carries the points: any isotropic description with the right d will do
    params = _lib.KernelParams(_lib.EXPQUAD, max(node.input_size, 1), np.ones(1))
    return leaves, steps, params


class _ArithmeticNode(CovarianceFunction):
    """Shared evaluation of Scaled / Sum / Product covariance functions: every operand's Gram
    block is built on the device and combined there in place (pncov_combine_device); only the
    result crosses PCIe.  The reference combines full host matrices with NumPy
    (_arithmetic_fallbacks.py:55-56, :101-104, :163-166)."""

    _float64_only = True

    def _kernel_params(self):
        raise NotImplementedError("arithmetic nodes are evaluated operand by operand")

    def _children(self):
        raise NotImplementedError

    def _combine_into(self, x0, x1, out, n_elems) -> None:
        raise NotImplementedError

    def _gram_to_device(self, x0, x1, out) -> None:
        n1 = x0.shape[0] if x1 is None else x1.shape[0]
        prog = _compile_program(self) if _FUSED_ARITHMETIC else None
        if prog is not None:  # one launch (see _evaluate_matrix)
            leaves, steps, params = prog
            dt = np.dtype(np.float64)
            with _lib.Operator(params, self._points(x0, dt), None if x1 is None else self._points(x1, dt),
                               device=out.device) as op:
                op.set_program(leaves, steps)
                op.gram_device(out.ptr, n1)
            return
        self._combine_into(x0, x1, out, x0.shape[0] * n1)

    def _result_dtype(self, compute_dtype: np.dtype) -> np.dtype:
        return np.dtype(np.float64)

    def _leaf_result_dtype(self, in_dtype: np.dtype) -> np.dtype:
        """NumPy promotion of the operands' result dtypes (float32 points: ExpQuad stays float32,
        everything else, and every NumPy-scalar factor, is float64)."""
        dts = []
        for c in self._children():
            if isinstance(c, _ArithmeticNode):
                dts.append(c._leaf_result_dtype(in_dtype))
            else:
                dts.append(c._result_dtype(np.dtype(np.float32) if in_dtype == np.float32 and not c._float64_only
                                           else np.dtype(np.float64)))
        if isinstance(self, ScaledCovarianceFunction):
            dts.append(np.dtype(np.float64))
        return np.result_type(*dts)

    def _evaluate_matrix(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> np.ndarray:
        n0 = x0.shape[0]
        n1 = n0 if x1 is None else x1.shape[0]
        in_dt = x0.dtype if x1 is None else np.result_type(x0.dtype, x1.dtype)
        out_dt = self._leaf_result_dtype(np.dtype(in_dt))
        if n0 == 0 or n1 == 0:
            return np.empty((n0, n1), dtype=out_dt)
        K = hostmem.empty((n0, n1), np.float64)
        prog = _compile_program(self) if _FUSED_ARITHMETIC else None
        if prog is not None and np.dtype(in_dt) == np.float64:
            # all leaves share S_ij: the whole expression in ONE launch per row block, written once
            leaves, steps, params = prog
            pts0 = self._points(x0, np.float64)
            with _lib.Operator(params, pts0, None if x1 is None else self._points(x1, np.float64)) as op:
                op.set_program(leaves, steps)
                op.gram(K)
            return K if out_dt == np.float64 else K.astype(out_dt)
        rows = max(1, min(n0, _ARITH_BLOCK_BYTES // (8 * n1)))
        if rows >= n0:  # one block: keeps x1 = None, so the symmetric kernels apply
            with _lib.DeviceBuffer(8 * n0 * n1) as acc:
                self._gram_to_device(x0, x1, acc)
                acc.to_host(K)
        else:
            x1e = x0 if x1 is None else x1
            with _lib.DeviceBuffer(8 * rows * n1) as acc:
                for r0 in range(0, n0, rows):
                    blk = x0[r0:r0 + rows]
                    self._gram_to_device(blk, x1e, acc)
                    acc.to_host(K[r0:r0 + blk.shape[0]])
        return K if out_dt == np.float64 else K.astype(out_dt)

    def _evaluate(self, x0: np.ndarray, x1: Optional[np.ndarray]) -> np.ndarray:
        nd = self.input_ndim
        if x1 is not None:
            b0 = x0.shape[: x0.ndim - nd]
            b1 = x1.shape[: x1.ndim - nd]
            if len(b0) == 2 and len(b1) == 2 and b0[1] == 1 and b1[0] == 1 and b0[0] * b1[1] > 1:
                return self._evaluate_matrix(x0.reshape((b0[0],) + self.input_shape),
                                             x1.reshape((b1[1],) + self.input_shape))
        # O(N d) broadcasts: operand by operand through the elementwise kernels
        return self._evaluate_elementwise(x0, x1)

    def _evaluate_linop(self, x0, x1):
        # Scaled and Sum nodes compose the operators of their operands (linop() below); a Product has
        # no operator form short of the dense matrix, which this module does not multiply on the CPU
        raise NotImplementedError(
            f"{type(self).__name__} has no matrix-free product on the device; use .matrix()")


def _check_same_shapes(operands, what: str) -> None:
    if not all(isinstance(o, CovarianceFunction) for o in operands):
        raise TypeError(f"All {what} must be covariance functions.")
    if not all(o.input_shape == operands[0].input_shape for o in operands):
        raise ValueError(f"All {what} must have the same in- and output shape.")


class ScaledCovarianceFunction(_ArithmeticNode):
    r""":math:`k(x_0,x_1)=o\,k'(x_0,x_1)` (_arithmetic_fallbacks.py:21-66)."""

    def __init__(self, covfunc: CovarianceFunction, scalar):
        if not isinstance(covfunc, CovarianceFunction):
            raise TypeError("`covfunc` must be a `CovarianceFunction`")
        if np.ndim(scalar) != 0:
            raise TypeError("`scalar` must be a scalar.")
        self._covfunc = covfunc
        self._scalar = _as_numpy_scalar(scalar)
        super().__init__(input_shape=covfunc.input_shape)

    def _children(self):
        return (self._covfunc,)

    def _combine_into(self, x0, x1, out, n_elems) -> None:
        self._covfunc._gram_to_device(x0, x1, out)
        _lib.combine(out, None, n_elems, _lib.COMBINE_LINEAR, alpha=float(self._scalar), beta=0.0)

    def _evaluate_elementwise(self, x0, x1):
        return self._scalar * self._covfunc(x0, x1)

    def linop(self, x0, x1=None):
        return self._scalar * self._covfunc.linop(x0, x1)

    def __repr__(self) -> str:
        return f"{self._scalar} * {self._covfunc}"


class SumCovarianceFunction(_ArithmeticNode):
    r""":math:`k(x_0,x_1)=\sum_i k_i(x_0,x_1)` (_arithmetic_fallbacks.py:69-135)."""

    def __init__(self, *summands: CovarianceFunction):
        _check_same_shapes(summands, "summands")
        expanded = []
        for s_ in summands:
            expanded.extend(s_._summands if isinstance(s_, SumCovarianceFunction) else (s_,))
        self._summands = tuple(expanded)
        super().__init__(input_shape=summands[0].input_shape)

    def _children(self):
        return self._summands

    def _combine_into(self, x0, x1, out, n_elems) -> None:
        self._summands[0]._gram_to_device(x0, x1, out)
        if len(self._summands) == 1:
            return
        with _lib.DeviceBuffer(8 * n_elems, device=out.device) as tmp:
            for s_ in self._summands[1:]:
                if isinstance(s_, ScaledCovarianceFunction):  # out += o * k': one pass instead of two
                    s_._covfunc._gram_to_device(x0, x1, tmp)
                    _lib.combine(out, tmp, n_elems, _lib.COMBINE_LINEAR, alpha=1.0, beta=float(s_._scalar))
                else:
                    s_._gram_to_device(x0, x1, tmp)
                    _lib.combine(out, tmp, n_elems, _lib.COMBINE_LINEAR, alpha=1.0, beta=1.0)

    def _evaluate_elementwise(self, x0, x1):
        return functools.reduce(operator.add, (s_(x0, x1) for s_ in self._summands))

    def linop(self, x0, x1=None):
        return functools.reduce(operator.add, (s_.linop(x0, x1) for s_ in self._summands))

    def __repr__(self):
        return "SumCovarianceFunction [\n\t" + ",\n\t".join(repr(s_) for s_ in self._summands) + "\n]"


class ProductCovarianceFunction(_ArithmeticNode):
    r""":math:`k(x_0,x_1)=\prod_i k_i(x_0,x_1)` (_arithmetic_fallbacks.py:138-199)."""

    def __init__(self, *factors: CovarianceFunction):
        _check_same_shapes(factors, "factors")
        expanded = []
        for f in factors:
            expanded.extend(f._factors if isinstance(f, ProductCovarianceFunction) else (f,))
        self._factors = tuple(expanded)
        super().__init__(input_shape=factors[0].input_shape)

    def _children(self):
        return self._factors

    def _combine_into(self, x0, x1, out, n_elems) -> None:
        self._factors[0]._gram_to_device(x0, x1, out)
        if len(self._factors) == 1:
            return
        with _lib.DeviceBuffer(8 * n_elems, device=out.device) as tmp:
            for f in self._factors[1:]:
                f._gram_to_device(x0, x1, tmp)
                _lib.combine(out, tmp, n_elems, _lib.COMBINE_MULTIPLY)

    def _evaluate_elementwise(self, x0, x1):
        return functools.reduce(operator.mul, (f(x0, x1) for f in self._factors))

    def __repr__(self):
        return "ProductCovarianceFunction [\n\t" + ",\n\t".join(repr(f) for f in self._factors) + "\n]"


def _mul(op1, op2):
    """_arithmetic_fallbacks.py:202-222 (``_mul_fallback``)."""
    if isinstance(op1, CovarianceFunction):
        if isinstance(op2, CovarianceFunction):
            return ProductCovarianceFunction(op1, op2)
        if np.ndim(op2) == 0:
            return ScaledCovarianceFunction(op1, scalar=op2)
    elif isinstance(op2, CovarianceFunction) and np.ndim(op1) == 0:
        return ScaledCovarianceFunction(op2, scalar=op1)
    return NotImplemented
