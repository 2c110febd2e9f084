"""ctypes binding of libpncov.so (the C ABI in ``include/pncov.h``).

There is exactly one implementation behind this module: the CUDA library.  If the
shared object is missing, or no B200 is visible when a compute call is made, the
call raises -- nothing here falls back to NumPy.
"""

from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpncov.so")

EXPQUAD, MATERN, RATQUAD = 0, 1, 2
LINEAR, POLYNOMIAL, WHITENOISE, PRODUCT_MATERN = 3, 4, 5, 6
COMBINE_LINEAR, COMBINE_MULTIPLY = 0, 1
F_LOAD, F_MULLEAF, F_SCALEVAL, F_SET, F_LIN, F_MUL, F_SCALEACC = range(7)  # pncov_set_program steps
FUSED_MAX_LEAVES, FUSED_MAX_STEPS = 6, 12
CHAIN_MAX_CAPACITY = 12800  # pncov_chain_create: 2 * capacity coefficients fit the shared memory of one CTA
OPT_IDENTITY, OPT_SYMMETRIC, OPT_FACTORISED = 0, 1, 2
F64, F32 = 0, 1
MAX_MATERN_P = 8

_c_i64 = ctypes.c_int64
_c_vp = ctypes.c_void_p
_c_dp = ctypes.POINTER(ctypes.c_double)

# every exported symbol of include/pncov.h with its signature
SIGNATURES = {
    "pncov_version": (ctypes.c_int, []),
    "pncov_last_error": (ctypes.c_char_p, []),
    "pncov_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "pncov_host_alloc": (ctypes.c_int, [ctypes.POINTER(_c_vp), _c_i64]),
    "pncov_host_free": (ctypes.c_int, [_c_vp]),
    "pncov_create": (
        ctypes.c_int,
        [ctypes.POINTER(_c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_i64, _c_dp, _c_i64,
         ctypes.c_double, ctypes.c_double, _c_vp, _c_i64, _c_vp, _c_i64, ctypes.c_int],
    ),
    "pncov_destroy": (ctypes.c_int, [_c_vp]),
    "pncov_shape": (ctypes.c_int, [_c_vp, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "pncov_gram_dtype": (ctypes.c_int, [_c_vp, ctypes.POINTER(ctypes.c_int)]),
    "pncov_gram": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, ctypes.c_int]),
    "pncov_gram_device": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, ctypes.c_int, _c_vp]),
    "pncov_matmul": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64, ctypes.c_int]),
    "pncov_matmul_device": (
        ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64, ctypes.c_int, _c_vp]),
    "pncov_matmul_sym_part_device": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, ctypes.c_int, ctypes.c_int, _c_vp]),
    "pncov_matmul_sym_part_bins_device": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, ctypes.c_int, ctypes.c_int, _c_vp]),
    "pncov_matmul_sym_finish_device": (ctypes.c_int, [_c_vp, _c_vp, _c_vp, _c_vp]),
    "pncov_matmulk_sym_part_bins_device": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, ctypes.c_int, ctypes.c_int, _c_vp]),
    "pncov_matmulk_sym_finish_device": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp]),
    "pncov_set_option": (ctypes.c_int, [_c_vp, ctypes.c_int, ctypes.c_int]),
    "pncov_set_program": (
        ctypes.c_int, [_c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), _c_dp, _c_dp, _c_dp, ctypes.c_int,
                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), _c_dp, _c_dp]),
    "pncov_pairs": (
        ctypes.c_int,
        [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_i64, _c_dp, _c_i64, ctypes.c_double,
         ctypes.c_double, _c_vp, _c_vp, _c_i64, _c_vp],
    ),
    "pncov_device_alloc": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_c_vp), _c_i64]),
    "pncov_device_free": (ctypes.c_int, [ctypes.c_int, _c_vp]),
    "pncov_copy_to_host": (ctypes.c_int, [ctypes.c_int, _c_vp, _c_vp, _c_i64, _c_vp]),
    "pncov_combine_device": (
        ctypes.c_int, [ctypes.c_int, _c_vp, _c_vp, _c_i64, ctypes.c_int, ctypes.c_double, ctypes.c_double, _c_vp]),
    "pncov_chol_create": (ctypes.c_int, [ctypes.POINTER(_c_vp), ctypes.c_int, _c_i64]),
    "pncov_chol_destroy": (ctypes.c_int, [_c_vp]),
    "pncov_chol_factor": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "pncov_chol_factor_gram": (ctypes.c_int, [_c_vp, _c_vp, ctypes.c_double, ctypes.POINTER(ctypes.c_int)]),
    "pncov_chol_solve": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64]),
    "pncov_chain_create": (ctypes.c_int, [ctypes.POINTER(_c_vp), ctypes.c_int, _c_i64, _c_i64]),
    "pncov_chain_destroy": (ctypes.c_int, [_c_vp]),
    "pncov_chain_set": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp]),
    "pncov_chain_apply": (ctypes.c_int, [_c_vp, _c_i64, ctypes.c_int, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64]),
    "pncov_chol_solve_gram": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_float)]),
    "pncov_chol_get_factor": (ctypes.c_int, [_c_vp, _c_vp, _c_i64]),
    "pncov_chol_last_factor_ms": (ctypes.c_int, [_c_vp, ctypes.POINTER(ctypes.c_float)]),
    "pncov_sharded_create": (
        ctypes.c_int,
        [ctypes.POINTER(_c_vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_i64, _c_dp, _c_i64,
         ctypes.c_double, ctypes.c_double, _c_vp, _c_i64, _c_vp, _c_i64]),
    "pncov_sharded_gram": (ctypes.c_int, [_c_vp, _c_vp, _c_i64]),
    "pncov_sharded_matmul": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_vp, _c_i64]),
    "pncov_sharded_destroy": (ctypes.c_int, [_c_vp]),
    "pncov_d2h_bytes": (ctypes.c_int64, []),
    "pncov_launch_count": (_c_i64, []),
    "pncov_matmul_path": (ctypes.c_int, [_c_vp, ctypes.POINTER(ctypes.c_int)]),
    "pncov_last_kernel_ms": (ctypes.c_int, [_c_vp, ctypes.POINTER(ctypes.c_float)]),
}

_lib = None
_lock = threading.Lock()


class PncovError(RuntimeError):
    """Device-side failure reported by libpncov (CUDA error, no device, ...)."""


def load() -> ctypes.CDLL:
    """Load libpncov.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PncovError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (nvcc, sm_100a).  probnum_b200 has no CPU fallback."
                )
            lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc == 0:
        return
    msg = load().pncov_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise NotImplementedError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise PncovError(f"libpncov error {rc}: {msg}")


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = load().pncov_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


def default_device() -> int:
    """One process per GPU: LOCAL_RANK selects the device under torchrun."""
    return int(os.environ.get("PNCOV_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def d2h_bytes() -> int:
    """Bytes ``pncov_gram`` has copied device-to-host in this process."""
    return int(load().pncov_d2h_bytes())


def launch_count() -> int:
    return int(load().pncov_launch_count())


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _dtype_code(dt: np.dtype) -> int:
    if dt == np.float64:
        return F64
    if dt == np.float32:
        return F32
    raise ValueError(f"unsupported dtype {dt}")


class KernelParams:
    """Plain description of one covariance function for the C ABI."""

    __slots__ = ("kind", "d", "lengthscales", "nu", "alpha")

    def __init__(self, kind: int, d: int, lengthscales, nu: float = 0.5, alpha: float = 1.0):
        self.kind = int(kind)
        self.d = int(d)
        self.lengthscales = np.ascontiguousarray(
            np.atleast_1d(lengthscales if lengthscales is not None else np.empty(0)), dtype=np.float64)
        self.nu = float(nu)
        self.alpha = float(alpha)

    def _ls(self):
        if self.lengthscales.size == 0:
            return None, 0
        return self.lengthscales.ctypes.data_as(_c_dp), self.lengthscales.size

    @property
    def matrix_free(self) -> bool:
        """Whether pncov_matmul serves this family (the isotropic kernels only)."""
        return self.kind in (EXPQUAD, MATERN, RATQUAD)


class Operator:
    """A covariance function bound to two point sets resident on one B200.

    ``x0`` / ``x1`` are host arrays of shape (n, d), or integer device addresses when
    ``on_device`` is set (then ``n0`` / ``n1`` / ``dtype`` must be given).
    """

    def __init__(self, params: KernelParams, x0, x1=None, *, device: Optional[int] = None,
                 on_device: bool = False, n0: Optional[int] = None, n1: Optional[int] = None,
                 dtype=None):
        lib = load()
        self._lib = lib
        self._h = _c_vp()
        self.params = params
        self.device = default_device() if device is None else int(device)
        if on_device:
            self.dtype = np.dtype(dtype)
            p0, p1 = _c_vp(int(x0)), (None if x1 is None else _c_vp(int(x1)))
            n0, n1 = int(n0), (0 if x1 is None else int(n1))
        else:
            x0 = np.ascontiguousarray(x0)
            x1 = None if x1 is None else np.ascontiguousarray(x1, dtype=x0.dtype)
            if x0.ndim != 2 or x0.shape[1] != params.d or (x1 is not None and (x1.ndim != 2 or x1.shape[1] != params.d)):
                raise ValueError("point sets must have shape (n, d)")
            self.dtype = x0.dtype
            self._keep = (x0, x1)
            p0, p1 = _ptr(x0), _ptr(x1)
            n0, n1 = x0.shape[0], (0 if x1 is None else x1.shape[0])
        ls, nls = params._ls()
        rc = lib.pncov_create(ctypes.byref(self._h), self.device, params.kind, _dtype_code(self.dtype),
                              params.d, ls, nls, params.nu, params.alpha, p0, n0, p1, n1,
                              1 if on_device else 0)
        _check(rc)
        self._keep = None
        a, b = _c_i64(), _c_i64()
        _check(lib.pncov_shape(self._h, ctypes.byref(a), ctypes.byref(b)))
        self.shape = (a.value, b.value)
        g = ctypes.c_int()
        _check(lib.pncov_gram_dtype(self._h, ctypes.byref(g)))
        self.gram_dtype = np.dtype(np.float32 if g.value == F32 else np.float64)

    # -- life cycle -----------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pncov_destroy(self._h)
            self._h = _c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- Gram -------------------------------------------------------------------------------
    def gram(self, out: np.ndarray, transpose: bool = False) -> np.ndarray:
        """Fill the host array ``out`` (C-contiguous rows, dtype ``gram_dtype``)."""
        rows, cols = self.shape[::-1] if transpose else self.shape
        if out.shape != (rows, cols) or out.dtype != self.gram_dtype:
            raise ValueError("output array has the wrong shape or dtype")
        if out.size and out.strides[1] != out.itemsize:
            raise ValueError("output rows must be contiguous")
        ld = out.strides[0] // out.itemsize if rows > 1 else max(cols, 1)
        _check(self._lib.pncov_gram(self._h, _ptr(out), ld, 1 if transpose else 0))
        return out

    def gram_device(self, out_ptr: int, ld: int, transpose: bool = False, stream: int = 0) -> None:
        _check(self._lib.pncov_gram_device(self._h, _c_vp(out_ptr), ld, 1 if transpose else 0,
                                           _c_vp(stream)))

    # -- matrix-free product ------------------------------------------------------------------
    def matmul(self, v: np.ndarray, transpose: bool = False) -> np.ndarray:
        """``K @ v`` for a host array ``v`` of shape (n1, k); returns (n0, k)."""
        rows, cols = self.shape[::-1] if transpose else self.shape
        v = np.ascontiguousarray(v, dtype=self.dtype)
        if v.ndim != 2 or v.shape[0] != cols:
            raise ValueError(f"operand must have shape ({cols}, k), got {v.shape}")
        k = v.shape[1]
        y = np.empty((rows, k), dtype=self.dtype)
        _check(self._lib.pncov_matmul(self._h, _ptr(v), k, max(k, 1), _ptr(y), max(k, 1),
                                      1 if transpose else 0))
        if cols == 0:
            y[...] = 0
        return y

    def matmul_device(self, v_ptr: int, k: int, ldv: int, y_ptr: int, ldy: int,
                      transpose: bool = False, stream: int = 0) -> None:
        _check(self._lib.pncov_matmul_device(self._h, _c_vp(v_ptr), k, ldv, _c_vp(y_ptr), ldy,
                                             1 if transpose else 0, _c_vp(stream)))

    def set_option(self, option: int, value: Optional[bool]) -> None:
        """OPT_IDENTITY / OPT_SYMMETRIC / OPT_FACTORISED: True, False or None (automatic)."""
        _check(self._lib.pncov_set_option(self._h, int(option), -1 if value is None else int(bool(value))))

    def set_program(self, leaves, steps) -> None:
        """Covariance arithmetic in one launch (pncov_set_program).  ``leaves``: sequence of
        ``(kind, lengthscale, nu, alpha)``; ``steps``: sequence of ``(op, leaf_index, a, b)``."""
        nl, ns = len(leaves), len(steps)
        kinds = (ctypes.c_int * max(nl, 1))(*[int(l[0]) for l in leaves])
        ls = (ctypes.c_double * max(nl, 1))(*[float(l[1]) for l in leaves])
        nus = (ctypes.c_double * max(nl, 1))(*[float(l[2]) for l in leaves])
        al = (ctypes.c_double * max(nl, 1))(*[float(l[3]) for l in leaves])
        ops = (ctypes.c_int * max(ns, 1))(*[int(t[0]) for t in steps])
        li = (ctypes.c_int * max(ns, 1))(*[int(t[1]) for t in steps])
        aa = (ctypes.c_double * max(ns, 1))(*[float(t[2]) for t in steps])
        bb = (ctypes.c_double * max(ns, 1))(*[float(t[3]) for t in steps])
        _check(self._lib.pncov_set_program(self._h, nl, kinds, ls, nus, al, ns, ops, li, aa, bb))

    def conservative(self) -> None:
        """Direct differences and a fixed summation order for every product of this operator."""
        for opt in (OPT_IDENTITY, OPT_SYMMETRIC, OPT_FACTORISED):
            self.set_option(opt, False)

    def matmul_path(self) -> str:
        """'undecided', 'norm-identity-dmma' or 'direct-differences' (see pncov_matmul_path)."""
        p = ctypes.c_int()
        _check(self._lib.pncov_matmul_path(self._h, ctypes.byref(p)))
        return ("undecided", "norm-identity-dmma", "direct-differences", "norm-identity-dmma-symmetric",
                "direct-differences-symmetric")[p.value]

    def matmul_sym_part_device(self, v_ptr: int, y_ptr: int, part: int, nparts: int, stream: int = 0) -> None:
        """Contribution of row blocks ``part, part + nparts, ...`` to ``K @ v`` (full-length ``y``);
        the sum over all parts is the product.  Raises NotImplementedError if the symmetric kernel
        does not apply (then use row sharding)."""
        _check(self._lib.pncov_matmul_sym_part_device(self._h, _c_vp(v_ptr), _c_vp(y_ptr), int(part), int(nparts),
                                                      _c_vp(stream)))

    def matmul_sym_part_bins_device(self, v_ptr: int, bins_ptr: int, part: int, nparts: int, stream: int = 0) -> None:
        """Integer planes (3 n0 words of 8 bytes) of the contribution of row blocks part, part + nparts, ..."""
        _check(self._lib.pncov_matmul_sym_part_bins_device(self._h, _c_vp(v_ptr), _c_vp(bins_ptr), int(part), int(nparts),
                                                           _c_vp(stream)))

    def matmulk_sym_part_bins_device(self, v_ptr: int, k: int, ldv: int, bins_ptr: int, part: int, nparts: int,
                                     stream: int = 0) -> None:
        """1 .. 8 right-hand sides: integer planes (3 n0 words for k = 1, 24 n0 for k > 1) of this part's row blocks."""
        _check(self._lib.pncov_matmulk_sym_part_bins_device(self._h, _c_vp(v_ptr), int(k), int(ldv), _c_vp(bins_ptr), int(part),
                                                            int(nparts), _c_vp(stream)))

    def matmulk_sym_finish_device(self, bins_ptr: int, k: int, y_ptr: int, ldy: int, stream: int = 0) -> None:
        _check(self._lib.pncov_matmulk_sym_finish_device(self._h, _c_vp(bins_ptr), int(k), _c_vp(y_ptr), int(ldy), _c_vp(stream)))

    def matmul_sym_finish_device(self, bins_ptr: int, y_ptr: int, stream: int = 0) -> None:
        """Summed integer planes -> y (n0 doubles)."""
        _check(self._lib.pncov_matmul_sym_finish_device(self._h, _c_vp(bins_ptr), _c_vp(y_ptr), _c_vp(stream)))

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        _check(self._lib.pncov_last_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)


class ShardedOperator:
    """One operator over several GPUs of this process (``pncov_sharded_*``): host arrays in and out, the rows
    of ``x0`` dealt contiguously to ``devices`` (an id may repeat).  No torch, no process launcher."""

    def __init__(self, params: KernelParams, x0: np.ndarray, x1: Optional[np.ndarray] = None, *, devices=None):
        self._lib = load()
        devices = list(range(device_count())) if devices is None else [int(i) for i in devices]
        self.dtype = np.dtype(x0.dtype)
        self.x0 = np.ascontiguousarray(x0)
        self.x1 = None if x1 is None else np.ascontiguousarray(x1, dtype=self.dtype)
        if self.x0.ndim != 2 or self.x0.shape[1] != params.d:
            raise ValueError("x0 must have shape (n0, d)")
        self.shape = (self.x0.shape[0], self.x0.shape[0] if self.x1 is None else self.x1.shape[0])
        self.gram_dtype = np.dtype(np.float32 if (self.dtype == np.float32 and params.kind == EXPQUAD) else np.float64)
        ids = (ctypes.c_int * len(devices))(*devices)
        ls, nls = params._ls()
        self._h = _c_vp()
        _check(self._lib.pncov_sharded_create(ctypes.byref(self._h), ids, len(devices), params.kind, _dtype_code(self.dtype),
                                              params.d, ls, nls, params.nu, params.alpha, _ptr(self.x0), self.shape[0],
                                              _ptr(self.x1), 0 if self.x1 is None else self.shape[1]))

    def gram(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=self.gram_dtype)
        if out.shape != self.shape or out.dtype != self.gram_dtype or (out.size and out.strides[1] != out.itemsize):
            raise ValueError("output array has the wrong shape, dtype or layout")
        ld = out.strides[0] // out.itemsize if self.shape[0] > 1 else max(self.shape[1], 1)
        _check(self._lib.pncov_sharded_gram(self._h, _ptr(out), ld))
        return out

    def matmul(self, v: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(v, dtype=self.dtype)
        if v.ndim != 2 or v.shape[0] != self.shape[1]:
            raise ValueError("v must have shape (n1, k)")
        k = v.shape[1]
        y = np.zeros((self.shape[0], k), dtype=self.dtype) if self.shape[1] == 0 else np.empty((self.shape[0], k), dtype=self.dtype)
        _check(self._lib.pncov_sharded_matmul(self._h, _ptr(v), k, max(k, 1), _ptr(y), max(k, 1)))
        return y

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.pncov_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def pairs(params: KernelParams, x0: np.ndarray, x1: np.ndarray, device: Optional[int] = None) -> np.ndarray:
    """out[b] = k(x0[b], x1[b]) for two host arrays of shape (B, d)."""
    lib = load()
    x0 = np.ascontiguousarray(x0)
    x1 = np.ascontiguousarray(x1, dtype=x0.dtype)
    if x0.shape != x1.shape or x0.ndim != 2 or x0.shape[1] != params.d:
        raise ValueError("pairs() needs two (B, d) arrays")
    out_dtype = np.float32 if (x0.dtype == np.float32 and params.kind == EXPQUAD) else np.float64
    out = np.empty(x0.shape[0], dtype=out_dtype)
    ls, nls = params._ls()
    _check(lib.pncov_pairs(default_device() if device is None else device, params.kind,
                           _dtype_code(x0.dtype), params.d, ls, nls, params.nu, params.alpha,
                           _ptr(x0), _ptr(x1), x0.shape[0], _ptr(out)))
    return out


# -- device buffers and covariance arithmetic (include/pncov.h, "covariance arithmetic") ------------
class DeviceBuffer:
    """``nbytes`` of device memory on one GPU, freed on close / garbage collection."""

    def __init__(self, nbytes: int, device: Optional[int] = None):
        self.device = default_device() if device is None else int(device)
        self.nbytes = int(nbytes)
        p = _c_vp()
        _check(load().pncov_device_alloc(self.device, ctypes.byref(p), self.nbytes))
        self.ptr = int(p.value or 0)

    def close(self) -> None:
        if getattr(self, "ptr", 0):
            load().pncov_device_free(self.device, _c_vp(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def to_host(self, out: np.ndarray, stream: int = 0) -> np.ndarray:
        if not out.flags.c_contiguous or out.nbytes > self.nbytes:
            raise ValueError("to_host needs a C-contiguous array no larger than the buffer")
        _check(load().pncov_copy_to_host(self.device, _ptr(out), _c_vp(self.ptr), out.nbytes, _c_vp(stream)))
        return out


def combine(dst: DeviceBuffer, src: Optional[DeviceBuffer], n: int, op: int, alpha: float = 1.0,
            beta: float = 1.0, stream: int = 0) -> None:
    """``dst = alpha * dst + beta * src`` (COMBINE_LINEAR) or ``dst *= src`` (COMBINE_MULTIPLY) over
    ``n`` float64 elements on the device."""
    _check(load().pncov_combine_device(dst.device, _c_vp(dst.ptr), None if src is None else _c_vp(src.ptr),
                                       int(n), int(op), float(alpha), float(beta), _c_vp(stream)))


# -- Cholesky (include/pncov.h, "Cholesky factor and solves on the device") ---------------------------
class Cholesky:
    """Device-resident Cholesky factor of ``a + jitter * I`` (float64, lower triangle of ``a`` read).

    ``a`` is a host array or, with ``on_device``, an integer device address of an (n, n) row-major
    matrix.  ``upper=True`` (host arrays): the UPPER triangle is the one that counts, as in
    ``scipy.linalg.cho_factor(a)`` with its default ``lower=False``; only that half is uploaded.
    Raises ``numpy.linalg.LinAlgError`` like ``scipy.linalg.cho_factor`` if the matrix is
    not positive definite."""

    def __init__(self, a, jitter: float = 0.0, *, device: Optional[int] = None, on_device: bool = False,
                 n: Optional[int] = None, ld: Optional[int] = None, upper: bool = False):
        lib = load()
        self._lib = lib
        self._h = _c_vp()
        self.device = default_device() if device is None else int(device)
        if isinstance(a, Operator):  # Gram matrix of a square operator, evaluated into the factor buffer
            if a.shape[0] != a.shape[1]:
                raise ValueError("expected a square operator")
            self.n = a.shape[0]
            self.device = a.device
            _check(lib.pncov_chol_create(ctypes.byref(self._h), self.device, self.n))
            info = ctypes.c_int(0)
            try:
                _check(lib.pncov_chol_factor_gram(self._h, a._h, float(jitter), ctypes.byref(info)))
            except Exception:
                self.close()
                raise
            if info.value != 0:
                self.close()
                raise np.linalg.LinAlgError(f"{info.value}-th leading minor of the array is not positive definite")
            return
        if on_device:
            self.n = int(n)
            ptr, lda = _c_vp(int(a)), int(ld if ld is not None else n)
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != a.shape[1]:
                raise ValueError("expected a square matrix")
            self.n = a.shape[0]
            ptr, lda = _ptr(a), a.shape[1]
        _check(lib.pncov_chol_create(ctypes.byref(self._h), self.device, self.n))
        info = ctypes.c_int(0)
        try:
            _check(lib.pncov_chol_factor(self._h, ptr, lda, float(jitter), 1 if on_device else (2 if upper else 0),
                                         ctypes.byref(info)))
        except Exception:
            self.close()
            raise
        if info.value != 0:
            self.close()
            raise np.linalg.LinAlgError(f"{info.value}-th leading minor of the array is not positive definite")

    def solve(self, b: np.ndarray) -> np.ndarray:
        """``(L L^T)^-1 b`` for a host array of shape (n,) or (n, k)."""
        b = np.asarray(b, dtype=np.float64)
        b2 = np.ascontiguousarray(b.reshape(self.n, -1))
        x = np.empty_like(b2)
        k = b2.shape[1]
        if k:
            _check(self._lib.pncov_chol_solve(self._h, _ptr(b2), k, k, _ptr(x), k))
        return x.reshape(b.shape)

    def solve_gram(self, op: "Operator", vecs=(), *, forward_only: bool = False):
        """Prediction path: with ``G = K(nodes, x)`` the Gram matrix of ``op``
        (``op.shape == (n, m)``, first point set = the factorised nodes) and ``W = (L L^T)^-1 G``
        returns ``(quad, dots)`` with ``quad[j] = sum_i W[i, j] G[i, j]`` and
        ``dots[t, j] = vecs[t] @ W[:, j]`` for up to four host vectors of length n.  G and W stay
        on the device.  ``self.last_solve_gram_ms`` holds the device time of the pass.
        ``forward_only``: the same quantities from ``Y = L^-1 G`` alone (``quad = |Y[:, j]|^2``,
        ``dots = (L^-1 vecs[t]) . Y[:, j]``): half the flops and no cancelling sums; the default
        follows the reference's order of operations (both substitutions, then the sums)."""
        if op.shape[0] != self.n:
            raise ValueError(f"the operator's first point set has {op.shape[0]} points, the factor has order {self.n}")
        m = op.shape[1]
        v = np.ascontiguousarray(np.asarray(vecs, dtype=np.float64).reshape(-1, self.n)) if len(vecs) else np.empty((0, self.n))
        if v.shape[0] > 4:
            raise ValueError("at most 4 vectors")
        quad = np.empty(m, dtype=np.float64)
        dots = np.empty((v.shape[0], m), dtype=np.float64)
        ms = ctypes.c_float(0.0)
        _check(self._lib.pncov_chol_solve_gram(self._h, op._h, v.shape[0], _ptr(v) if v.shape[0] else None, self.n,
                                               _ptr(dots) if v.shape[0] else None, max(m, 1), _ptr(quad),
                                               1 if forward_only else 0, ctypes.byref(ms)))
        self.last_solve_gram_ms = float(ms.value)
        return quad, dots

    def factor(self) -> np.ndarray:
        """Lower-triangular factor as a host array (upper triangle zeroed)."""
        L = np.empty((self.n, self.n), dtype=np.float64)
        _check(self._lib.pncov_chol_get_factor(self._h, _ptr(L), self.n))
        return np.tril(L)

    def last_factor_ms(self) -> float:
        ms = ctypes.c_float()
        _check(self._lib.pncov_chol_last_factor_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pncov_chol_destroy(self._h)
            self._h = _c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


class UpdateChain:
    """Device-resident chain of vector pairs (SURVEY.md 8f N2 ii): the rank-2 / rank-1 update sums of
    the matrix-based solver (linalg/solvers/matrixbased.py:751-775) as two (capacity, n) matrices in
    HBM.  ``append`` adds a pair, ``apply(k, mode, x)`` multiplies the sum of the first k updates with
    the columns of a host array; modes as in ``pncov_chain_apply``."""

    SYMMETRIC_RANK2, RANK1, RANK1_TRANSPOSED = 0, 1, 2

    def __init__(self, n: int, capacity: int = 64, *, device: Optional[int] = None):
        self._lib = load()
        self._h = _c_vp()
        self.n, self.capacity = int(n), int(capacity)
        self.device = default_device() if device is None else int(device)
        self.pairs = []  # host copies: growth, copy-on-write, dense views
        _check(self._lib.pncov_chain_create(ctypes.byref(self._h), self.device, self.n, self.capacity))

    def __len__(self) -> int:
        return len(self.pairs)

    def append(self, u: np.ndarray, v: np.ndarray) -> int:
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(self.n))
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(self.n))
        if len(self.pairs) == self.capacity:  # grow: new device matrices, vectors uploaded again
            old = self._h
            self._h = _c_vp()
            if self.capacity >= CHAIN_MAX_CAPACITY:
                raise MemoryError(f"an update chain holds at most {CHAIN_MAX_CAPACITY} rank-1 updates")
            self.capacity = min(2 * self.capacity, CHAIN_MAX_CAPACITY)
            _check(self._lib.pncov_chain_create(ctypes.byref(self._h), self.device, self.n, self.capacity))
            self._lib.pncov_chain_destroy(old)
            for i, (a, b) in enumerate(self.pairs):
                _check(self._lib.pncov_chain_set(self._h, i, _ptr(a), _ptr(b)))
        _check(self._lib.pncov_chain_set(self._h, len(self.pairs), _ptr(u), _ptr(v)))
        self.pairs.append((u, v))
        return len(self.pairs)

    def apply(self, k: int, mode: int, x: np.ndarray) -> np.ndarray:
        """``C_k @ x`` for a host array of shape (n, c)."""
        x2 = np.ascontiguousarray(x, dtype=np.float64)
        if x2.ndim != 2 or x2.shape[0] != self.n:
            raise ValueError(f"expected an operand of shape ({self.n}, c)")
        if not 0 <= k <= len(self.pairs):
            raise ValueError("chain length out of range")
        y = np.empty_like(x2)
        c = x2.shape[1]
        if c:
            _check(self._lib.pncov_chain_apply(self._h, k, mode, _ptr(x2), c, c, _ptr(y), c))
        return y

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pncov_chain_destroy(self._h)
            self._h = _c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass
