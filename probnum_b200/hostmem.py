"""Pinned (page-locked) host arrays for large results.

``CovarianceFunction.matrix`` must hand a NumPy array back to the caller
(/root/reference/src/probnum/randprocs/covfuncs/_covariance_function.py:407-411).
A Gram matrix of several GiB leaves the device at PCIe rate only if the destination
is page-locked, so large result arrays are carved from ``cudaHostAlloc`` blocks
owned by this module.  A block returns to a small free list when the array that
wraps it is garbage collected, which keeps repeated calls from paying the pinning
cost again.  Small results use ordinary pageable memory.
"""

from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib

PINNED_THRESHOLD_BYTES = 4 << 20
_MAX_CACHED_BYTES = 20 << 30
_free = {}  # nbytes -> [ptr, ...]
_cached_bytes = 0
_lock = threading.Lock()


class _PinnedBlock:
    """Owner object exposed to NumPy through ``__array_interface__``."""

    def __init__(self, nbytes: int, shape, dtype):
        global _cached_bytes
        self.nbytes = nbytes
        ptr = None
        with _lock:
            lst = _free.get(nbytes)
            if lst:
                ptr = lst.pop()
                _cached_bytes -= nbytes
        if ptr is None:
            p = ctypes.c_void_p()
            _lib._check(_lib.load().pncov_host_alloc(ctypes.byref(p), nbytes))
            ptr = p.value
        self.ptr = ptr
        self.__array_interface__ = {
            "data": (ptr, False),
            "shape": tuple(shape),
            "typestr": np.dtype(dtype).str,
            "version": 3,
        }

    def __del__(self):
        global _cached_bytes
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr is None:
            return
        try:
            with _lock:
                if _cached_bytes + self.nbytes <= _MAX_CACHED_BYTES:
                    _free.setdefault(self.nbytes, []).append(ptr)
                    _cached_bytes += self.nbytes
                    return
            _lib.load().pncov_host_free(ctypes.c_void_p(ptr))
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


def empty(shape, dtype) -> np.ndarray:
    """Uninitialised C-contiguous array; pinned if it is large."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < PINNED_THRESHOLD_BYTES:
        return np.empty(shape, dtype=dtype)
    return np.asarray(_PinnedBlock(nbytes, shape, dtype))


def release_cache() -> None:
    """Free every cached pinned block."""
    global _cached_bytes
    with _lock:
        blocks = [(n, p) for n, lst in _free.items() for p in lst]
        _free.clear()
        _cached_bytes = 0
    for _, p in blocks:
        _lib.load().pncov_host_free(ctypes.c_void_p(p))
