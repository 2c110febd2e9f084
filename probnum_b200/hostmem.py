"""Pinned (page-locked) host arrays for large results.

``CovarianceFunction.matrix`` must hand a NumPy array back to the caller
(/root/reference/src/probnum/randprocs/covfuncs/_covariance_function.py:407-411).
A Gram matrix of several GiB leaves the device at PCIe rate only if the destination
is page-locked, so large result arrays are carved from page-locked blocks owned by this
module (``pncov_host_alloc``: an anonymous mapping advised towards huge pages, populated by
all host worker threads of the library and registered with the driver -- 0.3-0.6 s for
8.6 GB where ``cudaHostAlloc`` takes 3.1-3.5 s).  A block returns to a small free list when the array that
wraps it is garbage collected, which keeps repeated calls from paying the pinning
cost again.  Small results use ordinary pageable memory.
"""

from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

from . import _lib

PINNED_THRESHOLD_BYTES = 4 << 20
# Page-locked memory kept on the free list per process after its arrays are gone (unswappable: with one process
# per GPU this is multiplied by the number of ranks).  PNCOV_PINNED_CACHE_GIB overrides the default of 18 GiB,
# which holds the two 8.6 GB results of BASELINE configs[1]; 0 disables the cache.
_MAX_CACHED_BYTES = int(float(os.environ.get("PNCOV_PINNED_CACHE_GIB", "18")) * (1 << 30))
_free = {}  # nbytes -> [ptr, ...] (dict order = insertion order: the oldest size class is evicted first)
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
            evicted = []
            with _lock:
                if self.nbytes <= _MAX_CACHED_BYTES:
                    # make room by dropping the oldest cached blocks (differently sized results never reuse them)
                    while _cached_bytes + self.nbytes > _MAX_CACHED_BYTES and _free:
                        size = next(iter(_free))
                        lst = _free[size]
                        evicted.append(lst.pop(0))
                        _cached_bytes -= size
                        if not lst:
                            del _free[size]
                    _free.setdefault(self.nbytes, []).append(ptr)
                    _cached_bytes += self.nbytes
                    ptr = None
            for p in evicted:
                _lib.load().pncov_host_free(ctypes.c_void_p(p))
            if ptr is not None:
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
