"""Build libpncov.so (sm_100a) in-tree with nvcc.  Used by ``__graft_entry__.build()``.

The product never JIT-compiles at import time: ``probnum_b200._lib`` only loads the
prebuilt ``probnum_b200/csrc/libpncov.so`` and raises if it is missing.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(CSRC, "libpncov.so")
SOURCES = ["pncov_api.cu", "pncov_gram.cu", "pncov_gram_sym_lo.cu", "pncov_gram_sym_hi.cu", "pncov_gram_mma.cu", "pncov_matvec.cu", "pncov_extra.cu", "pncov_chol.cu", "pncov_host.cu"]
HEADERS = ["pncov_common.cuh", "pncov_math.cuh", "pncov_tables.h", "pncov_internal.h", "pncov_gram_tile.cuh", "pncov_gram_sym.inc",
           os.path.join(ROOT, "include", "pncov.h")]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libpncov.so must be built where CUDA 12.9 is installed")
    return nvcc


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = find_nvcc()
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and _newer(obj, [path] + headers):
            return obj, False
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj, True

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"built {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
