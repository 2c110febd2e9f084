"""The C-ABI shared library loads and exports every symbol that include/pncov.h declares
(no compute calls: this runs without a GPU), and the product path fails loudly when no
device is present instead of falling back to a CPU implementation."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import probnum_b200 as pb
from probnum_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pncov.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PNCOV_API\s+[\w\s\*]+?\b(pncov_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(_lib.LIB_PATH).startswith(os.path.join(ROOT, "probnum_b200"))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in pncov.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in probnum_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (pncov_\w+)", nm)))
    assert exported == names, "the library exports symbols that the header does not declare"


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.pncov_version() == 100
    # invalid arguments are reported through return codes + pncov_last_error, not exceptions in C
    n = ctypes.c_int(-1)
    lib.pncov_device_count(ctypes.byref(n))
    assert n.value >= 0
    rc = lib.pncov_gram(None, None, 0, 0)
    assert rc == -1 and b"NULL" in lib.pncov_last_error()


def test_kernel_is_sm100a_sass_with_tma():
    """The shipped binary carries sm_100a SASS (not PTX for JIT) and uses the TMA bulk-copy
    path (UBLKCP) in the hot kernels."""
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    obj = os.path.join(ROOT, "probnum_b200", "csrc", "build", "pncov_gram.o")
    if os.path.exists(obj):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        assert "UBLKCP" in sass and "SYNCS" in sass and "DFMA" in sass


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    k = pb.ExpQuad((3,))
    x = np.zeros((4, 3))
    with pytest.raises(_lib.PncovError, match="no CPU fallback"):
        k.matrix(x)
    with pytest.raises(_lib.PncovError):
        k.linop(x) @ np.ones(4)
    with pytest.raises(_lib.PncovError):
        k(x[:, None, :], x[None, :, :])
    with pytest.raises(_lib.PncovError):
        k(x[:2], x[2:])


def test_product_does_not_import_the_oracle():
    """Nothing under probnum_b200/ may reference oracle/ (the checker is test-only)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "probnum_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "covfuncs_oracle" not in text, f
