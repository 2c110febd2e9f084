"""Driver for an ncu capture of chol_diag_kernel (csrc/pncov_chol.cu): one factorisation of n = 1024 (8 diagonal blocks).

    ncu --set full --import-source on -k regex:chol_diag --launch-skip 2 -c 1 -o out python tools/prof_chol_diag.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(0)
B = rng.normal(size=(n, n + 8))
A = B @ B.T / n + 0.1 * np.eye(n)
ch = _lib.Cholesky(A)
print("factor ms", ch.last_factor_ms())
ch.close()
