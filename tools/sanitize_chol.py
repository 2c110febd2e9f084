"""compute-sanitizer driver for the Cholesky kernels (factor, chained substitution, upper-triangle upload) at small sizes.

    compute-sanitizer --tool memcheck  python tools/sanitize_chol.py
    compute-sanitizer --tool racecheck python tools/sanitize_chol.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
for n in (5, 128, 300, 1000):
    B = rng.normal(size=(n, n + 8))
    A = B @ B.T / n + 0.1 * np.eye(n)
    for upper in (False, True):
        ch = _lib.Cholesky(A, jitter=1e-9, upper=upper)
        for k in (1, 2, 3, 7, 14):
            b = rng.normal(size=(n, k))
            x = ch.solve(b)
            assert np.linalg.norm(A @ x - b) <= 1e-8 * np.linalg.norm(b) * n
        ch.close()
print("sanitize chol done")
