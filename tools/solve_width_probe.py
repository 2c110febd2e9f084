"""pncov_chol_solve wall time by number of right-hand sides (8-column sweeps vs the transposed DMMA substitution).
    python tools/solve_width_probe.py [n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
B = torch.randn((n, n + 64), dtype=torch.float64, device="cuda")
A = B @ B.T / n
A.diagonal().add_(0.1)
del B
ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
rng = np.random.default_rng(0)
for k in (1, 2, 8, 9, 16, 64, 128, 129, 512):
    b = rng.normal(size=(n, k))
    ch.solve(b)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        x = ch.solve(b)
        best = min(best, time.perf_counter() - t0)
    print(f"n={n} nrhs={k:4d}: {best * 1e3:8.3f} ms  ({best * 1e3 / k:7.3f} ms per column)")
