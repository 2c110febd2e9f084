"""Driver for ncu captures of the Cholesky kernels: one factorisation of order n and one single-column solve.

    python tools/prof_chol_solve.py [n]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(n)
B = torch.randn((n, n + 64), dtype=torch.float64, device="cuda")
A = B @ B.T / n
A.diagonal().add_(0.1)
del B
ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
b = np.random.default_rng(0).normal(size=(n, 1))
x = ch.solve(b)
r = (A @ torch.from_numpy(x).cuda()).cpu().numpy() - b
print("factor ms", ch.last_factor_ms(), "residual", float(np.linalg.norm(r) / np.linalg.norm(b)))
ch.close()
