"""One factorisation (+ one 2-column solve) of a random SPD matrix (ncu driver): python tools/chol_probe.py [n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
B = torch.randn((n, n + 64), dtype=torch.float64, device="cuda")
A = B @ B.T / n
A.diagonal().add_(0.1)
torch.cuda.synchronize()
ch = _lib.Cholesky(A.data_ptr(), on_device=True, n=n)
print("factor ms", ch.last_factor_ms())
x = ch.solve(np.ones((n, 2)))
ch.close()
