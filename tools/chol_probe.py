import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from probnum_b200 import _lib
n = 1024
rng = np.random.default_rng(0)
B = rng.normal(size=(n, n + 8)); A = B @ B.T / n + 0.1 * np.eye(n)
ch = _lib.Cholesky(A); print(ch.last_factor_ms()); ch.close()
