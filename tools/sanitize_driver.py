"""Small driver for compute-sanitizer: touches every kernel family once at tiny sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402

rng = np.random.default_rng(0)
for d in (3, 8, 12, 40):
    x0, x1 = rng.normal(size=(1100, d)), rng.normal(size=(700, d))
    v = rng.normal(size=(700, 3))
    for k in (pb.ExpQuad((d,), lengthscales=np.sqrt(d)), pb.Matern((d,), nu=2.5, lengthscales=rng.gamma(2.0, size=d) + 1),
              pb.RatQuad((d,), lengthscale=np.sqrt(d), alpha=1.7)):
        k.matrix(x0, x1)          # fast interior + generic edges (d <= 8), strip (d <= 16) or chunk kernel
        k.matrix(x0[:1024])       # symmetric mirror kernel for the FP64-heavy epilogues
        k.matrix(x0[:333])        # odd leading dimension
        L = k.linop(x0, x1)
        L @ v                     # norm-identity DMMA / direct / generic
        L.T @ rng.normal(size=1100)
        k(x0[:50], x1[:50])       # pairs
        L.close()
x32 = rng.normal(size=(600, 3)).astype(np.float32)
for k in (pb.ExpQuad((3,)), pb.Matern((3,), nu=1.5)):
    k.matrix(x32)
    k.linop(x32) @ np.ones(600, dtype=np.float32)
os.environ["PNCOV_MATVEC_IDENT"] = "dfma"
pb.ExpQuad((4,)).linop(rng.normal(size=(900, 4))) @ np.ones(900)
print("sanitize driver done")
