"""Device time of the symmetric direct-difference K.v kernel (Matern, and ExpQuad on clouds too wide for the norm identity).

    python tools/matvec_direct_bench.py [n]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(0)
for d in (3, 4, 8):
    x = rng.normal(size=(n, d))
    for name, k in (("expquad l=0.25", pb.ExpQuad((d,), lengthscales=0.25)), ("matern52", pb.Matern((d,), nu=2.5, lengthscales=1.0)),
                    ("matern12", pb.Matern((d,), nu=0.5, lengthscales=1.0))):
        with _lib.Operator(k._kernel_params(), x) as op:
            v = rng.normal(size=(n, 1))
            op.matmul(v)
            ms = []
            for _ in range(3):
                op.matmul(v)
                ms.append(op.last_kernel_ms())
            print(f"d={d} {name:15s} path={op.matmul_path():32s} kernels {min(ms):8.3f} ms  {n * n / (min(ms) * 1e-3):.3e} ordered pairs/s", flush=True)
