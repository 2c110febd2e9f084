"""Wall time of one K @ v through the C ABI with host vectors next to the device time of its kernels.
    python tools/matvec_call_latency.py [n ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [8192, 65536, 262144]:
    x = rng.normal(size=(n, 4))
    k = pb.ExpQuad((4,), lengthscales=0.25)
    with _lib.Operator(k._kernel_params(), x) as op:
        v = rng.normal(size=(n, 1))
        op.matmul(v)
        walls, kern = [], []
        for _ in range(7):
            v = v + 1e-3
            t0 = time.perf_counter()
            op.matmul(v)
            walls.append(time.perf_counter() - t0)
            kern.append(op.last_kernel_ms())
        print(f"N={n:7d} path={op.matmul_path()} wall best {min(walls) * 1e3:7.3f} ms median {np.median(walls) * 1e3:7.3f} ms   "
              f"kernels {np.median(kern):7.3f} ms   pairs/s (kernel) {n * n / (np.median(kern) * 1e-3):.3e}")
