"""Gram timings of the dense-only families over d: python tools/extra_sweep.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(0)
out = torch.empty((n, n), dtype=torch.float64, device=dev)
for d in (3, 16, 17, 40, 100):
    for name, k in (("linear", pb.Linear((d,), constant=0.5)), ("poly3", pb.Polynomial((d,), constant=1.0, exponent=3)),
                    ("whitenoise", pb.WhiteNoise((d,), 0.1)), ("prodmatern32", pb.ProductMatern((d,), lengthscales=2.0, nus=1.5))):
        x = torch.from_numpy(rng.normal(size=(n, d))).to(dev)
        op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
        for _ in range(3):
            op.gram_device(out.data_ptr(), n, stream=stream)
        torch.cuda.synchronize()
        ms = op.last_kernel_ms()
        ops = {"linear": d + 1, "poly3": d + 3, "whitenoise": d, "prodmatern32": 5 * d + 9}[name]
        print(f"d={d:3d} {name:12s} {ms:8.3f} ms   fp64-bound {n * n * ops / (33.9e12 / 2) * 1e3:7.3f}  hbm-bound {n * n * 8 / 6.5e12 * 1e3:6.3f}")
        op.close()
