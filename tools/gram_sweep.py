"""Gram timings away from the benchmark shape: larger d (strip / chunk kernels), float32 points, cross matrices.
    python tools/gram_sweep.py [N]
"""
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
for dt in (np.float64, np.float32):
    for d in (3, 8, 10, 16, 32, 100):
        for name, k in (("expquad", pb.ExpQuad((d,))), ("matern52", pb.Matern((d,), nu=2.5))):
            x = torch.from_numpy(rng.normal(size=(n, d)).astype(dt)).to(dev)
            op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=dt)
            out = torch.empty((n, n), dtype=torch.float32 if op.gram_dtype == np.float32 else torch.float64, device=dev)
            for _ in range(3):
                op.gram_device(out.data_ptr(), n, stream=stream)
            torch.cuda.synchronize()
            ms = op.last_kernel_ms()
            fp64 = (2 * d if name == "expquad" else 2 * d) + (7 if name == "expquad" else 15)
            bound_ms = n * n * fp64 / (33.9e12 / 2) * 1e3 if dt == np.float64 else float("nan")
            hbm_ms = n * n * out.element_size() / 6.5e12 * 1e3
            print(f"{np.dtype(dt).name:8s} d={d:3d} {name:9s} {ms:8.3f} ms   fp64-bound {bound_ms:7.3f}  hbm-bound {hbm_ms:6.3f}")
            op.close()
            del out
