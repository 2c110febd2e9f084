"""K.v timings over the input dimension (main kernels d <= 8, generic kernel above): python tools/matvec_sweep.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(0)
for d in (4, 8, 9, 12, 16, 32, 64):
    for name, k in (("expquad", pb.ExpQuad((d,), lengthscales=np.sqrt(d))), ("matern52", pb.Matern((d,), nu=2.5, lengthscales=np.sqrt(d)))):
        x = torch.from_numpy(rng.normal(size=(n, d))).to(dev)
        v = torch.from_numpy(rng.normal(size=(n, 1))).to(dev)
        y = torch.empty((n, 1), dtype=torch.float64, device=dev)
        op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
        for _ in range(2):
            op.matmul_device(v.data_ptr(), 1, 1, y.data_ptr(), 1, stream=stream)
        torch.cuda.synchronize()
        ms = op.last_kernel_ms()
        print(f"d={d:3d} {name:9s} {ms:9.3f} ms  {n * n / ms / 1e6:8.1f} Gpair/s  path={op.matmul_path()}")
        op.close()
