"""K.v cliffs: float32 points, several columns, ARD lengthscales, rectangular operators: python tools/matvec_sweep2.py [N]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(0)


def run(label, k, d, dt, kcols, rect=False):
    x = torch.from_numpy(rng.normal(size=(n, d)).astype(dt)).to(dev)
    x1 = torch.from_numpy(rng.normal(size=(n, d)).astype(dt)).to(dev) if rect else None
    v = torch.from_numpy(rng.normal(size=(n, kcols)).astype(dt)).to(dev)
    y = torch.empty((n, kcols), dtype=v.dtype, device=dev)
    op = _lib.Operator(k._kernel_params(), x.data_ptr(), None if x1 is None else x1.data_ptr(), on_device=True, n0=n, n1=n, dtype=dt)
    for _ in range(2):
        op.matmul_device(v.data_ptr(), kcols, kcols, y.data_ptr(), kcols, stream=stream)
    torch.cuda.synchronize()
    ms = op.last_kernel_ms()
    print(f"{label:44s} {ms:9.3f} ms  {n * n / ms / 1e6:8.1f} Gpair/s x {kcols} col  path={op.matmul_path()}")
    op.close()


run("expquad d=8 f64 K=1 symmetric", pb.ExpQuad((8,), lengthscales=2.0), 8, np.float64, 1)
run("expquad d=8 f64 K=1 rectangular", pb.ExpQuad((8,), lengthscales=2.0), 8, np.float64, 1, rect=True)
run("expquad d=8 f64 K=4", pb.ExpQuad((8,), lengthscales=2.0), 8, np.float64, 4)
run("expquad d=8 f64 K=8", pb.ExpQuad((8,), lengthscales=2.0), 8, np.float64, 8)
run("expquad d=8 f64 ARD K=1", pb.ExpQuad((8,), lengthscales=np.linspace(1.5, 3.0, 8)), 8, np.float64, 1)
run("matern52 d=8 f64 ARD K=1", pb.Matern((8,), nu=2.5, lengthscales=np.linspace(1.5, 3.0, 8)), 8, np.float64, 1)
run("matern52 d=8 f64 K=4", pb.Matern((8,), nu=2.5, lengthscales=2.0), 8, np.float64, 4)
run("matern52 d=3 f64 K=1 rectangular", pb.Matern((3,), nu=2.5), 3, np.float64, 1, rect=True)
run("ratquad d=8 f64 K=1", pb.RatQuad((8,), lengthscale=2.0, alpha=1.5), 8, np.float64, 1)
run("expquad d=8 f32 K=1", pb.ExpQuad((8,), lengthscales=2.0), 8, np.float32, 1)
run("matern52 d=8 f32 K=1", pb.Matern((8,), nu=2.5, lengthscales=2.0), 8, np.float32, 1)
run("expquad d=3 f32 K=4", pb.ExpQuad((3,)), 3, np.float32, 4)
