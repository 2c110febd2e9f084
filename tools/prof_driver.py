"""Short driver for ncu captures: one launch of each hot kernel after a warm-up launch.

    python tools/prof_driver.py [gram|matvec|all] [N_gram] [N_matvec]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
nm = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
reps = 1 if os.environ.get("PROF_SINGLE") else 2
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(42)
if what in ("gram", "all"):
    x = torch.from_numpy(rng.normal(size=(n, 3))).to(dev)
    x1 = torch.from_numpy(rng.normal(size=(n, 3))).to(dev) if os.environ.get("PROF_CROSS") else None
    out = torch.empty((n, n), dtype=torch.float64, device=dev)
    kernels = (pb.ExpQuad((3,)), pb.Matern((3,), nu=2.5), pb.RatQuad((3,)), pb.RatQuad((3,), alpha=1.5))
    sel = os.environ.get("PROF_KERNELS")  # e.g. "1,3": indices into `kernels`
    chosen = [kernels[int(i)] for i in sel.split(",")] if sel else (kernels[:2] if reps == 1 else kernels)
    for k in chosen:
        op = _lib.Operator(k._kernel_params(), x.data_ptr(), None if x1 is None else x1.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
        for _ in range(reps):
            op.gram_device(out.data_ptr(), n, stream=stream)
        torch.cuda.synchronize()
        print(type(k).__name__, "gram ms", op.last_kernel_ms())
        op.close()
    del out
if what == "extra":  # SURVEY.md 8f N3 kernels: Linear, ProductMatern Gram and the combine kernel
    import ctypes

    x = torch.from_numpy(rng.normal(size=(n, 3))).to(dev)
    out = torch.empty((n, n), dtype=torch.float64, device=dev)
    tmp = torch.empty((n, n), dtype=torch.float64, device=dev)
    for k in (pb.Linear((3,), constant=0.5), pb.ProductMatern((3,), lengthscales=1.0, nus=2.5)):
        op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
        for _ in range(reps):
            op.gram_device(out.data_ptr(), n, stream=stream)
        torch.cuda.synchronize()
        print(type(k).__name__, "gram ms", op.last_kernel_ms())
        op.close()
    for _ in range(reps):
        _lib.load().pncov_combine_device(0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(tmp.data_ptr()), n * n, 0,
                                         1.0, 0.5, None)
    torch.cuda.synchronize()
if what in ("matvec", "all"):
    x = torch.from_numpy(rng.normal(size=(nm, 8))).to(dev)
    v = torch.from_numpy(rng.normal(size=(nm, 1))).to(dev)
    y = torch.empty((nm, 1), dtype=torch.float64, device=dev)
    mv_kernels = (pb.ExpQuad((8,)), pb.Matern((8,), nu=2.5))
    for k in mv_kernels[:1] if reps == 1 else mv_kernels:
        op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=nm, n1=nm, dtype=np.float64)
        for _ in range(reps):
            op.matmul_device(v.data_ptr(), 1, 1, y.data_ptr(), 1, stream=stream)
        torch.cuda.synchronize()
        ms = op.last_kernel_ms()
        print(type(k).__name__, "matvec ms", ms, "Gpairs/s", nm * nm / ms / 1e6)
        op.close()
