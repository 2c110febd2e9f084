"""Device-timed covariance arithmetic at N = 32768, d = 3: one launch (pncov_set_program) vs the
operand-by-operand device path, and the ProductMatern Gram."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib, covfuncs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = 3
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
x = torch.from_numpy(np.random.default_rng(42).normal(size=(n, d))).to(dev)
x1 = torch.from_numpy(np.random.default_rng(43).normal(size=(n, d))).to(dev)
out = torch.empty((n, n), dtype=torch.float64, device=dev)
tmp = torch.empty((n, n), dtype=torch.float64, device=dev)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


exprs = {
    "2*Matern52 + WhiteNoise": 2.0 * pb.Matern((d,), nu=2.5) + pb.WhiteNoise((d,), sigma_sq=0.5),
    "WhiteNoise + 0.5*ExpQuad": pb.WhiteNoise((d,), sigma_sq=0.5) + 0.5 * pb.ExpQuad((d,)),
    "2*Matern52 + 0.1*ExpQuad + WhiteNoise": 2.0 * pb.Matern((d,), nu=2.5) + 0.1 * pb.ExpQuad((d,)) + pb.WhiteNoise((d,), sigma_sq=0.5),
    "Matern32 * ExpQuad": pb.Matern((d,), nu=1.5) * pb.ExpQuad((d,)),
    "ExpQuad + RatQuad": pb.ExpQuad((d,)) + pb.RatQuad((d,)),
}
for name, k in exprs.items():
    leaves, steps, params = covfuncs._compile_program(k)
    for label, other in (("sym", None), ("cross", x1)):
        op = _lib.Operator(params, x.data_ptr(), None if other is None else other.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
        op.set_program(leaves, steps)
        ms = timed(lambda: op.gram_device(out.data_ptr(), n, stream=stream))
        print(f"fused {name} {label} n={n} ms={ms:.4f} GB/s={8.0*n*n/ms/1e6:.0f}", flush=True)
        op.close()
# operand-wise: 3 Gram launches + 2 combine passes for the first expression (device resident, no host copy)
import ctypes
lib = _lib.load()
ops = [_lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
       for k in (pb.Matern((d,), nu=2.5), pb.ExpQuad((d,)), pb.WhiteNoise((d,), sigma_sq=0.5))]


def operandwise():
    ops[0].gram_device(out.data_ptr(), n, stream=stream)
    lib.pncov_combine_device(0, ctypes.c_void_p(out.data_ptr()), None, n * n, 0, 2.0, 0.0, ctypes.c_void_p(stream))
    ops[1].gram_device(tmp.data_ptr(), n, stream=stream)
    lib.pncov_combine_device(0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(tmp.data_ptr()), n * n, 0, 1.0, 0.1, ctypes.c_void_p(stream))
    ops[2].gram_device(tmp.data_ptr(), n, stream=stream)
    lib.pncov_combine_device(0, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(tmp.data_ptr()), n * n, 0, 1.0, 1.0, ctypes.c_void_p(stream))


print(f"operand-wise 2*Matern52 + 0.1*ExpQuad + WhiteNoise sym n={n} ms={timed(operandwise):.4f}", flush=True)
pm = pb.ProductMatern((d,), lengthscales=1.0, nus=2.5)
for label, other in (("sym", None), ("cross", x1)):
    op = _lib.Operator(pm._kernel_params(), x.data_ptr(), None if other is None else other.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
    ms = timed(lambda: op.gram_device(out.data_ptr(), n, stream=stream))
    print(f"ProductMatern52 {label} n={n} ms={ms:.4f} GB/s={8.0*n*n/ms/1e6:.0f}", flush=True)
    op.close()
