import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb
from probnum_b200 import _lib
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev).cuda_stream
n = 32768
x = torch.from_numpy(np.random.default_rng(0).normal(size=(n, 3)).astype(np.float32)).to(dev)
for k, odt in ((pb.ExpQuad((3,)), torch.float32), (pb.Matern((3,), nu=2.5), torch.float64), (pb.RatQuad((3,)), torch.float64)):
    out = torch.empty((n, n), dtype=odt, device=dev)
    op = _lib.Operator(k._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float32)
    for _ in range(3):
        op.gram_device(out.data_ptr(), n, stream=stream)
    torch.cuda.synchronize()
    ms = op.last_kernel_ms()
    print(type(k).__name__, "f32 in ->", odt, "gram ms", round(ms, 3), "GB/s", round(out.element_size() * n * n / ms / 1e6, 1))
    op.close(); del out
# odd size (unaligned leading dimension) float64
n = 32767
x = torch.from_numpy(np.random.default_rng(0).normal(size=(n, 3))).to(dev)
out = torch.empty((n, n), dtype=torch.float64, device=dev)
op = _lib.Operator(pb.ExpQuad((3,))._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
for _ in range(3):
    op.gram_device(out.data_ptr(), n, stream=stream)
torch.cuda.synchronize()
print("f64 ExpQuad N=32767 (odd ld, generic kernel) ms", round(op.last_kernel_ms(), 3))
