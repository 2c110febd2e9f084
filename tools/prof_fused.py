"""ncu driver: the three-leaf sum 2*Matern52 + 0.1*ExpQuad + WhiteNoise in one launch (N = 32768, d = 3)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb
from probnum_b200 import _lib, covfuncs
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 3
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
x = torch.from_numpy(np.random.default_rng(42).normal(size=(n, d))).to(dev)
out = torch.empty((n, n), dtype=torch.float64, device=dev)
k = 2.0 * pb.Matern((d,), nu=2.5) + 0.1 * pb.ExpQuad((d,)) + pb.WhiteNoise((d,), sigma_sq=0.5)
leaves, steps, params = covfuncs._compile_program(k)
op = _lib.Operator(params, x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
op.set_program(leaves, steps)
for _ in range(2):
    op.gram_device(out.data_ptr(), n, stream=stream)
torch.cuda.synchronize()
print("fused ms", op.last_kernel_ms())
