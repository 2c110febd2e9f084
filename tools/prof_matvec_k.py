"""ncu driver: symmetric ExpQuad K @ V with several right-hand sides (N, K from argv; d = 8)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb
from probnum_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
rng = np.random.default_rng(42)
x = torch.from_numpy(rng.normal(size=(n, 8))).to(dev)
v = torch.from_numpy(rng.uniform(size=(n, k))).to(dev)
y = torch.empty((n, k), dtype=torch.float64, device=dev)
op = _lib.Operator(pb.ExpQuad((8,))._kernel_params(), x.data_ptr(), None, on_device=True, n0=n, n1=n, dtype=np.float64)
for _ in range(2):
    op.matmul_device(v.data_ptr(), k, k, y.data_ptr(), k, stream=stream)
torch.cuda.synchronize()
print("K", k, "ms", op.last_kernel_ms(), op.matmul_path())
