"""ncu driver: one warm-up + one timed launch of the wide-input Gram (N, d from argv; PROF_CROSS=1 for x1 != x0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402
from probnum_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
x = torch.from_numpy(np.random.default_rng(42).normal(size=(n, d))).to(dev)
x1 = torch.from_numpy(np.random.default_rng(43).normal(size=(n, d))).to(dev) if os.environ.get("PROF_CROSS") else None
out = torch.empty((n, n), dtype=torch.float64, device=dev)
k = pb.ExpQuad((d,), lengthscales=0.7 * d ** 0.5)
op = _lib.Operator(k._kernel_params(), x.data_ptr(), None if x1 is None else x1.data_ptr(), on_device=True, n0=n, n1=n, dtype=np.float64)
for _ in range(2):
    op.gram_device(out.data_ptr(), n, stream=stream)
torch.cuda.synchronize()
print("ExpQuad gram", "cross" if x1 is not None else "sym", n, d, "ms", op.last_kernel_ms())
op.close()
