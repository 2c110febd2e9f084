"""End-to-end time of k.matrix(X) / k.linop(X) @ v through the mirror API at small and medium sizes (host in, host out)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probnum_b200 as pb  # noqa: E402

rng = np.random.default_rng(0)
k = pb.Matern((3,), nu=2.5)
for n in (100, 1000, 1001, 4096, 8192, 8193, 16384):
    x = rng.normal(size=(n, 3))
    k.matrix(x)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); K = k.matrix(x); best = min(best, time.perf_counter() - t0); del K
    pcie = 8.0 * n * n / 55e9
    print(f"matrix N={n:6d}: {best * 1e3:8.3f} ms   (PCIe floor {pcie * 1e3:7.3f} ms)")
    y1 = rng.normal(size=(n // 2 + 1, 3))
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); K = k.matrix(x, y1); best = min(best, time.perf_counter() - t0); del K
    print(f"   cross ({n} x {n // 2 + 1}): {best * 1e3:8.3f} ms")
for n in (1000, 8192, 65536):
    x = rng.normal(size=(n, 4)); v = rng.normal(size=n)
    L = pb.ExpQuad((4,), lengthscales=0.5).linop(x)
    L @ v
    best = 1e9
    for _ in range(5):
        w = v + rng.normal()  # defeat the memo
        t0 = time.perf_counter(); y = L @ w; best = min(best, time.perf_counter() - t0)
    print(f"linop @ v N={n:6d}: {best * 1e3:8.3f} ms")
    L.close()
