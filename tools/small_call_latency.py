import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import probnum_b200 as pb
for n, d in [(100, 3), (1000, 10), (1000, 100), (4096, 3)]:
    X = np.random.default_rng(0).normal(size=(n, d))
    k = pb.ExpQuad((d,))
    k.matrix(X); k.matrix(X)
    t0 = time.perf_counter()
    for _ in range(50): k.matrix(X)
    dt = (time.perf_counter() - t0) / 50
    L = k.linop(X); v = np.ones(n); L @ v
    t0 = time.perf_counter()
    for _ in range(50): L @ v
    dm = (time.perf_counter() - t0) / 50
    print(f"N={n} d={d}: matrix() {dt*1e3:.3f} ms   linop@v {dm*1e3:.3f} ms")
