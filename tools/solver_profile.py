"""cProfile of problinsolve (BASELINE configs[3]) with the device operator: where the host time goes.
    python tools/solver_profile.py [n] [maxiter]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import probnum_b200.patch as gpu  # noqa: E402

gpu.numpy2_compat()
warnings.simplefilter("ignore")
import probnum as pn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(42)
X = rng.normal(size=(n, 4))
y = rng.normal(size=n)
gpu.enable()
k = pn.randprocs.covfuncs.ExpQuad((4,), lengthscales=0.25)


def run():
    A = k.linop(X) + 1e-3 * pn.linops.Identity(n)
    t0 = time.perf_counter()
    x, _, _, info = pn.linalg.problinsolve(A, y, maxiter=maxiter)
    return time.perf_counter() - t0, info


print("warm-up", run()[0])
for chains in (False, True):
    if chains:
        gpu.enable_solver_chains()  # update chains in HBM
        print("warm-up (chains)", run()[0])
    pr = cProfile.Profile()
    pr.enable()
    t, info = run()
    pr.disable()
    print("timed", "device chains" if chains else "reference lambdas", t, info["iter"], info["resid_l2norm"])
    pstats.Stats(pr).sort_stats("tottime").print_stats(10)
