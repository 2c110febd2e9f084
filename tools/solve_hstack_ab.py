"""Where the wall time of the reference's `problinsolve` goes once the products run on the device.

Times configs[3] (N = 65536, d = 4, 50 iterations) through probnum (baseline/_ref) + patch.enable_all(), with and
without the growing-hstack shim of patch.enable_solver_chains(), and prints a cProfile of one solve.

    python tools/solve_hstack_ab.py [--n 65536] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import numpy as np  # noqa: E402

import probnum_b200.patch as gpu  # noqa: E402

gpu.numpy2_compat()
warnings.simplefilter("ignore")
import probnum as pn  # noqa: E402
from probnum.linalg.solvers import matrixbased as mb  # noqa: E402

from bench import _solve  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    gpu.enable_all(sharded=False)
    _solve(pn, 4096, 0.25, 5)
    shim = mb.np
    xs = {}
    for name, mod in (("shim", shim), ("numpy", np), ("shim", shim), ("numpy", np)):
        mb.np = mod
        x, dt, it, res, _ = _solve(pn, args.n, 0.25, 50)
        xs.setdefault(name, x)
        print(f"{name:6s} n={args.n} iterations={it} seconds={dt:.4f} resid={res:.3e}", flush=True)
    print("iterates identical:", bool(np.array_equal(xs["shim"], xs["numpy"])))
    if args.profile:
        mb.np = shim
        pr = cProfile.Profile()
        pr.enable()
        _solve(pn, args.n, 0.25, 50)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    mb.np = shim
    gpu.disable()


if __name__ == "__main__":
    main()
