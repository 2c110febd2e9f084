"""Golden fixtures for the update operators of the matrix-based solver (SURVEY.md 8f N2 ii), from the
REAL reference (build container only; see make_golden.py for how the reference is imported).

    python tests/golden/make_golden_solver.py

Writes ``tests/golden/solver_golden.npz``: vector pairs, an operand and the products of the sums of the
reference's own ``_mean_update`` / ``_covariance_update`` LambdaLinearOperators with it.
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main() -> None:
    import_reference()
    from probnum.linalg.solvers import matrixbased as mb  # noqa: PLC0415

    rng = np.random.default_rng(77)
    out = {}
    for name, n, k, c in (("a", 37, 1, 1), ("b", 200, 6, 3), ("c", 513, 20, 2)):
        U, V, x = rng.normal(size=(n, k)), rng.normal(size=(n, k)), rng.normal(size=(n, c))
        me = types.SimpleNamespace(n=n)
        mean = cov = None
        for i in range(k):
            u, v = U[:, i : i + 1], V[:, i : i + 1]
            m_i = mb.SymmetricMatrixBasedSolver._mean_update(me, u=u, v=v)
            c_i = mb.SymmetricMatrixBasedSolver._covariance_update(me, u=u, Ws=v)
            mean = m_i if mean is None else mean + m_i
            cov = c_i if cov is None else cov + c_i
        out.update({f"{name}/U": U, f"{name}/V": V, f"{name}/x": x, f"{name}/mean": mean @ x, f"{name}/cov": cov @ x,
                    f"{name}/cov_T": cov.T @ x})
    out["names"] = np.array(["a", "b", "c"])
    np.savez_compressed(os.path.join(HERE, "solver_golden.npz"), **out)
    print("wrote solver_golden.npz")


if __name__ == "__main__":
    main()
