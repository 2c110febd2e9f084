"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the update operators of the matrix-based linear solver
(SURVEY.md 8f N2 ii).  Nothing under ``probnum_b200/`` imports this module.

Parity: PINNED.  ``tests/test_oracle_solver.py`` checks both functions against
``tests/golden/solver_golden.npz``, recorded from the real reference by
``tests/golden/make_golden_solver.py``.

Reference (relative to /root/reference/src/probnum): linalg/solvers/matrixbased.py
  :751-762  _mean_update        x -> u (v^T x) + v (u^T x)
  :764-775  _covariance_update  x -> Ws (u^T x)
  :897-924  the solver sums one such operator per iteration
"""

from __future__ import annotations

import numpy as np


def mean_update_chain(U: np.ndarray, V: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Sum over the columns i of ``U``, ``V`` (n, k) of the rank-2 updates :751-762 applied to ``x``
    (n, c), accumulated in iteration order like the solver's SumLinearOperator."""
    y = np.zeros_like(x, dtype=np.float64)
    for i in range(U.shape[1]):
        u, v = U[:, i : i + 1], V[:, i : i + 1]
        y = y + (u @ (v.T @ x) + v @ (u.T @ x))
    return y


def covariance_update_chain(U: np.ndarray, Ws: np.ndarray, x: np.ndarray, transpose: bool = False) -> np.ndarray:
    """Sum of the rank-1 updates :764-775, ``Ws_i (u_i^T x)``; ``transpose``: ``u_i (Ws_i^T x)``."""
    y = np.zeros_like(x, dtype=np.float64)
    for i in range(U.shape[1]):
        u, w = U[:, i : i + 1], Ws[:, i : i + 1]
        y = y + (u @ (w.T @ x) if transpose else w @ (u.T @ x))
    return y
