"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Bayesian-quadrature linear algebra around the
covariance hot path.  Nothing under ``probnum_b200/`` imports this module.

Parity: PINNED.  ``tests/test_oracle_quad.py`` checks every function against
``tests/golden/quad_golden.npz``, recorded from the real reference by
``tests/golden/make_golden_quad.py``.

Reference (paths relative to /root/reference/src/probnum):
  quad/solvers/belief_updates/_belief_update.py:61-103    compute_gram_cho_factor, gram_cho_solve
  quad/solvers/belief_updates/_standard_update.py:36-95   belief update (weights, scale, integral belief)
  quad/solvers/belief_updates/_standard_update.py:111-151 _estimate_scale, predict_integrand
  quad/solvers/acquisition_functions/_integral_variance_reduction.py:60-75   IVR acquisition values
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import scipy.linalg

from . import covfuncs_oracle as orc


def gram_cho_factor(gram: np.ndarray, jitter: float):
    """_belief_update.py:61-80: ``cho_factor(gram + jitter * eye)`` (upper factor, scipy's default)."""
    return scipy.linalg.cho_factor(gram + jitter * np.eye(gram.shape[0]))


def gram_cho_solve(factor, z: np.ndarray) -> np.ndarray:
    """_belief_update.py:82-103."""
    return scipy.linalg.cho_solve(factor, z)


def mle_scale_sq(factor, fun_evals: np.ndarray) -> float:
    """_standard_update.py:111-128, ``scale_estimation == "mle"``."""
    return fun_evals @ gram_cho_solve(factor, fun_evals) / fun_evals.shape[0]


def predict_integrand(spec: orc.KernelSpec, nodes: np.ndarray, fun_evals: np.ndarray, x: np.ndarray, jitter: float,
                      scale_sq: float) -> Tuple[np.ndarray, np.ndarray]:
    """_standard_update.py:131-151 for a zero-mean prior: predictive mean and variance at ``x``."""
    factor = gram_cho_factor(orc.gram(spec, nodes), jitter)
    predictive_var = orc.evaluate(spec, x, x)  # bq_state.kernel(x, x)
    kXx = orc.gram(spec, nodes, x)
    weights = gram_cho_solve(factor, kXx)
    predictive_mean = np.zeros(x.shape[0]) + weights.T @ (fun_evals - np.zeros(fun_evals.shape[0]))
    predictive_var = predictive_var - np.sum(weights * kXx, axis=0)
    return predictive_mean, scale_sq * predictive_var


def integral_variance_reduction(spec: orc.KernelSpec, nodes: np.ndarray, fun_evals: np.ndarray, x: np.ndarray,
                                jitter: float, scale_sq: float, kernel_means: np.ndarray, kernel_mean_x: np.ndarray,
                                integral_cov: float) -> np.ndarray:
    """_integral_variance_reduction.py:60-75.  ``kernel_means`` / ``kernel_mean_x`` are the kernel
    embedding at the nodes / at ``x`` (closed forms of quad/kernel_embeddings, outside this oracle)."""
    _, y_predictive_var = predict_integrand(spec, nodes, fun_evals, x, jitter, scale_sq)
    y_predictive_var = y_predictive_var + 0.0  # observation noise placeholder of the reference
    factor = gram_cho_factor(orc.gram(spec, nodes), jitter)
    weights = gram_cho_solve(factor, orc.gram(spec, nodes, x))
    predictive_embedding = kernel_mean_x - np.dot(kernel_means, weights)
    return (scale_sq * predictive_embedding) ** 2 / (integral_cov * y_predictive_var)
