"""Golden fixtures for the Bayesian-quadrature prediction path, from the REAL
reference (build container only; see make_golden.py for how the reference is imported).

    python tests/golden/make_golden_quad.py

Writes ``tests/golden/quad_golden.npz``: for each case the nodes, function values, prediction points,
jitter and -- from the reference's BQStandardBeliefUpdate / IntegralVarianceReduction -- the scale,
kernel means, integral belief, predictive mean / variance and acquisition values.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

CASES = [  # (name, n, m, d, lengthscale, jitter)
    ("small_d1", 12, 5, 1, 0.7, 1e-6),
    ("d2", 60, 17, 2, 0.5, 1e-5),
    ("d3_padded", 130, 33, 3, 0.8, 1e-4),
    ("d4", 200, 9, 4, 1.0, 1e-3),
]


def main() -> None:
    pn = import_reference()
    from probnum.quad.solvers import BQState  # noqa: PLC0415
    from probnum.quad.solvers.acquisition_functions import IntegralVarianceReduction  # noqa: PLC0415
    from probnum.quad.solvers.belief_updates import BQStandardBeliefUpdate  # noqa: PLC0415

    out = {"names": np.array([c[0] for c in CASES])}
    rng = np.random.default_rng(2024)
    for name, n, m, d, ell, jitter in CASES:
        nodes, x = rng.normal(size=(n, d)), rng.normal(size=(m, d))
        fun_evals = np.sin(nodes).sum(axis=1) + 0.1 * nodes[:, 0] ** 2
        measure = pn.quad.integration_measures.GaussianMeasure(mean=np.zeros(d), cov=np.eye(d))
        kernel = pn.randprocs.covfuncs.ExpQuad((d,), lengthscales=ell)
        state = BQState(measure=measure, kernel=kernel)
        belief, state = BQStandardBeliefUpdate(jitter=jitter, scale_estimation="mle")(state, nodes, fun_evals)
        mean, var = BQStandardBeliefUpdate.predict_integrand(x, state)
        acq, _ = IntegralVarianceReduction()(x, state)
        for key, val in dict(nodes=nodes, x=x, fun_evals=fun_evals, lengthscale=ell, jitter=jitter,
                             scale_sq=state.scale_sq, kernel_means=state.kernel_means,
                             kernel_mean_x=state.kernel_embedding.kernel_mean(x),
                             integral_mean=belief.mean, integral_cov=belief.cov,
                             predictive_mean=mean, predictive_var=var, acquisition=acq).items():
            out[f"{name}/{key}"] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, "quad_golden.npz"), **out)
    print("wrote quad_golden.npz:", len(CASES), "cases")


if __name__ == "__main__":
    main()
