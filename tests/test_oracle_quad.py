"""oracle/quad_oracle.py pinned against fixtures recorded from the real reference
(tests/golden/make_golden_quad.py): BQ belief update quantities, predict_integrand and the
integral-variance-reduction acquisition.  GPU part: pncov_chol_solve_gram against the same fixtures."""

import os

import numpy as np
import pytest

from oracle import covfuncs_oracle as orc
from oracle import quad_oracle as qo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "quad_golden.npz"))
NAMES = [str(n) for n in GOLD["names"]]


def _case(name):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    d = g["nodes"].shape[1]
    return g, orc.KernelSpec("expquad", (d,), float(g["lengthscale"]))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference(name):
    g, spec = _case(name)
    factor = qo.gram_cho_factor(orc.gram(spec, g["nodes"]), float(g["jitter"]))
    scale_sq = qo.mle_scale_sq(factor, g["fun_evals"])
    # same LAPACK, same order of operations: equal to the last bits (1e-13 leaves room for BLAS threading)
    np.testing.assert_allclose(scale_sq, g["scale_sq"], rtol=1e-13)
    weights = qo.gram_cho_solve(factor, g["kernel_means"])
    np.testing.assert_allclose(weights @ g["fun_evals"], g["integral_mean"], rtol=1e-12, atol=1e-14)
    mean, var = qo.predict_integrand(spec, g["nodes"], g["fun_evals"], g["x"], float(g["jitter"]), float(g["scale_sq"]))
    np.testing.assert_allclose(mean, g["predictive_mean"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(var, g["predictive_var"], rtol=1e-10, atol=1e-14)
    acq = qo.integral_variance_reduction(spec, g["nodes"], g["fun_evals"], g["x"], float(g["jitter"]), float(g["scale_sq"]),
                                         g["kernel_means"], g["kernel_mean_x"], float(g["integral_cov"]))
    np.testing.assert_allclose(acq, g["acquisition"], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("forward_only", [False, True])
def test_device_prediction_matches_the_reference_fixtures(name, forward_only):
    """pncov_chol_factor_gram + pncov_chol_solve_gram against values recorded from the reference's
    predict_integrand / IntegralVarianceReduction.  Tolerances: 1e-10 on the scale of the sums (a host
    run with a different factorisation path deviates by 1e-14 .. 1e-15 on these cases), 1e-6 on the
    acquisition, a quotient by variances down to 1e-7 (host deviation 5e-9)."""
    import probnum_b200 as pb
    from probnum_b200 import _lib

    g, spec = _case(name)
    d = g["nodes"].shape[1]
    k = pb.ExpQuad((d,), lengthscales=float(g["lengthscale"]))
    with _lib.Operator(k._kernel_params(), g["nodes"]) as sq:
        ch = _lib.Cholesky(sq, jitter=float(g["jitter"]))
    with _lib.Operator(k._kernel_params(), g["nodes"], g["x"]) as op:
        quad, dots = ch.solve_gram(op, [g["fun_evals"], g["kernel_means"]], forward_only=forward_only)
    scale_sq = float(g["scale_sq"])
    np.testing.assert_allclose(dots[0], g["predictive_mean"], rtol=0, atol=1e-10 * np.abs(g["fun_evals"]).max())
    var = scale_sq * (k(g["x"], g["x"]) - quad)
    np.testing.assert_allclose(var, g["predictive_var"], rtol=0, atol=1e-10 * scale_sq)
    emb = g["kernel_mean_x"] - dots[1]
    acq = (scale_sq * emb) ** 2 / (float(g["integral_cov"]) * var)
    np.testing.assert_allclose(acq, g["acquisition"], rtol=1e-6)
    # the mle scale through the device solve
    w = ch.solve(g["fun_evals"])
    np.testing.assert_allclose(g["fun_evals"] @ w / w.shape[0], scale_sq, rtol=1e-9)
    ch.close()
