"""oracle/solver_oracle.py pinned against products of the reference's own update operators
(tests/golden/make_golden_solver.py); GPU part: pncov_chain_apply against the same fixtures."""

import os

import numpy as np
import pytest

from oracle import solver_oracle as so

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "solver_golden.npz"))
NAMES = [str(n) for n in GOLD["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference_operators(name):
    U, V, x = GOLD[f"{name}/U"], GOLD[f"{name}/V"], GOLD[f"{name}/x"]
    np.testing.assert_array_equal(so.mean_update_chain(U, V, x), GOLD[f"{name}/mean"])  # same ops, same order
    np.testing.assert_array_equal(so.covariance_update_chain(U, V, x), GOLD[f"{name}/cov"])
    np.testing.assert_allclose(so.covariance_update_chain(U, V, x, transpose=True), GOLD[f"{name}/cov_T"], rtol=1e-13,
                               atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_chain_matches_the_reference_fixtures(name):
    from probnum_b200 import _lib

    U, V, x = GOLD[f"{name}/U"], GOLD[f"{name}/V"], GOLD[f"{name}/x"]
    chain = _lib.UpdateChain(U.shape[0], capacity=4)
    for i in range(U.shape[1]):
        chain.append(U[:, i], V[:, i])
    k = U.shape[1]
    bound = np.abs(U) @ (np.abs(V).T @ np.abs(x)) + np.abs(V) @ (np.abs(U).T @ np.abs(x))
    for mode, key in ((chain.SYMMETRIC_RANK2, "mean"), (chain.RANK1, "cov"), (chain.RANK1_TRANSPOSED, "cov_T")):
        assert np.all(np.abs(chain.apply(k, mode, x) - GOLD[f"{name}/{key}"]) <= 1e-12 * bound)
    chain.close()
