"""The NumPy oracle must reproduce the real reference bit-for-bit.

Golden data come from ``tests/golden/make_golden.py`` (run against
/root/reference in the build container).  These tests pin the oracle; the GPU
parity tests then compare the CUDA path with the pinned oracle.
"""

import numpy as np
import pytest

from oracle import covfuncs_oracle as orc
from tests.conftest import load_golden, spec_from_dict

G = load_golden()


@pytest.mark.parametrize("entry", G.kat, ids=[e["cite"].split("/")[-1] for e in G.kat])
def test_docstring_known_answers(entry):
    """Doctest values printed in the reference (8 significant digits)."""
    spec = spec_from_dict(entry["spec"])
    x = G.get(entry["x"])
    got = orc.gram(spec, x)
    np.testing.assert_allclose(got, np.asarray(entry["doc"]), rtol=5e-8)
    np.testing.assert_array_equal(got, G.get(entry["ref"]))


GRAM_CASES = [c for c in G.cases if c["type"] == "gram"]
CALL_CASES = [c for c in G.cases if c["type"] == "call"]


def test_golden_is_not_empty():
    assert len(GRAM_CASES) > 300 and len(CALL_CASES) > 100


@pytest.mark.parametrize("chunk", range(8))
def test_gram_bit_identical_to_reference(chunk):
    for case in GRAM_CASES[chunk::8]:
        spec = spec_from_dict(case["spec"])
        x0, x1 = G.get(case["x0"]), G.get(case["x1"])
        want = G.get(case["K"])
        got = orc.gram(spec, x0, x1, row_block=7)  # odd block: blocking must not matter
        assert got.dtype == want.dtype, case
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("chunk", range(4))
def test_call_bit_identical_to_reference(chunk):
    for case in CALL_CASES[chunk::4]:
        spec = spec_from_dict(case["spec"])
        got = orc.evaluate(spec, G.get(case["x0"]), G.get(case["x1"]))
        want = G.get(case["out"])
        assert np.shape(got) == want.shape
        assert np.asarray(got).dtype == want.dtype
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("chunk", range(4))
def test_matvec_matches_reference(chunk):
    """Reference: todense() @ v (BLAS); oracle: row-blocked.  Summation order may
    differ, so the componentwise bound of SURVEY 8d applies, not bit equality."""
    for case in GRAM_CASES[chunk::4]:
        spec = spec_from_dict(case["spec"])
        x0, x1, v = G.get(case["x0"]), G.get(case["x1"]), G.get(case["v"])
        got = orc.matvec(spec, x0, x1, v, row_block=16)
        want = G.get(case["Kv"])
        bound = orc.abs_matvec(spec, x0, x1, v)
        tol = 1e-12 if case["dtype"] == "float64" else 1e-5
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= tol * bound + 1e-300)


def test_matern_coefficients_exact():
    # p = 2: 1 + r + r^2/3  (_matern.py:61-64 closed forms)
    np.testing.assert_array_equal(orc.matern_coefficients(0), [1.0])
    np.testing.assert_array_equal(orc.matern_coefficients(1), [1.0, 1.0])
    np.testing.assert_array_equal(orc.matern_coefficients(2), [1.0, 1.0, 1.0 / 3.0])
    np.testing.assert_allclose(orc.matern_coefficients(3), [1.0, 1.0, 0.4, 1.0 / 15.0])
    assert orc.matern_p(2.5) == 2 and orc.matern_p(0.5) == 0 and orc.matern_p(0.7) is None
