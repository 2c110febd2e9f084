"""The extra oracle (Linear, Polynomial, WhiteNoise, ProductMatern, covariance arithmetic) must
reproduce the real reference bit-for-bit on the golden fixtures of make_golden_extra.py."""

import numpy as np
import pytest

from oracle import covfuncs_extra_oracle as xo
from tests.conftest import load_golden_extra

G = load_golden_extra()
GRAM_CASES = [c for c in G.cases if c["type"] == "gram"]
CALL_CASES = [c for c in G.cases if c["type"] == "call"]


def test_golden_is_not_empty():
    assert len(GRAM_CASES) > 200 and len(CALL_CASES) > 100


@pytest.mark.parametrize("entry", G.kat, ids=[e["cite"].split("/")[-1] for e in G.kat])
def test_docstring_known_answers(entry):
    np.testing.assert_array_equal(xo.gram(entry["spec"], G.get(entry["x"])), np.asarray(entry["doc"]))


@pytest.mark.parametrize("chunk", range(6))
def test_gram_bit_identical_to_reference(chunk):
    for case in GRAM_CASES[chunk::6]:
        want = G.get(case["K"])
        got = xo.gram(case["spec"], G.get(case["x0"]), G.get(case["x1"]), row_block=7)
        assert got.dtype == want.dtype and got.shape == want.shape, case["spec"]
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("chunk", range(3))
def test_call_bit_identical_to_reference(chunk):
    for case in CALL_CASES[chunk::3]:
        want = G.get(case["out"])
        got = np.asarray(xo.evaluate(case["spec"], G.get(case["x0"]), G.get(case["x1"])))
        assert got.shape == want.shape and got.dtype == want.dtype, case["spec"]
        np.testing.assert_array_equal(got, want)
