"""pytest configuration: the ``gpu`` marker and shared golden-fixture loading."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """Inputs and outputs recorded from the real reference (tests/golden/make_golden.py)."""

    def __init__(self):
        with open(os.path.join(GOLDEN_DIR, "covfuncs_golden.json")) as fh:
            meta = json.load(fh)
        self.kat = meta["kat"]
        self.cases = meta["cases"]
        self.arrays = np.load(os.path.join(GOLDEN_DIR, "covfuncs_golden.npz"))

    def get(self, name):
        return None if name is None else self.arrays[name]


_GOLDEN = None


def load_golden():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


@pytest.fixture(scope="session")
def golden():
    return load_golden()


def spec_from_dict(d):
    from oracle.covfuncs_oracle import KernelSpec

    ls = d.get("lengthscales", 1.0)
    return KernelSpec(
        kind=d["kind"],
        input_shape=tuple(d["input_shape"]),
        lengthscales=np.asarray(ls) if np.ndim(ls) else float(ls),
        nu=float(d.get("nu", 1.5)),
        alpha=float(d.get("alpha", 1.0)),
    )
