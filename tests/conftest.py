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

    def __init__(self, stem="covfuncs_golden"):
        with open(os.path.join(GOLDEN_DIR, stem + ".json")) as fh:
            meta = json.load(fh)
        self.kat = meta["kat"]
        self.cases = meta["cases"]
        self.arrays = np.load(os.path.join(GOLDEN_DIR, stem + ".npz"))

    def get(self, name):
        return None if name is None else self.arrays[name]


_GOLDEN = None


def load_golden():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


_GOLDEN_EXTRA = None


def load_golden_extra():
    """Fixtures of tests/golden/make_golden_extra.py (SURVEY.md 8f N3 families and arithmetic)."""
    global _GOLDEN_EXTRA
    if _GOLDEN_EXTRA is None:
        _GOLDEN_EXTRA = Golden("covfuncs_extra_golden")
    return _GOLDEN_EXTRA


def mirror_from_spec(spec):
    """probnum_b200 covariance function for a spec dict of oracle/covfuncs_extra_oracle.py."""
    import probnum_b200 as pb

    kind = spec["kind"]
    if kind == "scaled":
        return spec["scalar"] * mirror_from_spec(spec["child"])
    if kind in ("sum", "product"):
        parts = [mirror_from_spec(c) for c in spec["children"]]
        out = parts[0]
        for p in parts[1:]:
            out = out + p if kind == "sum" else out * p
        return out
    shape = tuple(spec["input_shape"])
    if kind == "expquad":
        return pb.ExpQuad(shape, lengthscales=np.asarray(spec["lengthscales"]))
    if kind == "matern":
        return pb.Matern(shape, nu=spec["nu"], lengthscales=np.asarray(spec["lengthscales"]))
    if kind == "ratquad":
        return pb.RatQuad(shape, lengthscale=spec["lengthscales"], alpha=spec["alpha"])
    if kind == "linear":
        return pb.Linear(shape, constant=spec["constant"])
    if kind == "polynomial":
        return pb.Polynomial(shape, constant=spec["constant"], exponent=spec["exponent"])
    if kind == "whitenoise":
        return pb.WhiteNoise(shape, sigma_sq=spec["sigma_sq"])
    if kind == "productmatern":
        ls, nus = spec["lengthscales"], spec["nus"]
        return pb.ProductMatern(shape, lengthscales=np.asarray(ls) if np.ndim(ls) else ls,
                                nus=np.asarray(nus) if np.ndim(nus) else nus)
    raise ValueError(kind)


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
