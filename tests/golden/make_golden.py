"""Generate golden fixtures for the covariance hot path from the REAL reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports ProbNum's own NumPy implementation from a scratch copy of
``/root/reference/src/probnum`` (plus the ``_version.py`` stub that setuptools_scm
would generate and four NumPy-2 aliases), evaluates
``CovarianceFunction.__call__ / matrix / linop(...) @ v`` on seeded inputs and
stores inputs + outputs in ``tests/golden/covfuncs_golden.npz``.  The fixture
sizes follow the reference's own covfuncs test-suite
(tests/test_randprocs/test_covfuncs/conftest.py:12-137, test_call.py:12-39).
"""

from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/probnum"


def import_reference():
    scratch = tempfile.mkdtemp(prefix="probnum_ref_")
    shutil.copytree(REF_SRC, os.path.join(scratch, "probnum"))
    with open(os.path.join(scratch, "probnum", "_version.py"), "w") as fh:
        fh.write('version = "0.0.0+ref"\n__version__ = version\n')
    np.float_ = np.float64
    np.Inf = np.inf
    np.find_common_type = lambda a, s: np.result_type(*a, *s)
    np.AxisError = np.exceptions.AxisError
    sys.path.insert(0, scratch)
    warnings.simplefilter("ignore")
    import probnum  # noqa: PLC0415

    return probnum


def build_kernel(pn, spec):
    cf = pn.randprocs.covfuncs
    shape = tuple(spec["input_shape"])
    if spec["kind"] == "expquad":
        return cf.ExpQuad(shape, lengthscales=np.asarray(spec["lengthscales"]))
    if spec["kind"] == "matern":
        return cf.Matern(
            shape, nu=spec["nu"], lengthscales=np.asarray(spec["lengthscales"])
        )
    if spec["kind"] == "ratquad":
        return cf.RatQuad(shape, lengthscale=spec["lengthscales"], alpha=spec["alpha"])
    raise ValueError(spec["kind"])


def kernel_specs(rng):
    """Kernel zoo: the reference fixtures (conftest.py:38-48), the asv defaults
    (benchmarks/covfuncs.py:31-40) and ARD variants as in test_matern.py:34."""
    specs = []
    for d in (None, 1, 3, 10):
        shape = () if d is None else (d,)
        specs += [
            dict(kind="expquad", input_shape=shape, lengthscales=1.5),
            dict(kind="expquad", input_shape=shape, lengthscales=1.0),
            dict(kind="ratquad", input_shape=shape, lengthscales=0.5, alpha=2.0),
            dict(kind="ratquad", input_shape=shape, lengthscales=1.0, alpha=1.0),
            dict(kind="ratquad", input_shape=shape, lengthscales=1.3, alpha=0.7),
            dict(kind="matern", input_shape=shape, lengthscales=0.5, nu=0.5),
            dict(kind="matern", input_shape=shape, lengthscales=0.5, nu=1.5),
            dict(kind="matern", input_shape=shape, lengthscales=1.5, nu=2.5),
            dict(kind="matern", input_shape=shape, lengthscales=1.2, nu=3.5),
        ]
        if d is not None and d > 1:
            ard = (rng.gamma(1.0, size=shape) + 0.5).tolist()
            specs += [
                dict(kind="expquad", input_shape=shape, lengthscales=ard),
                dict(kind="matern", input_shape=shape, lengthscales=ard, nu=0.5),
                dict(kind="matern", input_shape=shape, lengthscales=ard, nu=2.5),
            ]
    return specs


# (x0 batch shape, x1 batch shape or None) for __call__: test_call.py:12-39
CALL_SHAPES = [
    [(), None],
    [(), ()],
    [(), (3,)],
    [(1,), (2,)],
    [(10,), None],
    [(10,), (10,)],
    [(30, 1), (9,)],
    [(10, 1), (1, 2)],
    [(2, 4, 1, 1, 3), (1, 4, 5, 1, 1)],
]

# (N0, N1 or None) for matrix / linop
GRAM_SHAPES = [(1, None), (2, 3), (10, None), (10, 1), (33, 17), (64, None), (100, 10)]


def main():
    pn = import_reference()
    rng = np.random.default_rng(20261019)
    arrays = {}
    cases = []

    def put(name, value):
        arrays[name] = np.asarray(value)
        return name

    # ---- known-answer values printed in the reference docstrings -------------
    kat = []
    xs = np.linspace(0, 1, 3)
    kat.append(
        dict(
            spec=dict(kind="expquad", input_shape=[], lengthscales=0.1),
            x=put("kat_x_expquad", xs),
            doc=[
                [1.00000000e00, 3.72665317e-06, 1.92874985e-22],
                [3.72665317e-06, 1.00000000e00, 3.72665317e-06],
                [1.92874985e-22, 3.72665317e-06, 1.00000000e00],
            ],
            cite="randprocs/covfuncs/_exponentiated_quadratic.py:47-52",
        )
    )
    kat.append(
        dict(
            spec=dict(kind="matern", input_shape=[], lengthscales=0.1, nu=2.5),
            x=put("kat_x_matern", xs),
            doc=[
                [1.00000000e00, 7.50933789e-04, 3.69569622e-08],
                [7.50933789e-04, 1.00000000e00, 7.50933789e-04],
                [3.69569622e-08, 7.50933789e-04, 1.00000000e00],
            ],
            cite="randprocs/covfuncs/_matern.py:91-96",
        )
    )
    kat.append(
        dict(
            spec=dict(kind="ratquad", input_shape=[1], lengthscales=0.1, alpha=3.0),
            x=put("kat_x_ratquad", xs[:, None]),
            doc=[
                [1.00000000e00, 7.25051190e-03, 1.81357765e-04],
                [7.25051190e-03, 1.00000000e00, 7.25051190e-03],
                [1.81357765e-04, 7.25051190e-03, 1.00000000e00],
            ],
            cite="randprocs/covfuncs/_rational_quadratic.py:51-56",
        )
    )
    kat.append(
        dict(
            spec=dict(kind="expquad", input_shape=[], lengthscales=1.0),
            x=put("kat_x_gp", np.linspace(-1, 1, 5)),
            doc=[
                [1.0, 0.8824969, 0.60653066, 0.32465247, 0.13533528],
                [0.8824969, 1.0, 0.8824969, 0.60653066, 0.32465247],
                [0.60653066, 0.8824969, 1.0, 0.8824969, 0.60653066],
                [0.32465247, 0.60653066, 0.8824969, 1.0, 0.8824969],
                [0.13533528, 0.32465247, 0.60653066, 0.8824969, 1.0],
            ],
            cite="randprocs/_gaussian_process.py:51-56",
        )
    )
    # the reference itself must reproduce its docstrings (sanity of this script)
    for entry in kat:
        k = build_kernel(pn, entry["spec"])
        got = k.matrix(arrays[entry["x"]])
        np.testing.assert_allclose(got, np.asarray(entry["doc"]), rtol=5e-8)  # printed to 8-9 digits
        entry["ref"] = put(entry["x"].replace("_x_", "_ref_"), got)

    # ---- seeded parity cases ------------------------------------------------------
    idx = 0
    for spec in kernel_specs(rng):
        k = build_kernel(pn, spec)
        shape = tuple(spec["input_shape"])
        for dtype in ("float64", "float32"):
            # float32 goldens only for a subset (dtype behaviour, SURVEY 7 last bullet)
            if dtype == "float32" and not (shape in ((3,), ()) and np.ndim(spec["lengthscales"]) == 0):
                continue
            for n0, n1 in GRAM_SHAPES:
                loc = 0.0 if idx % 3 else 50.0  # every third case is an un-centred cloud
                x0 = (rng.normal(size=(n0,) + shape) + loc).astype(dtype)
                x1 = None if n1 is None else (rng.normal(size=(n1,) + shape) + loc).astype(dtype)
                ncols = n0 if n1 is None else n1
                if n0 >= 10 and n1 is None:
                    x0[3] = x0[7]  # duplicated point: exact zero distance off the diagonal
                v = rng.normal(size=(ncols, 3)).astype(dtype)
                tag = f"g{idx:04d}"
                K = k.matrix(x0, x1)
                Kv = k.linop(x0, x1) @ v
                cases.append(
                    dict(
                        type="gram", spec=spec, dtype=dtype,
                        x0=put(tag + "_x0", x0),
                        x1=None if x1 is None else put(tag + "_x1", x1),
                        v=put(tag + "_v", v),
                        K=put(tag + "_K", K), Kv=put(tag + "_Kv", Kv),
                    )
                )
                idx += 1
        if shape in ((), (3,)):
            for b0, b1 in CALL_SHAPES:
                x0 = rng.normal(size=tuple(b0) + shape)
                x1 = None if b1 is None else rng.normal(size=tuple(b1) + shape)
                tag = f"c{idx:04d}"
                cases.append(
                    dict(
                        type="call", spec=spec, dtype="float64",
                        x0=put(tag + "_x0", x0),
                        x1=None if x1 is None else put(tag + "_x1", x1),
                        out=put(tag + "_out", k(x0, x1)),
                    )
                )
                idx += 1

    # ---- a larger far-field / underflow case (exp underflow, subnormals) ---------
    spec = dict(kind="expquad", input_shape=[2], lengthscales=0.05)
    k = build_kernel(pn, spec)
    x0 = rng.uniform(-1.5, 1.5, size=(40, 2))
    cases.append(
        dict(type="gram", spec=spec, dtype="float64", x0=put("uf_x0", x0), x1=None,
             v=put("uf_v", rng.normal(size=(40, 1))), K=put("uf_K", k.matrix(x0)),
             Kv=put("uf_Kv", k.linop(x0) @ arrays["uf_v"]))
    )

    np.savez_compressed(os.path.join(HERE, "covfuncs_golden.npz"), **arrays)
    with open(os.path.join(HERE, "covfuncs_golden.json"), "w") as fh:
        json.dump(dict(kat=kat, cases=cases, numpy=np.__version__,
                       generator="tests/golden/make_golden.py"), fh, indent=0)
    size = os.path.getsize(os.path.join(HERE, "covfuncs_golden.npz"))
    print(f"{len(cases)} cases, {len(kat)} KATs, {size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
