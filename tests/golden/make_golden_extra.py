"""Golden fixtures for the remaining covariance functions and covariance arithmetic, from the
REAL reference (build container only; see make_golden.py for how the reference is imported).

    python tests/golden/make_golden_extra.py

Writes ``tests/golden/covfuncs_extra_golden.{npz,json}``.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_kernel as build_isotropic  # noqa: E402
from make_golden import import_reference  # noqa: E402


def build(pn, spec):
    cf = pn.randprocs.covfuncs
    kind = spec["kind"]
    if kind in ("expquad", "matern", "ratquad"):
        return build_isotropic(pn, spec)
    if kind == "scaled":
        return spec["scalar"] * build(pn, spec["child"])
    if kind == "sum":
        parts = [build(pn, c) for c in spec["children"]]
        out = parts[0]
        for p in parts[1:]:
            out = out + p
        return out
    if kind == "product":
        parts = [build(pn, c) for c in spec["children"]]
        out = parts[0]
        for p in parts[1:]:
            out = out * p
        return out
    shape = tuple(spec["input_shape"])
    if kind == "linear":
        return cf.Linear(shape, constant=spec["constant"])
    if kind == "polynomial":
        return cf.Polynomial(shape, constant=spec["constant"], exponent=spec["exponent"])
    if kind == "whitenoise":
        return cf.WhiteNoise(shape, sigma_sq=spec["sigma_sq"])
    if kind == "productmatern":
        ls, nus = spec["lengthscales"], spec["nus"]
        return cf.ProductMatern(shape, lengthscales=np.asarray(ls) if np.ndim(ls) else ls,
                                nus=np.asarray(nus) if np.ndim(nus) else nus)
    raise ValueError(kind)


def specs(shape):
    d = list(shape)
    n = max(1, int(np.prod(shape)))
    eq = dict(kind="expquad", input_shape=d, lengthscales=1.3)
    m52 = dict(kind="matern", input_shape=d, lengthscales=0.8, nu=2.5)
    rq = dict(kind="ratquad", input_shape=d, lengthscales=0.7, alpha=2.0)
    lin = dict(kind="linear", input_shape=d, constant=0.5)
    wn = dict(kind="whitenoise", input_shape=d, sigma_sq=0.3)
    out = [
        dict(kind="linear", input_shape=d, constant=0.0),
        lin,
        dict(kind="polynomial", input_shape=d, constant=1.0, exponent=3),
        dict(kind="polynomial", input_shape=d, constant=0.25, exponent=2.0),
        dict(kind="polynomial", input_shape=d, constant=2.0, exponent=1.0),
        dict(kind="polynomial", input_shape=d, constant=-0.5, exponent=5),
        wn,
        dict(kind="productmatern", input_shape=d, lengthscales=0.9, nus=1.5),
        dict(kind="scaled", scalar=2.5, child=eq),
        dict(kind="sum", children=[eq, m52]),
        dict(kind="sum", children=[dict(kind="scaled", scalar=1.7, child=m52), wn]),
        dict(kind="product", children=[eq, lin]),
        dict(kind="sum", children=[dict(kind="scaled", scalar=0.4, child=rq),
                                   dict(kind="product", children=[m52, dict(kind="polynomial", input_shape=d,
                                                                            constant=1.0, exponent=2)]),
                                   wn]),
        dict(kind="scaled", scalar=3.0, child=dict(kind="sum", children=[eq, lin])),
    ]
    if shape != ():
        out.append(dict(kind="productmatern", input_shape=d,
                        lengthscales=[0.5 + 0.3 * i for i in range(n)],
                        nus=[[0.5, 1.5, 2.5, 3.5][i % 4] for i in range(n)]))
    return out


GRAM_SHAPES = [(1, None), (2, 3), (10, None), (33, 17), (70, None), (100, 10)]
CALL_SHAPES = [[(), None], [(), ()], [(10,), None], [(10,), (10,)], [(30, 1), (9,)], [(10, 1), (1, 2)]]


def main():
    pn = import_reference()
    rng = np.random.default_rng(20261020)
    arrays, cases = {}, []

    def put(name, value):
        arrays[name] = np.asarray(value)
        return name

    # docstring known answers: _linear.py:36-41, _polynomial.py:40-45
    kat = []
    for spec, xs, doc, cite in [
        (dict(kind="linear", input_shape=[2], constant=0.0), np.array([[1, 2], [2, 3]]),
         [[5.0, 8.0], [8.0, 13.0]], "randprocs/covfuncs/_linear.py:36-41"),
        (dict(kind="polynomial", input_shape=[2], constant=1.0, exponent=3), np.array([[1, -1], [-1, 0]]),
         [[27.0, 0.0], [0.0, 8.0]], "randprocs/covfuncs/_polynomial.py:40-45"),
    ]:
        got = build(pn, spec).matrix(xs)
        np.testing.assert_array_equal(got, np.asarray(doc))
        kat.append(dict(spec=spec, x=put(f"kat{len(kat)}_x", xs), doc=doc, cite=cite))

    idx = 0
    for shape in ((), (3,), (20,)):
        for spec in specs(shape):
            k = build(pn, spec)
            for n0, n1 in GRAM_SHAPES:
                x0 = rng.normal(size=(n0,) + shape)
                x1 = None if n1 is None else rng.normal(size=(n1,) + shape)
                if n0 >= 10 and n1 is None:
                    x0[3] = x0[7]  # duplicated point (WhiteNoise off the diagonal)
                if n1 is not None and n1 >= 10:
                    x1[2] = x0[1]  # shared point between the two sets
                tag = f"x{idx:04d}"
                cases.append(dict(type="gram", spec=spec, x0=put(tag + "_x0", x0),
                                  x1=None if x1 is None else put(tag + "_x1", x1),
                                  K=put(tag + "_K", k.matrix(x0, x1))))
                idx += 1
            if shape in ((), (3,)):
                for b0, b1 in CALL_SHAPES:
                    x0 = rng.normal(size=tuple(b0) + shape)
                    x1 = None if b1 is None else rng.normal(size=tuple(b1) + shape)
                    tag = f"y{idx:04d}"
                    cases.append(dict(type="call", spec=spec, x0=put(tag + "_x0", x0),
                                      x1=None if x1 is None else put(tag + "_x1", x1),
                                      out=put(tag + "_out", k(x0, x1))))
                    idx += 1
    # float32 points: dtype behaviour of the arithmetic
    for spec in specs((3,))[:3] + specs((3,))[8:11]:
        k = build(pn, spec)
        x0 = rng.normal(size=(40, 3)).astype(np.float32)
        tag = f"z{idx:04d}"
        cases.append(dict(type="gram", spec=spec, x0=put(tag + "_x0", x0), x1=None, K=put(tag + "_K", k.matrix(x0))))
        idx += 1

    np.savez_compressed(os.path.join(HERE, "covfuncs_extra_golden.npz"), **arrays)
    with open(os.path.join(HERE, "covfuncs_extra_golden.json"), "w") as fh:
        json.dump(dict(kat=kat, cases=cases, numpy=np.__version__,
                       generator="tests/golden/make_golden_extra.py"), fh, indent=0)
    print(len(cases), "cases", os.path.getsize(os.path.join(HERE, "covfuncs_extra_golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
