"""Host logic of the one-launch covariance arithmetic (no GPU needed): `covfuncs._compile_program` turns an
expression tree into the step list of `pncov_set_program`, mirroring the operand-by-operand evaluation order of
`_ArithmeticNode._combine_into` (reference: _arithmetic_fallbacks.py:55-56, :101-104, :163-166)."""

import numpy as np

import probnum_b200 as pb
from probnum_b200 import _lib, covfuncs

L, ML, SV, SET, LIN, MUL, SA = (_lib.F_LOAD, _lib.F_MULLEAF, _lib.F_SCALEVAL, _lib.F_SET, _lib.F_LIN, _lib.F_MUL,
                                _lib.F_SCALEACC)


def ops(prog):
    return [st[0] for st in prog[1]]


def test_signal_plus_noise():
    k = 2.0 * pb.Matern((3,), nu=2.5) + pb.WhiteNoise((3,), sigma_sq=0.5)
    leaves, steps, params = covfuncs._compile_program(k)
    assert [l[0] for l in leaves] == [_lib.MATERN, _lib.WHITENOISE]
    assert leaves[1][3] == 0.5 and leaves[0][2] == 2.5
    assert ops((leaves, steps, params)) == [L, SET, SA, L, LIN]
    assert steps[2][2] == 2.0 and steps[4][2:] == (1.0, 1.0)
    assert params.d == 3


def test_scaled_summands_fold_their_scalar_into_the_combination():
    k = pb.ExpQuad((2,)) + 0.1 * pb.RatQuad((2,), alpha=1.5) + pb.Matern((2,), nu=0.5) * pb.ExpQuad((2,), lengthscales=0.3)
    leaves, steps, _ = covfuncs._compile_program(k)
    assert ops((leaves, steps, None)) == [L, SET, L, LIN, L, ML, LIN]
    assert steps[3][2:] == (1.0, 0.1) and steps[6][2:] == (1.0, 1.0)
    assert len(leaves) == 4  # ExpQuad(l=1), RatQuad, Matern, ExpQuad(l=0.3)


def test_repeated_leaves_share_one_description():
    e = pb.ExpQuad((4,), lengthscales=0.7)
    leaves, steps, _ = covfuncs._compile_program(e * e + e)
    assert len(leaves) == 1
    assert ops((leaves, steps, None)) == [L, SET, L, MUL, L, LIN]


def test_products_and_outer_scale():
    k = 1.5 * (pb.ExpQuad((1,)) * pb.Matern((1,), nu=1.5))
    _, steps, _ = covfuncs._compile_program(k)
    assert ops((None, steps, None)) == [L, SET, L, MUL, SA]
    k2 = (3.0 * pb.ExpQuad((1,))) * pb.Matern((1,), nu=1.5)
    _, steps2, _ = covfuncs._compile_program(k2)
    assert ops((None, steps2, None)) == [L, SET, SA, L, MUL]


def test_expressions_outside_the_program_form():
    d = 3
    ard = pb.ExpQuad((d,), lengthscales=np.array([1.0, 2.0, 3.0]))
    assert covfuncs._compile_program(ard + pb.Matern((d,), nu=1.5)) is None                    # per-dimension lengthscales
    assert covfuncs._compile_program(pb.Linear((d,), constant=1.0) + pb.ExpQuad((d,))) is None  # not isotropic
    assert covfuncs._compile_program(pb.ExpQuad((20,)) + pb.Matern((20,), nu=1.5)) is None      # d > 16
    assert covfuncs._compile_program(pb.Matern((d,), nu=0.7) + pb.ExpQuad((d,))) is None        # Bessel branch
    nested = (pb.ExpQuad((d,)) + pb.Matern((d,), nu=0.5)) * (pb.RatQuad((d,)) + pb.Matern((d,), nu=1.5))
    assert covfuncs._compile_program(nested) is None                                             # product of sums
    many = pb.ExpQuad((d,), lengthscales=0.5)
    for ls in (0.7, 0.9, 1.1, 1.3, 1.5, 1.7):
        many = many + pb.Matern((d,), nu=1.5, lengthscales=ls)
    assert covfuncs._compile_program(many) is None                                               # seven leaf evaluations
    six = pb.ExpQuad((d,), lengthscales=0.5)
    for ls in (0.7, 0.9, 1.1, 1.3, 1.5):
        six = six + pb.Matern((d,), nu=1.5, lengthscales=ls)
    assert covfuncs._compile_program(six) is not None
