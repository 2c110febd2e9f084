"""`pncov_sharded_*`: one operator over several GPUs of ONE process through the C ABI alone (SURVEY.md 8b:
`device_ids[], n_dev`).  On a one-GPU box the device id repeats, which exercises the same decomposition."""

import numpy as np
import pytest

import probnum_b200 as pb
from oracle import covfuncs_oracle as orc
from probnum_b200 import _lib

pytestmark = pytest.mark.gpu


def device_lists():
    n = _lib.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_gram_and_products_match_one_device(devices):
    rng = np.random.default_rng(4)
    n, d = 4100, 3
    x = rng.normal(size=(n, d))
    x1 = rng.normal(size=(1500, d))
    k = pb.Matern((d,), nu=2.5, lengthscales=0.8)
    spec = orc.KernelSpec("matern", (d,), 0.8, nu=2.5)
    with _lib.Operator(k._kernel_params(), x, None) as one, _lib.ShardedOperator(k._kernel_params(), x, None, devices=devices) as sh:
        K1 = one.gram(np.empty((n, n)))
        K = sh.gram()
        assert np.array_equal(K, K1)  # direct differences: the same operations per entry on every device
        v = rng.normal(size=(n, 1))
        y1 = one.matmul(v)
        assert one.matmul_path() == "direct-differences-symmetric"
        y = sh.matmul(v)
        assert np.array_equal(y, y1)  # integer planes added on the host: the bits of the one-device product
        assert np.array_equal(sh.matmul(v), y)
        V = rng.normal(size=(n, 3))
        Y = sh.matmul(V)
        assert np.all(np.abs(Y - orc.matvec(spec, x, None, V)) <= 1e-12 * orc.abs_matvec(spec, x, None, V))
    with _lib.ShardedOperator(k._kernel_params(), x, x1, devices=devices) as sh:
        np.testing.assert_allclose(sh.gram(), orc.gram(spec, x, x1), rtol=1e-12, atol=1e-12)
        w = rng.normal(size=(1500, 2))
        assert np.all(np.abs(sh.matmul(w) - orc.matvec(spec, x, x1, w)) <= 1e-12 * orc.abs_matvec(spec, x, x1, w))


@pytest.mark.parametrize("devices", device_lists()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_expquad_identity_kernel_and_edges(devices):
    rng = np.random.default_rng(8)
    n, d = 5000, 8
    x = rng.normal(size=(n, d))
    kp = pb.ExpQuad((d,))._kernel_params()
    spec = orc.KernelSpec("expquad", (d,), 1.0)
    v = rng.normal(size=(n, 1))
    with _lib.Operator(kp, x, None) as one, _lib.ShardedOperator(kp, x, None, devices=devices) as sh:
        assert np.array_equal(sh.matmul(v), one.matmul(v))  # DMMA identity kernel, symmetric halving over the devices
        assert one.matmul_path() == "norm-identity-dmma-symmetric"
    # fewer rows than devices, and a cloud too wide for the identity (direct differences take over on every device)
    xs = rng.normal(size=(2, d))
    with _lib.ShardedOperator(kp, xs, None, devices=devices) as sh:
        np.testing.assert_allclose(sh.gram(), orc.gram(spec, xs), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(sh.matmul(np.ones((2, 1))), orc.gram(spec, xs) @ np.ones((2, 1)), rtol=1e-12)
    wide = _lib.KernelParams(_lib.EXPQUAD, d, np.array([0.01]))
    specw = orc.KernelSpec("expquad", (d,), 0.01)
    with _lib.ShardedOperator(wide, x[:2500], None, devices=devices) as sh:
        y = sh.matmul(v[:2500])
        assert np.all(np.abs(y - orc.matvec(specw, x[:2500], None, v[:2500])) <= 1e-12 * orc.abs_matvec(specw, x[:2500], None, v[:2500]))
    with pytest.raises(ValueError):
        _lib.ShardedOperator(kp, x, None, devices=[])
