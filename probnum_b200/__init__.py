"""probnum_b200 -- ProbNum's covariance-function hot path on NVIDIA B200 (sm_100a).

Drop-in for ``probnum.randprocs.covfuncs.{ExpQuad, Matern, RatQuad}``:
``k(x0, x1)``, ``k.matrix(x0, x1)`` and ``k.linop(x0, x1) @ v`` run as hand-written CUDA
kernels behind the C ABI of ``include/pncov.h``.  See DESIGN.md.

* :mod:`probnum_b200.covfuncs` -- mirror of the reference classes (same API)
* :mod:`probnum_b200.linops`   -- the slice of ``probnum.linops`` the path touches
* :mod:`probnum_b200.patch`    -- ``enable()`` installs the device path into an
  importable ``probnum`` in place (same classes, so ``isinstance`` checks in
  ``probnum.quad`` keep working)
* :mod:`probnum_b200.sharded`  -- row-sharded operator over ``torch.distributed``
"""

from . import covfuncs, linops
from .covfuncs import (CovarianceFunction, CovarianceLinearOperator, DeviceMatrix, ExpQuad, Linear, Matern, Polynomial,
                       ProductCovarianceFunction, ProductMatern, RatQuad, ScaledCovarianceFunction,
                       SumCovarianceFunction, WhiteNoise)

__all__ = [
    "covfuncs",
    "linops",
    "CovarianceFunction",
    "CovarianceLinearOperator",
    "DeviceMatrix",
    "ExpQuad",
    "Matern",
    "RatQuad",
    "Linear",
    "Polynomial",
    "WhiteNoise",
    "ProductMatern",
    "ScaledCovarianceFunction",
    "SumCovarianceFunction",
    "ProductCovarianceFunction",
]
__version__ = "0.1.0"
