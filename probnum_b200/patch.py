"""Install the B200 path into an importable ``probnum`` in place (the drop-in).

    import probnum
    import probnum_b200.patch as pb_patch
    pb_patch.enable()          # ExpQuad / Matern(nu = p + 1/2) / RatQuad now run on the GPU
    ...
    pb_patch.disable()         # restore the reference's NumPy methods

Only the two override points named in SURVEY.md 8b are replaced, on the reference's own
classes, so ``isinstance(kernel, ExpQuad)`` checks elsewhere in ProbNum (e.g. the kernel
embeddings of ``probnum.quad``, quad/kernel_embeddings/_kernel_embedding.py:95-138) keep
working and every caller -- ``CovarianceFunction.matrix / linop``, ``GaussianProcess.__call__``,
``problinsolve`` on ``k.linop(X) + sigma^2 I``, the Bayesian-quadrature belief update -- picks
the device path up without modification:

* ``_evaluate_matrix(x0, x1)``  (randprocs/covfuncs/_covariance_function.py:553-582)
* ``_evaluate_linop(x0, x1)``   (randprocs/covfuncs/_covariance_function.py:584-604)

plus ``_evaluate`` for the Gram-shaped broadcast ``(N0, 1, d) x (1, N1, d)``: that is what the
covariance arithmetic of the reference calls on each operand
(``ScaledCovarianceFunction._evaluate`` / ``SumCovarianceFunction._evaluate`` /
``ProductCovarianceFunction._evaluate``, randprocs/covfuncs/_arithmetic_fallbacks.py:55-56,
:101-104, :163-166), so ``(k1 + 0.1 * k2).matrix(X)`` builds every operand on the device and
only the final combination stays in NumPy (first step).

Kernels outside the accelerated scope (other classes, Matern with non-half-integer nu,
multi-dimensional input shapes) keep the reference's own methods untouched: that is the
reference's behaviour for them, not a fallback of ours.
"""

from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import covfuncs as _mirror

_ORIGINALS = {}


def numpy2_compat() -> None:
    """Aliases removed in NumPy 2 that the reference still uses.  Needed to
    import the reference at all under NumPy >= 2; harmless otherwise."""
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if not hasattr(np, "find_common_type"):
        np.find_common_type = lambda array_types, scalar_types: np.result_type(*array_types, *scalar_types)
    if not hasattr(np, "AxisError"):
        np.AxisError = np.exceptions.AxisError


def _mirror_of(k) -> Optional[_mirror.CovarianceFunction]:
    """The probnum_b200 twin of a reference covariance function, or None if it is outside
    the accelerated scope."""
    import probnum  # noqa: PLC0415

    cf = probnum.randprocs.covfuncs
    try:
        if len(k.input_shape) > 1:
            return None
    except ValueError:
        return None
    if type(k) is cf.ExpQuad:
        return _mirror.ExpQuad(k.input_shape, lengthscales=k.lengthscales)
    if type(k) is cf.Matern:
        if not k.is_half_integer or k.p > _mirror._lib.MAX_MATERN_P:
            return None
        return _mirror.Matern(k.input_shape, nu=k.nu, lengthscales=k.lengthscales)
    if type(k) is cf.RatQuad:
        if k.lengthscale == 0:
            return None
        return _mirror.RatQuad(k.input_shape, lengthscale=k.lengthscale, alpha=k.alpha)
    # the remaining families and covariance arithmetic (dense Gram only)
    if type(k) is cf.Linear:
        return _mirror.Linear(k.input_shape, constant=k.constant)
    if type(k) is cf.Polynomial:
        q = float(k.exponent)
        if not (0 <= q <= 64 and q == int(q)):
            return None
        return _mirror.Polynomial(k.input_shape, constant=k.constant, exponent=k.exponent)
    if type(k) is cf.WhiteNoise:
        return _mirror.WhiteNoise(k.input_shape, sigma_sq=k.sigma_sq)
    if type(k) is cf.ProductMatern:
        nus = np.atleast_1d(np.asarray(k.nus, dtype=float))
        if not all((nu - 0.5) == int(nu - 0.5) and nu - 0.5 <= _mirror._lib.MAX_MATERN_P for nu in nus):
            return None
        if k.input_shape == ():
            return _mirror.ProductMatern((), lengthscales=float(np.ravel(k.lengthscales)[0]), nus=float(nus[0]))
        return _mirror.ProductMatern(k.input_shape, lengthscales=np.asarray(k.lengthscales, dtype=float), nus=nus)
    from probnum.randprocs.covfuncs import _arithmetic_fallbacks as af  # noqa: PLC0415

    if type(k) is af.ScaledCovarianceFunction:
        inner = _mirror_of(k._covfunc)
        return None if inner is None else _mirror.ScaledCovarianceFunction(inner, k._scalar)
    if type(k) in (af.SumCovarianceFunction, af.ProductCovarianceFunction):
        parts = [_mirror_of(c) for c in (k._summands if type(k) is af.SumCovarianceFunction else k._factors)]
        if any(c is None for c in parts):
            return None
        return (_mirror.SumCovarianceFunction if type(k) is af.SumCovarianceFunction
                else _mirror.ProductCovarianceFunction)(*parts)
    return None


def _make_linop_class():
    import probnum  # noqa: PLC0415

    class GpuCovarianceLinearOperator(probnum.linops.LinearOperator):
        """``probnum.linops.LinearOperator`` backed by the resident device operator
        (replaces CovarianceLinearOperator, _covariance_linear_operator.py:22-90)."""

        def __init__(self, inner: _mirror.CovarianceLinearOperator):
            self._inner = inner
            super().__init__(inner.shape, inner.dtype)

        def _matmul(self, x: np.ndarray) -> np.ndarray:
            return self._inner._matmul(np.asarray(x))

        def _todense(self) -> np.ndarray:
            return self._inner._todense()

        def _transpose(self):
            return GpuCovarianceLinearOperator(self._inner._transpose())

        def _diagonal(self) -> np.ndarray:
            return self._inner._diagonal()

        def _trace(self):
            return np.sum(self._inner._diagonal())

    return GpuCovarianceLinearOperator


_SHARDED: Optional[bool] = None
_LOCAL_FACTORY = None


def _use_sharded() -> bool:
    """One process per GPU under torchrun: operators shard over the ranks of the default group."""
    if _SHARDED is not None:
        return _SHARDED
    try:
        import torch.distributed as dist  # noqa: PLC0415
    except ImportError:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _make_sharded_linop_class():
    import probnum  # noqa: PLC0415

    from .sharded import ShardedCovarianceOperator  # noqa: PLC0415

    class GpuShardedCovarianceLinearOperator(probnum.linops.LinearOperator):
        """``k.linop(X)`` with the product spread over all ranks (probnum_b200.sharded): every rank
        runs the same (sequential) solver on the same iterate and receives the full product after one
        collective (BASELINE.json configs[3]: problinsolve via the GPU operator + NCCL)."""

        def __init__(self, twin, x0, x1):
            self._twin, self._x0, self._x1 = twin, x0, x1
            n0 = x0.shape[0]
            n1 = n0 if x1 is None else x1.shape[0]
            dtype = np.promote_types(x0.dtype, x1.dtype) if x1 is not None else x0.dtype
            if dtype not in (np.float32, np.float64):
                dtype = np.dtype(np.float64)
            self._sop = None
            super().__init__((n0, n1), dtype)

        def _sharded(self):
            if self._sop is None:
                self._sop = ShardedCovarianceOperator(self._twin, self._x0, self._x1, local_factory=_LOCAL_FACTORY)
            return self._sop

        def _matmul(self, x: np.ndarray) -> np.ndarray:
            x = np.asarray(x)
            lead, (n, k) = x.shape[:-2], x.shape[-2:]
            if lead:  # stacked operands: fold the stack into columns
                B = int(np.prod(lead, dtype=np.int64))
                v = np.moveaxis(x.reshape(B, n, k), 0, 1).reshape(n, B * k)
                y = self._sharded().matmul(v)
                y = np.moveaxis(y.reshape(self.shape[0], B, k), 1, 0).reshape(lead + (self.shape[0], k))
            else:
                y = self._sharded().matmul(x)
            return y.astype(np.result_type(self.dtype, x.dtype), copy=False)

        def _transpose(self):
            if self._x1 is None:
                return self  # isotropic kernels: K(X, X) is symmetric
            return GpuShardedCovarianceLinearOperator(self._twin, self._x1, self._x0)

        def _todense(self) -> np.ndarray:
            return self._twin._evaluate_matrix(self._x0, self._x1)

        def _diagonal(self) -> np.ndarray:
            n = min(self.shape)
            if self._x1 is None:
                return np.ones(n, dtype=self._twin._result_dtype(np.dtype(self.dtype)))
            return np.array(self._twin(self._x0[:n], self._x1[:n]))

        def _trace(self):
            return np.sum(self._diagonal())

    return GpuShardedCovarianceLinearOperator


def enable(sharded: Optional[bool] = None, _local_factory=None) -> None:
    """Patch ExpQuad, Matern and RatQuad (Gram, operator, Gram-shaped calls) and the Gram hook of
    Linear, Polynomial, WhiteNoise, ProductMatern and the arithmetic nodes of the importable
    ``probnum``.

    ``sharded``: ``None`` (default) shards ``k.linop(X)`` over the ranks of ``torch.distributed``
    whenever a process group with more than one rank is initialised (one process per GPU under
    torchrun), ``True`` / ``False`` force the choice.  ``_local_factory`` replaces the per-rank device
    operator (CPU tests of the host logic only)."""
    global _SHARDED, _LOCAL_FACTORY
    numpy2_compat()
    import probnum  # noqa: PLC0415

    if enabled():
        return
    _SHARDED, _LOCAL_FACTORY = sharded, _local_factory
    cf = probnum.randprocs.covfuncs
    linop_cls = _make_linop_class()
    sharded_cls = _make_sharded_linop_class()

    def make_matrix(original):
        def _evaluate_matrix(self, x0, x1):
            twin = _mirror_of(self)
            if twin is None:
                return original(self, x0, x1)
            return twin._evaluate_matrix(x0, x1)

        return _evaluate_matrix

    def make_linop(original):
        def _evaluate_linop(self, x0, x1):
            twin = _mirror_of(self)
            if twin is None:
                return original(self, x0, x1)
            if _use_sharded():
                return sharded_cls(twin, x0, x1)
            return linop_cls(twin._evaluate_linop(x0, x1))

        return _evaluate_linop

    def make_evaluate(original):
        def _evaluate(self, x0, x1=None):
            # only the Gram-shaped broadcast with enough pairs to be worth a launch; every other
            # pattern is O(N d) work and stays with the reference
            if (x1 is not None and x0.ndim == self.input_ndim + 2 and x1.ndim == self.input_ndim + 2
                    and x0.shape[1] == 1 and x1.shape[0] == 1 and x0.shape[0] * x1.shape[1] >= 4096):
                twin = _mirror_of(self)
                if twin is not None:
                    return twin._evaluate(x0, x1)
            return original(self, x0, x1)

        return _evaluate

    for cls in (cf.ExpQuad, cf.Matern, cf.RatQuad):
        _ORIGINALS[cls] = (cls._evaluate_matrix, cls._evaluate_linop, cls.__dict__["_evaluate"])
        cls._evaluate_matrix = make_matrix(cls._evaluate_matrix)
        cls._evaluate_linop = make_linop(cls._evaluate_linop)
        cls._evaluate = make_evaluate(cls.__dict__["_evaluate"])
    # dense-only families and arithmetic nodes: only the Gram hook.  Their operators keep the
    # reference's dense-cached CovarianceLinearOperator, whose matrix now comes from the device.
    from probnum.randprocs.covfuncs import _arithmetic_fallbacks as af  # noqa: PLC0415

    for cls in (cf.Linear, cf.Polynomial, cf.WhiteNoise, cf.ProductMatern, af.ScaledCovarianceFunction,
                af.SumCovarianceFunction, af.ProductCovarianceFunction):
        _ORIGINALS[("matrix", cls)] = cls._evaluate_matrix
        cls._evaluate_matrix = make_matrix(cls._evaluate_matrix)


# ---------------------------------------------------------------------------------------------
# Cholesky + solves of the Bayesian-quadrature belief update on the device
# ---------------------------------------------------------------------------------------------
class DeviceCholesky:
    """Stand-in for the ``(c, lower)`` pair of ``scipy.linalg.cho_factor``: the factor of
    ``gram + jitter I`` stays in HBM (the hand-written blocked FP64 Cholesky of libpncov,
    csrc/pncov_chol.cu, DMMA trailing updates); only right-hand sides and solutions cross PCIe.
    (The cuSOLVER baseline the survey names for this row lives in tools/cusolver_baseline.py.)"""

    def __init__(self, factor):
        self.factor = factor  # _lib.Cholesky

    # scipy returns a 2-tuple; the reference only ever passes it back to gram_cho_solve
    def __iter__(self):
        return iter((self, True))

    def __len__(self):
        return 2

    def solve(self, z: np.ndarray) -> np.ndarray:
        zz = np.asarray(z, dtype=np.float64)
        return self.factor.solve(zz.reshape(zz.shape[0], -1)).reshape(zz.shape)


def enable_quad_cholesky() -> None:
    """Route ``BQBeliefUpdate.compute_gram_cho_factor`` / ``gram_cho_solve``
    (quad/solvers/belief_updates/_belief_update.py:61-103) through the GPU.  At 16384 nodes
    the host ``dpotrf`` (1.5 TFLOP) takes ~15 s and dominates ``bayesquad_from_data`` once the
    Gram matrix itself is built on the device."""
    numpy2_compat()
    import probnum  # noqa: F401, PLC0415
    from probnum.quad.solvers.belief_updates import _belief_update as bu  # noqa: PLC0415

    cls = bu.BQBeliefUpdate
    if ("quad", cls) in _ORIGINALS:
        factor0, solve0 = _ORIGINALS.pop(("quad", cls))  # enabling twice: start from the originals
        cls.compute_gram_cho_factor = factor0
        cls.gram_cho_solve = solve0
    _ORIGINALS[("quad", cls)] = (cls.__dict__["compute_gram_cho_factor"], cls.__dict__["gram_cho_solve"])

    def compute_gram_cho_factor(self, gram):
        # raises numpy.linalg.LinAlgError for a non-positive pivot, like scipy.  The array that is
        # handed in is what gets factorised (uploaded), exactly like the reference.  Opt-in
        # (TRUST_GRAM_PROVENANCE): a Gram matrix that this module built itself and whose sampled
        # entries are unchanged is re-evaluated on the device straight into the factorisation
        # buffer -- a millisecond instead of an N^2 host-to-device copy, at the price of not
        # noticing an in-place edit that misses the samples.
        origin = (_mirror.gram_provenance(gram)
                  if TRUST_GRAM_PROVENANCE and isinstance(gram, np.ndarray) else None)
        if origin is not None:
            params, pts = origin
            with _mirror._lib.Operator(params, pts) as op:
                chol = _mirror._lib.Cholesky(op, jitter=float(self.jitter))
        else:
            # cho_factor(a) reads the upper triangle (lower=False): so does the upload, and only that half crosses PCIe
            chol = _mirror._lib.Cholesky(gram, jitter=float(self.jitter), upper=True)
        return DeviceCholesky(chol), True

    def gram_cho_solve(gram_cho_factor, z):
        chol = gram_cho_factor[0] if isinstance(gram_cho_factor, tuple) else gram_cho_factor
        if isinstance(chol, DeviceCholesky):
            return chol.solve(z)
        return _ORIGINALS[("quad", cls)][1].__func__(gram_cho_factor, z)

    cls.compute_gram_cho_factor = compute_gram_cho_factor
    cls.gram_cho_solve = staticmethod(gram_cho_solve)
    _enable_quad_prediction()


# compute_gram_cho_factor: rebuild a Gram matrix of known provenance on the device instead of uploading the
# host array (sampled check only, see there).  Off by default: the reference always factors the array it is
# handed.  PNCOV_TRUST_GRAM_PROVENANCE=1 in the environment switches it on.
TRUST_GRAM_PROVENANCE = os.environ.get("PNCOV_TRUST_GRAM_PROVENANCE", "0") not in ("", "0", "off", "false")

# pncov_chol_solve_gram mode of the patched prediction path.  False: both substitutions and the sums in
# the reference's order.  True: everything from L^-1 K(nodes, x) alone (half the flops; sums of squares
# instead of cancelling sums, so the result is the more accurate one when the Gram matrix is
# ill-conditioned -- it then differs from the reference's by the reference's own rounding error).
PREDICTION_FORWARD_ONLY = False


def _device_prediction(x, bq_state, vecs):
    """``(k(x, x), quad, dots)`` of ``Cholesky.solve_gram`` for a BQ state whose factor lives on the
    device, or None when the state is outside that path (no data yet, host factor, a kernel without
    a single-launch Gram)."""
    chol = bq_state.gram_cho_factor
    chol = chol[0] if isinstance(chol, tuple) else chol
    if not isinstance(chol, DeviceCholesky) or bq_state.fun_evals.shape[0] == 0:
        return None
    twin = _mirror_of(bq_state.kernel)
    try:
        params = twin._kernel_params()
    except (AttributeError, NotImplementedError):  # outside the scope, or an arithmetic node
        return None
    x = np.asarray(x, dtype=np.float64)
    nodes = np.asarray(bq_state.nodes, dtype=np.float64)
    if x.ndim != 2 or nodes.ndim != 2 or x.shape[0] == 0:
        return None
    with _mirror._lib.Operator(params, nodes, x) as op:
        quad, dots = chol.factor.solve_gram(op, vecs, forward_only=PREDICTION_FORWARD_ONLY)
    return bq_state.kernel(x, x), quad, dots


def _enable_quad_prediction() -> None:
    """``BQStandardBeliefUpdate.predict_integrand``
    (quad/solvers/belief_updates/_standard_update.py:131-151) and
    ``IntegralVarianceReduction.__call__`` (acquisition_functions/_integral_variance_reduction.py:60-75)
    with the cross Gram ``K(nodes, x)``, the solve against the device factor and the column
    reductions in one device pass (``pncov_chol_solve_gram``); only length-m vectors come back."""
    from probnum.quad.solvers.acquisition_functions import _integral_variance_reduction as ivr  # noqa: PLC0415
    from probnum.quad.solvers.belief_updates import _standard_update as su  # noqa: PLC0415

    upd, acq = su.BQStandardBeliefUpdate, ivr.IntegralVarianceReduction
    if ("quad_predict", upd) in _ORIGINALS:
        return
    _ORIGINALS[("quad_predict", upd)] = upd.__dict__["predict_integrand"]
    _ORIGINALS[("quad_predict", acq)] = acq.__dict__["__call__"]
    predict0 = upd.__dict__["predict_integrand"].__func__
    call0 = acq.__dict__["__call__"]

    def predict_integrand(x, bq_state):
        res = _device_prediction(x, bq_state, [bq_state.fun_evals])
        if res is None:
            return predict0(x, bq_state)
        prior_var, quad, dots = res
        return dots[0], bq_state.scale_sq * (prior_var - quad)

    def __call__(self, x, bq_state):
        res = _device_prediction(x, bq_state, [bq_state.kernel_means])
        if res is None:
            return call0(self, x, bq_state)
        prior_var, quad, dots = res
        y_predictive_var = bq_state.scale_sq * (prior_var - quad) + 0.0  # observation noise placeholder of the reference
        predictive_embedding = bq_state.kernel_embedding.kernel_mean(x) - dots[0]
        values = (bq_state.scale_sq * predictive_embedding) ** 2 / (bq_state.integral_belief.cov * y_predictive_var)
        return values, None

    upd.predict_integrand = staticmethod(predict_integrand)
    acq.__call__ = __call__


# ---------------------------------------------------------------------------------------------
# the update chains of the matrix-based solver on the device
# ---------------------------------------------------------------------------------------------
_CHAIN_CLASSES = None


def _make_chain_classes():
    """``ChainOperator`` subclasses of ``probnum.linops.LinearOperator``: the sum of the first k
    updates of an ``_lib.UpdateChain``.  A fresh update (one vector pair, no device storage yet) is a
    chain of length one; adding it to a sum that already holds a chain extends that chain instead of
    growing the sum (registered in the reference's own dispatch table ``linops._arithmetic._add_fns``)."""
    global _CHAIN_CLASSES
    if _CHAIN_CLASSES is not None:
        return _CHAIN_CLASSES
    import probnum  # noqa: F401, PLC0415
    from probnum import linops  # noqa: PLC0415
    from probnum.linops import _arithmetic as arith  # noqa: PLC0415
    from probnum.linops._arithmetic_fallbacks import SumLinearOperator  # noqa: PLC0415

    UpdateChain = _mirror._lib.UpdateChain

    class ChainOperator(linops.LinearOperator):
        MODE = UpdateChain.SYMMETRIC_RANK2

        def __init__(self, n, pairs, store=None, transposed=False):
            self._pairs = pairs      # tuple of (u, v) host vectors of length n
            self._store = store      # shared _lib.UpdateChain holding a prefix-compatible history, or None
            self._transposed = transposed
            super().__init__(shape=(n, n), dtype=np.dtype(np.float64))

        # -- chain bookkeeping ----------------------------------------------------------------
        def extended(self, other: "ChainOperator") -> "ChainOperator":
            """This chain followed by the pairs of ``other`` (a new operator; storage is shared while
            the history is linear, which is how the solver uses it)."""
            n = self.shape[0]
            store = self._store
            if store is None or len(store) != len(self._pairs):  # first merge, or a fork: own storage
                store = UpdateChain(n)
                for u, v in self._pairs:
                    store.append(u, v)
            for u, v in other._pairs:
                store.append(u, v)
            return type(self)(n, self._pairs + other._pairs, store, self._transposed)

        def _mode(self) -> int:
            return self.MODE

        # -- LinearOperator interface --------------------------------------------------------------
        def _matmul(self, x: np.ndarray) -> np.ndarray:
            n, k = self.shape[0], len(self._pairs)
            if self._store is None:
                self._store = UpdateChain(n)
                for u, v in self._pairs:
                    self._store.append(u, v)
            x = np.asarray(x, dtype=np.float64)
            if x.ndim == 2:
                return self._store.apply(k, self._mode(), x)
            batch = x.shape[:-2]
            cols = np.moveaxis(x.reshape((-1,) + x.shape[-2:]), 0, -1).reshape(n, -1)  # (n, c * B)
            y = self._store.apply(k, self._mode(), cols).reshape(n, x.shape[-1], -1)
            return np.moveaxis(y, -1, 0).reshape(batch + (n, x.shape[-1]))

        def _factors(self):
            U = np.stack([u for u, _ in self._pairs], axis=1)
            V = np.stack([v for _, v in self._pairs], axis=1)
            return U, V

        def _todense(self) -> np.ndarray:
            U, V = self._factors()
            return U @ V.T + V @ U.T

        def _transpose(self):
            return self

        def _trace(self):
            return 2.0 * sum(float(u @ v) for u, v in self._pairs)

        def _diagonal(self) -> np.ndarray:
            U, V = self._factors()
            return 2.0 * np.sum(U * V, axis=1)

    class Rank1ChainOperator(ChainOperator):
        """sum_i Ws_i u_i^T with pairs stored as (u_i, Ws_i)."""

        def _mode(self) -> int:
            return UpdateChain.RANK1_TRANSPOSED if self._transposed else UpdateChain.RANK1

        def _todense(self) -> np.ndarray:
            U, V = self._factors()
            return U @ V.T if self._transposed else V @ U.T

        def _transpose(self):
            return Rank1ChainOperator(self.shape[0], self._pairs, self._store, not self._transposed)

        def _trace(self):
            return sum(float(u @ v) for u, v in self._pairs)

        def _diagonal(self) -> np.ndarray:
            U, V = self._factors()
            return np.sum(U * V, axis=1)

    def add_to_sum(total, update):
        """SumLinearOperator + fresh update: extend the chain the sum already holds, if any."""
        parts = list(total._summands)
        for i, part in enumerate(parts):
            if type(part) is type(update) and part._transposed == update._transposed:
                parts[i] = part.extended(update)
                return SumLinearOperator(*parts)
        return SumLinearOperator(total, update)

    def add_chains(chain, update):
        if chain._transposed != update._transposed:
            return SumLinearOperator(chain, update)
        return chain.extended(update)

    for cls in (ChainOperator, Rank1ChainOperator):
        arith._add_fns[(SumLinearOperator, cls)] = add_to_sum
        arith._add_fns[(cls, cls)] = add_chains
    _CHAIN_CLASSES = (ChainOperator, Rank1ChainOperator)
    return _CHAIN_CLASSES


class _GrowingHstack:
    """``np.hstack`` for the solver's growing lists of (n, 1) columns.

    The reference's solver loop calls ``np.hstack(self.obs_list)`` / ``np.hstack(self.search_dir_list)`` two to
    three times per iteration (linalg/solvers/matrixbased.py:615-616, :661, :927, :936): O(n k) copies of k small
    arrays each time, O(n k^2) over a solve -- 0.7 s of a 1.2 s solve at n = 65536, k = 50 once the products run on
    the device.  A list that only grew since the last call (same leading array objects) gets its new columns
    appended to a C-ordered buffer, and the caller receives the same values in the same (row-major) orientation, as
    a view of that buffer.  Anything else goes to NumPy.  The module-level name ``np`` of ``matrixbased`` is pointed
    at this shim by :func:`enable_solver_chains` and restored by :func:`disable`."""

    def __init__(self):
        # id(list) -> [columns held, buffer (n, capacity)]; the columns themselves are kept (not their ids) so that
        # an address reused by a later solve can never be mistaken for a column already copied
        self._entries = {}

    def __getattr__(self, name):
        return getattr(np, name)

    def hstack(self, tup, *args, **kwargs):
        if args or kwargs or not isinstance(tup, list) or len(tup) < 2:
            return np.hstack(tup, *args, **kwargs)
        first = tup[0]
        if not (isinstance(first, np.ndarray) and first.ndim == 2 and first.shape[1] == 1 and first.dtype == np.float64
                and first.shape[0] >= 4096):
            return np.hstack(tup)
        n, k = first.shape[0], len(tup)
        entry = self._entries.get(id(tup))
        if (entry is None or entry[1].shape[0] != n or len(entry[0]) > k
                or any(a is not b for a, b in zip(entry[0], tup))):
            entry = [[], np.empty((n, max(64, 2 * k)), dtype=np.float64)]
            if len(self._entries) >= 8:
                self._entries.pop(next(iter(self._entries)))
            self._entries[id(tup)] = entry
        held, buf = entry
        if k > buf.shape[1]:
            grown = np.empty((n, 2 * k), dtype=np.float64)
            grown[:, :len(held)] = buf[:, :len(held)]
            entry[1] = buf = grown
        for j in range(len(held), k):
            col = tup[j]
            if not (isinstance(col, np.ndarray) and col.shape == (n, 1) and col.dtype == np.float64):
                self._entries.pop(id(tup), None)
                return np.hstack(tup)
            buf[:, j] = col[:, 0]
        entry[0] = list(tup)
        return buf[:, :k]


def enable_solver_chains() -> None:
    """Route ``SymmetricMatrixBasedSolver._mean_update`` / ``_covariance_update``
    (linalg/solvers/matrixbased.py:751-775) through device-resident update chains.  The solver's own
    loop is untouched: it still writes ``aslinop(mean) + self._mean_update(u, v)``; the sum it gets
    back holds ONE chain operator instead of k lambdas, and a product with it is two kernel launches."""
    numpy2_compat()
    import probnum  # noqa: F401, PLC0415
    from probnum.linalg.solvers import matrixbased as mb  # noqa: PLC0415

    cls = mb.SymmetricMatrixBasedSolver
    if ("solver_chain", cls) in _ORIGINALS:
        return
    Rank2, Rank1 = _make_chain_classes()
    _ORIGINALS[("solver_chain", cls)] = (cls.__dict__["_mean_update"], cls.__dict__["_covariance_update"])

    def _vec(a, n):
        a = np.asarray(a, dtype=np.float64)
        return a.reshape(n).copy() if a.size == n else None

    def _mean_update(self, u, v):
        uu, vv = _vec(u, self.n), _vec(v, self.n)
        if uu is None or vv is None:  # several right-hand sides: not a vector pair
            return _ORIGINALS[("solver_chain", cls)][0](self, u, v)
        return Rank2(self.n, ((uu, vv),))

    def _covariance_update(self, u, Ws):
        uu, ww = _vec(u, self.n), _vec(Ws, self.n)
        if uu is None or ww is None:
            return _ORIGINALS[("solver_chain", cls)][1](self, u, Ws)
        return Rank1(self.n, ((uu, ww),))

    cls._mean_update = _mean_update
    cls._covariance_update = _covariance_update
    # the loop's repeated np.hstack over its growing column lists (see _GrowingHstack)
    if not isinstance(mb.np, _GrowingHstack):
        _ORIGINALS[("solver_numpy", mb)] = mb.np
        mb.np = _GrowingHstack()


def enable_all(sharded: Optional[bool] = None) -> None:
    """:func:`enable` plus the opt-in rows of SURVEY.md 8f: the device Cholesky and prediction path of
    the BQ belief update (N1, N4) and the solver's update chains (N2 ii)."""
    enable(sharded)
    enable_quad_cholesky()
    enable_solver_chains()


def disable() -> None:
    """Undo :func:`enable`, :func:`enable_quad_cholesky` and :func:`enable_solver_chains`."""
    global _SHARDED, _LOCAL_FACTORY
    _SHARDED, _LOCAL_FACTORY = None, None
    quad = [k for k in _ORIGINALS if isinstance(k, tuple) and k[0] == "quad"]
    for key in quad:
        factor, solve = _ORIGINALS.pop(key)
        key[1].compute_gram_cho_factor = factor
        key[1].gram_cho_solve = solve
    for key in [k for k in _ORIGINALS if isinstance(k, tuple) and k[0] == "solver_chain"]:
        mean_update, covariance_update = _ORIGINALS.pop(key)
        key[1]._mean_update = mean_update
        key[1]._covariance_update = covariance_update
    for key in [k for k in _ORIGINALS if isinstance(k, tuple) and k[0] == "solver_numpy"]:
        key[1].np = _ORIGINALS.pop(key)
    for key in [k for k in _ORIGINALS if isinstance(k, tuple) and k[0] == "quad_predict"]:
        orig = _ORIGINALS.pop(key)
        setattr(key[1], "predict_integrand" if isinstance(orig, staticmethod) else "__call__", orig)
    for key in [k for k in _ORIGINALS if isinstance(k, tuple) and k[0] == "matrix"]:
        orig = _ORIGINALS.pop(key)
        if "_evaluate_matrix" in key[1].__dict__:
            delattr(key[1], "_evaluate_matrix")
        if getattr(key[1], "_evaluate_matrix") is not orig:
            key[1]._evaluate_matrix = orig
    for cls, (matrix, linop, evaluate) in _ORIGINALS.items():
        # the two hooks are inherited from CovarianceFunction: remove our overrides
        for name, orig in (("_evaluate_matrix", matrix), ("_evaluate_linop", linop)):
            if name in cls.__dict__:
                delattr(cls, name)
            if getattr(cls, name) is not orig:
                setattr(cls, name, orig)
        cls._evaluate = evaluate  # defined on the class itself
    _ORIGINALS.clear()
    # page-locked result blocks that no array uses any more go back to the system
    from . import hostmem  # noqa: PLC0415

    hostmem.release_cache()


def enabled() -> bool:
    return any(not isinstance(k, tuple) for k in _ORIGINALS)
