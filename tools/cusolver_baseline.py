"""Benchmark baseline only (NOT part of the product): the library Cholesky SURVEY.md 8f N1 names --
cuSOLVER ``potrf`` / ``potrs`` through torch -- behind the same two hooks of the reference's
``BQBeliefUpdate`` that ``probnum_b200.patch.enable_quad_cholesky()`` routes to libpncov.

    import tools.cusolver_baseline as base
    base.enable(); ...; base.disable()
"""
import numpy as np

_saved = None


class _TorchCholesky:
    def __init__(self, factor):
        self.factor = factor

    def __iter__(self):
        return iter((self, True))

    def __len__(self):
        return 2


def enable(device_index: int = 0) -> None:
    global _saved
    import torch
    from probnum.quad.solvers.belief_updates import _belief_update as bu

    cls = bu.BQBeliefUpdate
    if _saved is None:
        _saved = (cls.__dict__["compute_gram_cho_factor"], cls.__dict__["gram_cho_solve"])
    device = torch.device("cuda", device_index)

    def compute_gram_cho_factor(self, gram):
        g = torch.from_numpy(np.ascontiguousarray(gram, dtype=np.float64)).to(device)
        g.diagonal().add_(self.jitter)
        factor, info = torch.linalg.cholesky_ex(g, upper=False)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(f"{int(info.item())}-th leading minor of the array is not positive definite")
        return _TorchCholesky(factor), True

    def gram_cho_solve(gram_cho_factor, z):
        chol = gram_cho_factor[0] if isinstance(gram_cho_factor, tuple) else gram_cho_factor
        zz = np.asarray(z, dtype=np.float64)
        rhs = torch.from_numpy(np.ascontiguousarray(zz.reshape(zz.shape[0], -1))).to(chol.factor.device)
        return torch.cholesky_solve(rhs, chol.factor, upper=False).cpu().numpy().reshape(zz.shape)

    cls.compute_gram_cho_factor = compute_gram_cho_factor
    cls.gram_cho_solve = staticmethod(gram_cho_solve)


def disable() -> None:
    global _saved
    if _saved is None:
        return
    from probnum.quad.solvers.belief_updates import _belief_update as bu

    bu.BQBeliefUpdate.compute_gram_cho_factor, bu.BQBeliefUpdate.gram_cho_solve = _saved
    _saved = None
