"""Dictionary learning front end -- mirrors lasso/linear/dict_learning.py.

E-step = ``sparse_encode`` (the FISTA kernels).  M-step in Gram space: one pass
over (Z, X) produces A = Z^T Z and B = Z^T X; these -- plus the two loss sums --
are the only buffers that cross GPUs (one all-reduce per EM step when the batch
is row-sharded over a process group), after which every rank runs the same
tiny atom sweep so the dictionary stays replicated without a broadcast.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _cabi
from .sparse_encode import sparse_encode
from .utils import default_device

__all__ = ["lasso_loss", "dict_evaluate", "dict_learning", "update_dict", "update_dict_ridge"]


def _world(group):
    return torch.distributed.get_world_size(group) if group is not None else 1


def _global_rows(n, device, group):
    if _world(group) == 1:
        return n
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    torch.distributed.all_reduce(cnt, group=group)
    return int(cnt.item())


def lasso_loss(X, Z, weight, alpha=1.0, group=None):
    """(0.5 |X - Z W^T|^2 + alpha |Z|_1) / n as a float32 scalar tensor (dict_learning.py:10-13)."""
    terms = _cabi.loss_terms(X, Z, weight)
    if _world(group) > 1:
        torch.distributed.all_reduce(terms, group=group)
    n = _global_rows(X.size(0), X.device, group)
    return ((0.5 * terms[0] + alpha * terms[1]) / n).to(torch.float32)


def dict_evaluate(X, weight, alpha, group=None, **kwargs):
    """Held-out loss: encode then ``lasso_loss`` (dict_learning.py:16-20)."""
    X = X.to(weight.device)
    Z = sparse_encode(X, weight, alpha, group=group, **kwargs) if group is not None \
        else sparse_encode(X, weight, alpha, **kwargs)
    return lasso_loss(X, Z, weight, alpha, group=group)


def _statistics(X, Z, group):
    gzz, gzx = _cabi.gram(Z, X)
    if _world(group) > 1:
        k, d = gzx.shape
        packed = torch.cat([gzz.reshape(-1), gzx.reshape(-1)])
        torch.distributed.all_reduce(packed, group=group)
        gzz = packed[:k * k].reshape(k, k).contiguous()
        gzx = packed[k * k:].reshape(k, d).contiguous()
    return gzz, gzx


def update_dict(dictionary, X, Z, random_seed=None, positive=False, eps=1e-10, group=None):
    """Block-coordinate atom update, in place on ``dictionary`` (and on ``Z`` for
    degenerate atoms) like the reference (dict_learning.py:56-103)."""
    if random_seed is not None:
        torch.manual_seed(random_seed)
    if not (dictionary.is_cuda and dictionary.is_contiguous() and dictionary.dtype == torch.float32):
        raise _cabi.LassoB200Error("dictionary must be a contiguous float32 CUDA tensor")
    gzz, gzx = _statistics(X, Z, group)
    zeroed = _cabi.dict_update_gram(dictionary, gzz, gzx, eps=eps, redraw=None, positive=positive)
    if bool(zeroed.any()):  # one sync per sweep (the reference syncs once per atom)
        # degenerate atoms (dict_learning.py:91-98): fresh N(0,1) atoms, unit norm, their codes dropped.
        # All of them are drawn in ONE call (the reference draws them one by one inside its atom loop;
        # early EM steps of a large dictionary can have dozens, and a handful of tiny launches per
        # atom dominated the M-step).
        idx = zeroed.nonzero().flatten()
        d = dictionary.size(0)
        atoms = torch.empty(idx.numel(), d, device=dictionary.device).normal_()
        if positive:
            atoms.clamp_(0, None)      # dict_learning.py:94-95
        if _world(group) > 1:
            torch.distributed.broadcast(atoms, src=torch.distributed.get_global_rank(group, 0)
                                        if group is not torch.distributed.group.WORLD else 0,
                                        group=group)
        dictionary[:, idx] = (atoms / atoms.norm(dim=1, keepdim=True)).T
        Z[:, idx] = 0
    return dictionary


def update_dict_ridge(x, z, lambd=1e-4, group=None):
    """Unconstrained M-step: V = ((Z^T Z + lambd n I)^-1 Z^T X)^T (dict_learning.py:106-123)."""
    gzz, gzx = _statistics(x, z, group)
    n = _global_rows(x.size(0), x.device, group)
    gzz.diagonal().add_(lambd * n)
    chol = torch.linalg.cholesky(gzz)
    return torch.cholesky_solve(gzx, chol).T.to(torch.float32).contiguous()


def dict_learning(X, n_components, alpha=1.0, constrained=True, persist=False,
                  lambd=1e-2, steps=60, device='cpu', progbar=True, group=None,
                  **solver_kwargs):
    """EM dictionary learning (dict_learning.py:23-53); returns ``(weight[d,k], losses[steps])``.

    ``device`` keeps its reference meaning for where the inputs/outputs live and
    which RNG draws the initial dictionary; the arithmetic runs on the current
    CUDA device either way.  With ``group`` each rank passes its row shard of X.
    """
    out_device = torch.device(device)
    compute = out_device if out_device.type == 'cuda' else default_device()
    n_features = X.shape[1]
    X = X.to(compute)
    weight = torch.empty(n_features, n_components, device=out_device)
    torch.nn.init.orthogonal_(weight)
    if constrained:
        weight = F.normalize(weight, dim=0)
    weight = weight.to(compute).contiguous()
    if _world(group) > 1:
        torch.distributed.broadcast(weight, src=torch.distributed.get_global_rank(group, 0)
                                    if group is not torch.distributed.group.WORLD else 0,
                                    group=group)
    if group is not None:
        solver_kwargs = dict(solver_kwargs, group=group)
    Z0 = None
    losses = torch.zeros(steps, device=compute)

    try:
        from tqdm import tqdm
        bar = tqdm(total=steps, disable=not progbar)
    except ImportError:  # pragma: no cover
        bar = None
    for i in range(steps):
        Z = sparse_encode(X, weight, alpha, Z0, **solver_kwargs)
        losses[i] = lasso_loss(X, Z, weight, alpha, group=group)
        if persist:
            Z0 = Z
        if constrained:
            weight = update_dict(weight, X, Z, group=group)
        else:
            weight = update_dict_ridge(X, Z, lambd=lambd, group=group)
        if bar is not None and progbar:
            bar.set_postfix(loss=losses[i].item())
            bar.update(1)
    if bar is not None:
        bar.close()
    return weight.to(out_device), losses.to(out_device)
