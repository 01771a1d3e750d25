"""ISTA / FISTA solver front end -- mirrors ``lasso.linear.solvers.ista``.

Same signature, defaults and error behaviour as the reference function
(lasso/linear/solvers/ista.py:57-58), but the loop of ista.py:79-102 runs in
``liblasso_b200.so`` (one fused kernel per iteration, no host synchronisation
between iterations).  Differences, all deliberate:

* ``lr='auto'`` uses an on-device float64 power iteration instead of ARPACK on
  the host (ista.py:8-14); the reference value itself varies ~1e-6 run to run.
* the batch-global stop test (ista.py:93) is evaluated on the device; kernels
  of later iterations turn into no-ops once it fires.
* ``backtrack=True`` (ista.py:17-54) and ``verbose=True`` (ista.py:80-81) take a
  slower path: the host drives one iteration at a time -- like the reference,
  whose line search branches on the host after every trial -- but every
  arithmetic step is a kernel of the library (gradient, trial, momentum, loss).
* CPU tensors go through the C ABI's host entry point (H2D, solve, D2H) -- the
  arithmetic always runs on the GPU.  There is no PyTorch fallback.
* extra keywords: ``path`` ('auto' | 'ffma' | 'tcgen05') selects the kernel,
  ``group`` a torch.distributed process group whose ranks hold row shards of one
  batch (the stop test is then made global with one deferred all-reduce), ``out``
  a preallocated result buffer.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch

from .. import utils as _utils
from ... import _cabi

__all__ = ["ista", "lipschitz_constant"]


def lipschitz_constant(weight: torch.Tensor, iters: int = 2000) -> float:
    """L = lambda_max(W^T W); replaces ``_lipschitz_constant`` (ista.py:8-14)."""
    if weight.is_cuda:
        return _cabi.lipschitz(weight.detach(), iters)
    dev = _utils.default_device()
    return _cabi.lipschitz(weight.detach().to(dev), iters)


def _validate(x, z0, weight):
    for name, t in (("x", x), ("weight", weight)) + ((("z0", z0),) if z0 is not None else ()):
        if not torch.is_tensor(t):
            raise TypeError("{} must be a tensor".format(name))
        if t.dtype != torch.float32:
            raise NotImplementedError(
                "lasso_b200 computes in float32 only; {} has dtype {}".format(name, t.dtype))
        if t.requires_grad:
            raise NotImplementedError(
                "lasso_b200 does not record an autograd graph through the solver; "
                "detach {} first".format(name))
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[0]:
        raise ValueError("expected x[n,d] and weight[d,k]; got {} and {}".format(
            tuple(x.shape), tuple(weight.shape)))
    if z0 is not None and tuple(z0.shape) != (x.shape[0], weight.shape[1]):
        raise ValueError("z0 must have shape [n,k]")
    if z0 is not None and z0.device != x.device:
        raise ValueError("z0 and x must live on the same device")
    if weight.device != x.device:
        raise ValueError("weight and x must live on the same device")


def _abs_tolerance(numel: int, tol: float) -> float:
    # ista.py:64 -- python float; torch then compares the float32 sum against the
    # scalar rounded to float32
    return float(np.float32(numel * tol))


def _first_stop(hist: torch.Tensor, tol_abs: float) -> int:
    """Number of iterations the reference would have executed given the global deltas."""
    hits = (hist[:-1] <= tol_abs).nonzero()
    return int(hits[0]) + 1 if hits.numel() else int(hist.numel())


def _global(values: torch.Tensor, group):
    """Sum a small device tensor of partial sums over the shards, return python floats."""
    if group is not None and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(values, group=group)
    return [float(v) for v in values.tolist()]


def solve(x, z0, weight, alpha=1.0, fast=True, lr='auto', maxiter=10, tol=1e-5,
          path='auto', group=None, return_iters=False, out=None):
    """Fast path (constant step): ``z0`` may be None for the all-zero start; ``out`` is an
    optional preallocated [n,k] float32 result buffer on x's device."""
    _validate(x, z0, weight)
    n, k = x.shape[0], weight.shape[1]
    if lr == 'auto':
        lr = 1.0 / lipschitz_constant(weight)
    lr = float(lr)
    alpha = float(alpha)
    sharded = group is not None and torch.distributed.get_world_size(group) > 1

    if not x.is_cuda:
        if sharded:
            raise NotImplementedError("sharded solves need CUDA tensors")
        z, iters = _cabi.fista_host(x, weight, z0, alpha, lr, maxiter, fast, _abs_tolerance(n * k, tol),
                                    path=path, want_iters=return_iters, out=out)
        return (z, iters) if return_iters else z

    if not sharded or maxiter <= 1:
        z, iters, _ = _cabi.fista_device(x, weight, z0, alpha, lr, maxiter, fast, _abs_tolerance(n * k, tol),
                                         path=path, want_iters=return_iters, out=out)
        return (z, iters) if return_iters else z

    # Row-sharded batch: run all iterations with the local test disabled, make the
    # per-iteration deltas global with ONE all-reduce (the shard's element count, which the
    # threshold z0.numel() * tol of ista.py:64 needs globally, rides in the same buffer), and replay a
    # shorter run in the (rare) case the global test fired early.  Deterministic kernels make the
    # replay identical to stopping in place (ista.py:93-95 semantics).
    z, _, hist = _cabi.fista_device(x, weight, z0, alpha, lr, maxiter, fast, -1.0,
                                    path=path, want_hist=True, out=out)
    buf = torch.cat([hist, hist.new_tensor([float(n * k)])])
    torch.distributed.all_reduce(buf, group=group)
    host = buf.tolist()                                   # the solve's one read-back
    tol_abs = _abs_tolerance(int(round(host[-1])), tol)
    done = maxiter
    for i, delta in enumerate(host[:maxiter - 1]):
        if delta <= tol_abs:
            done = i + 1
            break
    if done < maxiter:
        z, _, _ = _cabi.fista_device(x, weight, z0, alpha, lr, done, fast, -1.0, path=path,
                                     out=out)
    return (z, done) if return_iters else z


def _backtracking(point, x, weight, alpha, lr0, eta, group, maxiter=1000, verbose=False):
    """Beck-Teboulle line search with batch-global F and Q (ista.py:17-54); one scalar
    read-back per trial, exactly where the reference branches on the host."""
    grad, f_sum = _cabi.gradient(x, point, weight)
    (f_sum0,) = _global(f_sum, group)
    fval0 = 0.5 * f_sum0
    step = lr0
    cand = None
    for i in range(maxiter):
        cand, sums = _cabi.linesearch_trial(x, point, grad, weight, step, alpha, cand_out=cand)
        r2, l1, dzg, dz2 = _global(sums, group)
        big_f = 0.5 * r2 + alpha * l1
        big_q = fval0 + dzg + (0.5 / step) * dz2 + alpha * l1
        if verbose:
            print('iter: %4d,  t: %0.5f,  F-Q: %0.5f' % (i, step, big_f - big_q))
        if big_f <= big_q:
            return cand, step
        step = step / eta
    warnings.warn('backtracking line search failed. Reverting to initial '
                  'step size')
    cand, _ = _cabi.linesearch_trial(x, point, grad, weight, lr0, alpha, cand_out=cand)
    return cand, lr0


def _solve_stepwise(x, z0, weight, alpha, fast, lr, maxiter, tol, backtrack, eta_backtrack,
                    verbose, group):
    """Host-driven loop for backtrack / verbose (one iteration at a time, ista.py:79-102)."""
    _validate(x, z0, weight)
    if not x.is_cuda:
        dev = _utils.default_device()
        z = _solve_stepwise(x.to(dev), z0.to(dev), weight.to(dev), alpha, fast, lr, maxiter, tol,
                            backtrack, eta_backtrack, verbose, group)
        return z.cpu()
    if lr == 'auto':
        lr = 1.0 / lipschitz_constant(weight)
    lr, alpha = float(lr), float(alpha)
    numel = z0.numel()
    n_rows = x.shape[0]
    if group is not None and torch.distributed.get_world_size(group) > 1:
        cnt = torch.tensor([numel, n_rows], dtype=torch.int64, device=x.device)
        torch.distributed.all_reduce(cnt, group=group)
        numel, n_rows = int(cnt[0]), int(cnt[1])
    tol_abs = _abs_tolerance(numel, tol)

    z = z0.contiguous()
    y, t = z, 1
    for _ in range(maxiter):
        if verbose:
            r2, l1 = _global(_cabi.loss_terms(x, z, weight), group)
            print('loss: %0.4f' % ((0.5 * r2 + alpha * l1) / n_rows))
        point = y if fast else z
        if backtrack:
            z_next, _ = _backtracking(point, x, weight, alpha, lr, eta_backtrack, group)
        else:
            grad, _ = _cabi.gradient(x, point, weight)
            z_next, _ = _cabi.linesearch_trial(x, point, grad, weight, lr, alpha)
        beta = 0.0
        if fast:
            t_next = (1 + math.sqrt(1 + 4 * t ** 2)) / 2
            beta = (t - 1) / t_next
        y_next, delta = _cabi.momentum(z_next, z, beta, want_y=fast)
        (delta,) = _global(delta, group)
        if delta <= tol_abs:
            z = z_next
            break
        if fast:
            y, t = y_next, t_next
        z = z_next
    return z


def ista(x, z0, weight, alpha=1.0, fast=True, lr='auto', maxiter=10,
         tol=1e-5, backtrack=False, eta_backtrack=1.5, verbose=False,
         path='auto', group=None, out=None):
    """Drop-in for ``lasso.linear.solvers.ista`` (ista.py:57-104)."""
    if maxiter == 0:
        return z0  # the reference returns the z0 object itself
    if backtrack or verbose:
        if backtrack and eta_backtrack <= 1:
            raise ValueError('eta must be > 1.')  # ista.py:18-19
        z = _solve_stepwise(x, z0, weight, alpha, fast, lr, maxiter, tol, backtrack,
                            eta_backtrack, verbose, group)
        if out is not None:
            out.copy_(z)
            return out
        return z
    return solve(x, z0, weight, alpha=alpha, fast=fast, lr=lr, maxiter=maxiter, tol=tol,
                 path=path, group=group, out=out)
