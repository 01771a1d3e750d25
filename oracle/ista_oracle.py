"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's ISTA/FISTA path.

The reference (rfeinman/pytorch-lasso) is pure Python whose arithmetic lives in
``torch`` (ATen matmul / softshrink / elementwise / reductions; MKL on CPU) and
``scipy`` (ARPACK ``eigsh``).  This module restates the algorithm of

    lasso/linear/solvers/ista.py          (ista.py:8-104)
    lasso/linear/sparse_encode.py         (sparse_encode.py:19-73, ista branch)
    lasso/linear/dict_learning.py         (dict_learning.py:10-123)
    lasso/linear/utils.py                 (utils.py:28-40, ridge)

with the SAME torch calls at the reference's call sites and in the same order,
so that on one machine the fp32 results are bit-identical to the reference run
with ``lr`` pinned.  ``ista_f64`` is a numpy float64 restatement used as the
"gold" solution (noise floor of the fp32 reference: 5e-7..3e-6 relative).

Pinned by ``tests/golden/*.npz`` (outputs of the real reference, produced by
``tests/golden/make_golden.py`` in the build container).  Never imported by the
product package.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# step size / momentum schedule
# --------------------------------------------------------------------------

def lipschitz_constant(weight: torch.Tensor) -> float:
    """Largest eigenvalue of W^T W (ista.py:8-14).

    The reference calls ARPACK ``eigsh`` on the float32 Gram, which is not
    run-to-run reproducible.  The oracle uses a dense float64
    eigensolve of the smaller of W^T W / W W^T -- same quantity, deterministic.
    """
    w = weight.detach().to(torch.float64).cpu()
    d, k = w.shape
    gram = w @ w.T if d <= k else w.T @ w
    return float(torch.linalg.eigvalsh(gram)[-1])


def beta_schedule(maxiter: int) -> list:
    """FISTA momentum coefficients (t-1)/t_next as Python floats (ista.py:77-78, 98-101)."""
    out, t = [], 1
    for _ in range(maxiter):
        t_next = (1 + math.sqrt(1 + 4 * t ** 2)) / 2
        out.append((t - 1) / t_next)
        t = t_next
    return out


# --------------------------------------------------------------------------
# ISTA / FISTA
# --------------------------------------------------------------------------

def _gradient(code, x, weight):
    # ista.py:71-73 -- residual first, then back-projection
    residual = torch.matmul(code, weight.T) - x
    return torch.matmul(residual, weight)


def _line_search(point, x, weight, alpha, step0, shrink=1.5, max_trials=1000):
    """Beck-Teboulle backtracking with batch-global F and Q (ista.py:17-54)."""
    if shrink <= 1:
        raise ValueError('eta must be > 1.')
    residual0 = torch.matmul(point, weight.T) - x
    f0 = 0.5 * residual0.pow(2).sum()
    grad0 = torch.matmul(residual0, weight)

    step = step0
    for _ in range(max_trials):
        cand = F.softshrink(point - step * grad0, alpha * step)
        res1 = torch.matmul(cand, weight.T) - x
        big_f = 0.5 * res1.pow(2).sum() + alpha * cand.abs().sum()
        dz = cand - point
        big_q = (f0 + (dz * grad0).sum() + (0.5 / step) * dz.pow(2).sum()
                 + alpha * cand.abs().sum())
        if big_f <= big_q:
            return cand, step
        step = step / shrink
    warnings.warn('backtracking line search failed. Reverting to initial '
                  'step size')
    return F.softshrink(point - step0 * grad0, alpha * step0), step0


def ista(x, z0, weight, alpha=1.0, fast=True, lr='auto', maxiter=10,
         tol=1e-5, backtrack=False, eta_backtrack=1.5, return_info=False):
    """Restatement of ista.py:57-104 (verbose printing omitted).

    ``return_info=True`` additionally returns the number of executed iterations
    and the per-iteration sum |z - z_next| (not part of the reference API; used
    by the parity tests of the deferred stopping rule).
    """
    if lr == 'auto':
        lr = 1 / lipschitz_constant(weight)
    threshold = z0.numel() * tol

    z = z0
    y, t = z0, 1
    deltas = []
    done = 0
    for _ in range(maxiter):
        point = y if fast else z
        if backtrack:
            z_new, _ = _line_search(point, x, weight, alpha, lr, eta_backtrack)
        else:
            z_new = F.softshrink(point - lr * _gradient(point, x, weight), alpha * lr)
        done += 1
        delta = (z - z_new).abs().sum()
        deltas.append(float(delta))
        if delta <= threshold:
            z = z_new
            break
        if fast:
            t_new = (1 + math.sqrt(1 + 4 * t ** 2)) / 2
            y = z_new + ((t - 1) / t_new) * (z_new - z)
            t = t_new
        z = z_new
    if return_info:
        return z, done, deltas
    return z


def ista_f64(x, z0, weight, alpha, lr, maxiter, fast=True, tol=0.0):
    """float64 numpy gold restatement of ista.py:79-102 (no backtracking)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    z = np.asarray(z0, dtype=np.float64)
    y, t = z, 1.0
    threshold = z.size * tol
    lam = alpha * lr
    for _ in range(maxiter):
        p = y if fast else z
        v = p - lr * (((p @ w.T) - x) @ w)
        z_new = np.sign(v) * np.maximum(np.abs(v) - lam, 0.0)
        if np.abs(z - z_new).sum() <= threshold:
            z = z_new
            break
        if fast:
            t_new = (1 + math.sqrt(1 + 4 * t * t)) / 2
            y = z_new + ((t - 1) / t_new) * (z_new - z)
            t = t_new
        z = z_new
    return z


# --------------------------------------------------------------------------
# sparse_encode boundary (ista branch only)
# --------------------------------------------------------------------------

def _ridge(b, a, alpha):
    # utils.py:28-40
    rhs = torch.matmul(a.T, b)
    gram = torch.matmul(a.T, a)
    gram.diagonal().add_(alpha)
    chol, info = torch.linalg.cholesky_ex(gram)
    if info != 0:
        raise RuntimeError("The Gram matrix is not positive definite. "
                           "Try increasing 'alpha'.")
    return torch.cholesky_solve(rhs, chol)


def _lstsq(b, a):
    # utils.py:13-25: QR least-norm solution when the system is under-determined (d < k),
    # QR least-squares solution otherwise
    m, n = a.shape[-2:]
    if m < n:
        q, r = torch.linalg.qr(a.transpose(-1, -2), mode='reduced')
        d = torch.linalg.solve_triangular(r.transpose(-1, -2), b, upper=False)
        return torch.matmul(q, d)
    q, r = torch.linalg.qr(a, mode='reduced')
    d = torch.matmul(q.transpose(-1, -2), b)
    return torch.linalg.solve_triangular(r, d, upper=True)


def initialize_code(x, weight, alpha, mode):
    """sparse_encode.py:19-35."""
    n, k = x.size(0), weight.size(1)
    if mode == 'zero':
        return x.new_zeros(n, k)
    if mode == 'unif':
        return x.new(n, k).uniform_(-0.1, 0.1)
    if mode == 'lstsq':
        return _lstsq(x.T, weight).T
    if mode == 'ridge':
        return _ridge(x.T, weight, alpha).T
    if mode == 'transpose':
        return torch.matmul(x, weight)
    raise ValueError("invalid init parameter '{}'.".format(mode))


def sparse_encode(x, weight, alpha=1.0, z0=None, algorithm='ista', init=None, **kwargs):
    """sparse_encode.py:38-73 restricted to algorithm='ista'."""
    n, k = x.size(0), weight.size(1)
    if z0 is not None:
        assert z0.shape == (n, k)
    else:
        z0 = initialize_code(x, weight, alpha, 'zero' if init is None else init)
    if algorithm != 'ista':
        raise ValueError("invalid algorithm parameter '{}'.".format(algorithm))
    return ista(x, z0, weight, alpha, **kwargs)


# --------------------------------------------------------------------------
# dictionary learning
# --------------------------------------------------------------------------

def lasso_loss(x, z, weight, alpha=1.0):
    # dict_learning.py:10-13
    recon = torch.matmul(z, weight.T)
    return (0.5 * (x - recon).pow(2).sum() + alpha * z.abs().sum()) / x.size(0)


def update_dict(dictionary, x, z, eps=1e-10, redraw=None, positive=False):
    """Sequential atom update of dict_learning.py:56-103 (``positive``: the clamp of :87-88, :94-95).

    In place on ``dictionary`` and ``z`` like the reference.  ``redraw`` is an
    optional callable ``(d,) -> tensor`` that supplies the replacement for a
    degenerate atom instead of the global RNG (dict_learning.py:93).
    """
    resid = x - torch.matmul(z, dictionary.T)
    for j in range(dictionary.size(1)):
        resid += torch.outer(z[:, j], dictionary[:, j])
        dictionary[:, j] = torch.matmul(z[:, j], resid)
        if positive:
            dictionary[:, j].clamp_(0, None)
        nrm = dictionary[:, j].norm()
        if nrm < eps:
            if redraw is None:
                dictionary[:, j].normal_()
            else:
                dictionary[:, j] = redraw(dictionary.size(0))
            if positive:
                dictionary[:, j].clamp_(0, None)
            dictionary[:, j] /= dictionary[:, j].norm()
            z[:, j].zero_()
        else:
            dictionary[:, j] /= nrm
            resid -= torch.outer(z[:, j], dictionary[:, j])
    return dictionary


def update_dict_gram(dictionary, gram_zz, gram_zx, eps=1e-10, redraw=None, positive=False):
    """Gram-space restatement of the same Gauss-Seidel sweep.

    With A = Z^T Z (k x k) and B = Z^T X (k x d) the un-normalised atom is
    u_j = B[j] - D A[:, j] + A[j, j] d_j, evaluated with the already-updated
    atoms.  float64 so it can serve as the checker of the
    CUDA single-CTA kernel.  Returns (dictionary, zeroed_atoms).
    """
    dmat = dictionary.to(torch.float64).clone()
    a = gram_zz.to(torch.float64).clone()
    b = gram_zx.to(torch.float64).clone()
    zeroed = []
    for j in range(dmat.size(1)):
        u = b[j] - dmat @ a[:, j] + a[j, j] * dmat[:, j]
        if positive:
            u = u.clamp(min=0)
        nrm = u.norm()
        if nrm < eps:
            u = torch.randn(dmat.size(0), dtype=torch.float64) if redraw is None \
                else redraw(dmat.size(0)).to(torch.float64)
            if positive:
                u = u.clamp(min=0)
            dmat[:, j] = u / u.norm()
            a[j, :] = 0
            a[:, j] = 0
            b[j, :] = 0
            zeroed.append(j)
        else:
            dmat[:, j] = u / nrm
    return dmat.to(dictionary.dtype), zeroed


def update_dict_ridge(x, z, lambd=1e-4):
    # dict_learning.py:106-123
    rhs = torch.mm(z.T, x)
    m = torch.mm(z.T, z)
    m.diagonal().add_(lambd * x.size(0))
    chol = torch.linalg.cholesky(m)
    return torch.cholesky_solve(rhs, chol).T


def dict_learning(x, n_components, alpha=1.0, constrained=True, persist=False,
                  lambd=1e-2, steps=60, weight0=None, **solver_kwargs):
    """EM loop of dict_learning.py:23-53 on CPU, no progress bar.

    ``weight0`` injects the initial dictionary (the reference always draws it
    with nn.init.orthogonal_ -- dict_learning.py:28-31 -- which is what happens
    here when it is None).
    """
    n, d = x.shape
    if weight0 is None:
        weight = torch.empty(d, n_components)
        torch.nn.init.orthogonal_(weight)
        if constrained:
            weight = F.normalize(weight, dim=0)
    else:
        weight = weight0.clone()
    z0 = None
    losses = torch.zeros(steps)
    for i in range(steps):
        z = sparse_encode(x, weight, alpha, z0, **solver_kwargs)
        losses[i] = lasso_loss(x, z, weight, alpha)
        if persist:
            z0 = z
        if constrained:
            weight = update_dict(weight, x, z)
        else:
            weight = update_dict_ridge(x, z, lambd=lambd)
    return weight, losses


# --------------------------------------------------------------------------
# convolutional ISTA / FISTA  (lasso/conv2d/ista.py:7-49)
# --------------------------------------------------------------------------

def conv2d_ista(x, z0, weight, alpha=1.0, stride=1, padding=0, fast=True, maxiter=10, lr=None,
                tol=1e-5, return_iters=False):
    """Restatement of ``ista_conv2d`` with the reference's torch calls in the reference's order:
    the decoder is ``conv_transpose2d`` (ista.py:18), its adjoint ``conv2d`` (ista.py:19), the
    momentum update comes BEFORE the stop test (ista.py:40-47).  ``lr`` must be a float here
    (the 'auto' bound is checked separately against the reference)."""
    thresh = z0.numel() * tol                                     # ista.py:16

    def grad(code):
        recon = F.conv_transpose2d(code, weight, stride=stride, padding=padding)
        return F.conv2d(recon - x, weight, stride=stride, padding=padding)

    z = z0
    point, t = z0, 1
    done = 0
    for _ in range(maxiter):
        done += 1
        base = point if fast else z
        z_next = F.softshrink(base - lr * grad(base), alpha * lr)   # ista.py:27-28
        if fast:
            t_next = (1 + math.sqrt(1 + 4 * t ** 2)) / 2
            point = z_next + ((t - 1) / t_next) * (z_next - z)
            t = t_next
        if (z - z_next).abs().sum() <= thresh:
            z = z_next
            break
        z = z_next
    return (z, done) if return_iters else z
