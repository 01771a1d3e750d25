"""``sparse_encode`` / ``initialize_code`` -- mirrors lasso/linear/sparse_encode.py.

Same signature and error behaviour as the reference entry point
(sparse_encode.py:38-73); only ``algorithm='ista'`` is built (the hot path).
"""
from __future__ import annotations

import torch

from .. import _cabi
from .solvers.ista import ista as _ista_fn, solve as _solve
from .utils import default_device, lstsq

__all__ = ["sparse_encode", "initialize_code"]

# default start per algorithm (sparse_encode.py:8-16)
_DEFAULT_INIT = {'ista': 'zero', 'cd': 'zero', 'gpsr': 'zero', 'iter-ridge': 'ridge',
                 'interior-point': 'ridge', 'split-bregman': 'zero', 'own': 'zero'}
# solvers of the reference that are outside the hot path this build accelerates
_NOT_BUILT = ('cd', 'gpsr', 'iter-ridge', 'interior-point', 'split-bregman', 'own')


def _zeros(x, weight, alpha):
    return x.new_zeros(x.size(0), weight.size(1))


def _uniform(x, weight, alpha):
    return x.new_empty(x.size(0), weight.size(1)).uniform_(-0.1, 0.1)


def _ridge_start(x, weight, alpha):
    """Ridge warm start ((W^T W + alpha I)^-1 W^T x^T)^T (sparse_encode.py:28-29, utils.py:28-40) as ONE
    [n,d] x [d,k] product z0 = x T with T = (W W^T + alpha I)^-1 W = W (W^T W + alpha I)^-1 from the SMALLER
    of the two systems.  min(d,k) <= 64 (BASELINE configs 2, 4): all in the library
    (``lasso_b200_ridge_init_f32``: float64 Gram, one-CTA Cholesky, warp-per-column solves, FFMA product;
    0.30 ms at config 2 against 2.0 ms for the reference's formula in stock torch).  Larger systems (the
    notebook's 289): the one-CTA factorisation is slower than cuSOLVER there (1.7 ms vs 0.1 ms), so the
    m x m factor / solve stay in torch and only the n-sized product runs in the library."""
    if not x.is_cuda:
        dev = default_device()
        return _ridge_start(x.to(dev), weight.to(dev), alpha).to(x.device)
    d, k = weight.shape
    if min(d, k) <= 64:
        return _cabi.ridge_init(x, weight, alpha)
    by_rows = d <= k
    gram = weight @ weight.T if by_rows else weight.T @ weight
    gram.diagonal().add_(alpha)
    chol, info = torch.linalg.cholesky_ex(gram)
    if info != 0:
        raise RuntimeError("The Gram matrix is not positive definite. "
                           "Try increasing 'alpha'.")          # utils.py:36-38
    t = torch.cholesky_solve(weight, chol) if by_rows else torch.cholesky_solve(weight.T.contiguous(), chol).T
    return _cabi.matmul(x, t.contiguous())


_INITIALISERS = {
    'zero': _zeros,
    'unif': _uniform,
    'lstsq': lambda x, weight, alpha: lstsq(x.T, weight).T.contiguous(),
    'ridge': _ridge_start,
    'transpose': lambda x, weight, alpha: (_cabi.matmul(x, weight) if x.is_cuda else torch.matmul(x, weight)),
}


def initialize_code(x, weight, alpha, mode):
    """Starting code z0[n,k] (sparse_encode.py:19-35)."""
    try:
        make = _INITIALISERS[mode]
    except KeyError:
        raise ValueError("invalid init parameter '{}'.".format(mode)) from None
    return make(x, weight, alpha)


def sparse_encode(x, weight, alpha=1.0, z0=None, algorithm='ista', init=None, **kwargs):
    """Lasso codes z[n,k] of the rows of x for dictionary ``weight`` [d,k]."""
    if z0 is not None:
        assert z0.shape == (x.size(0), weight.size(1))
    if algorithm in _NOT_BUILT:
        raise NotImplementedError(
            "algorithm '{}' is outside the hot path lasso_b200 builds; only 'ista' "
            "(ISTA / FISTA) is available".format(algorithm))
    if algorithm != 'ista':
        raise ValueError("invalid algorithm parameter '{}'.".format(algorithm))

    if z0 is None:
        mode = _DEFAULT_INIT[algorithm] if init is None else init
        if mode == 'zero' and not kwargs.get('backtrack') and not kwargs.get('verbose') \
                and kwargs.get('maxiter', 10) != 0:
            # all-zero start: let the kernel clear its own buffer instead of
            # materialising and copying an [n,k] tensor
            opts = {key: kwargs[key] for key in kwargs
                    if key not in ('backtrack', 'eta_backtrack', 'verbose')}
            return _solve(x, None, weight, alpha=alpha, **opts)
        z0 = initialize_code(x, weight, alpha, mode)
    return _ista_fn(x, z0, weight, alpha, **kwargs)
