"""``sparse_encode`` / ``initialize_code`` -- mirrors lasso/linear/sparse_encode.py.

Same signature and error behaviour as the reference entry point
(sparse_encode.py:38-73); only ``algorithm='ista'`` is built (the hot path).
"""
from __future__ import annotations

import torch

from .solvers.ista import ista as _ista_fn, solve as _solve
from .utils import lstsq, ridge

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


_INITIALISERS = {
    'zero': _zeros,
    'unif': _uniform,
    'lstsq': lambda x, weight, alpha: lstsq(x.T, weight).T.contiguous(),
    'ridge': lambda x, weight, alpha: ridge(x.T, weight, alpha=alpha).T.contiguous(),
    'transpose': lambda x, weight, alpha: torch.matmul(x, weight),
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
