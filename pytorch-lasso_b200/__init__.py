"""lasso_b200 -- B200-native ISTA/FISTA sparse-encode engine.

Drop-in for the ISTA path of ``lasso.linear`` (rfeinman/pytorch-lasso):

    from lasso_b200.linear import sparse_encode, dict_learning

Host code is Python/PyTorch (device memory, streams, torch.distributed); the
arithmetic runs in ``csrc/liblasso_b200.so`` behind the C ABI declared in
``include/lasso_b200.h``.
"""
from . import _cabi  # noqa: F401
from . import linear  # noqa: F401
from . import conv2d  # noqa: F401
from . import testing  # noqa: F401
from ._cabi import LassoB200Error  # noqa: F401

__version__ = "0.1.0"
