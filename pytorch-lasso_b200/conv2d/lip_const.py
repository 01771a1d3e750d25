"""Upper bound of the Lipschitz constant of a stride-1 convolution (Araujo et al.), the quantity the
reference's ``lip_bound_conv2d`` returns (lasso/conv2d/lip_const.py:96-135): the maximum over a
frequency grid of the per-frequency Gram trace of the kernel's Fourier symbol.  Host-side helper on a
tiny tensor; kept in torch like ``linear/utils.ridge``."""
from __future__ import annotations

import math

import torch

from .. import _cabi
from ..linear import utils as _utils

__all__ = ["lip_bound_conv2d", "lip_constant"]


def _one_int(value):
    """torch's conv2d also takes a pair per axis; a pair of equal integers means the same thing as the integer."""
    if isinstance(value, (tuple, list)) and len(value) == 2 and value[0] == value[1] and isinstance(value[0], int):
        return value[0]
    return value


@torch.no_grad()
def lip_constant(kernel, imsize, transpose=False, sqrt=False, stride=1, padding=0):
    """Largest eigenvalue of conv2d^T conv2d (``transpose=False``) / conv2d conv2d^T (``True``) on images
    of ``imsize`` -- the reference's ``lip_constant`` (lasso/conv2d/lip_const.py:8-31: ARPACK ``eigsh`` on a
    host ``LinearOperator``, a conv2d + conv_transpose2d + two host copies per Lanczos step) on the device
    (``lasso_b200_conv2d_lipschitz_f32``: the operator formed densely in image space, then the eigen-solver
    of the dictionary's Lipschitz constant).  Any kernel size, stride and padding (``stride`` / ``padding``
    are the reference's ``**kwargs``); ``cin*h*w <= 4096``.  Both operators share their non-zero spectrum, so
    ``transpose`` only says which side ``imsize`` describes: the image (False) or the code grid (True)."""
    stride, padding = _one_int(stride), _one_int(padding)
    if not (isinstance(stride, int) and isinstance(padding, int)):
        raise NotImplementedError("one integer stride / padding for both axes")
    out_channels, in_channels, kh, kw = kernel.shape
    height, width = imsize
    if transpose:       # imsize is the code grid: the image it decodes to
        height = (height - 1) * stride - 2 * padding + kh
        width = (width - 1) * stride - 2 * padding + kw
    dev = kernel.device if kernel.is_cuda else _utils.default_device()
    eig = _cabi.conv2d_lipschitz(kernel.detach().to(dev).contiguous(), (height, width), stride=stride,
                                 padding=padding)
    return math.sqrt(eig) if sqrt else eig


def lip_bound_conv2d(kernel, padding, stride=1, sample=50, sqrt=False):
    """Same arguments, checks and result as the reference function (lip_const.py:96-135)."""
    assert kernel.dim() == 4
    if kernel.size(-1) != kernel.size(-2):
        raise ValueError("The last 2 dim of the kernel must be equal.")
    if kernel.size(-1) % 2 != 1:
        raise ValueError("The dimension of the kernel must be odd.")
    if stride != 1:
        raise NotImplementedError("LipBound not implemented for stride > 1.")
    size = kernel.size(-1)
    taps = kernel.transpose(0, 1) if kernel.size(0) > kernel.size(1) else kernel
    taps = taps.flatten(2)                                       # [A, B, size*size], A <= B
    # symbol of the kernel at `sample` x `sample` frequencies in [0, 2 pi]^2
    freq = torch.linspace(0, 2 * math.pi, sample, device=kernel.device, dtype=kernel.dtype)
    f0, f1 = torch.meshgrid(freq, freq, indexing="ij")
    pos = 1.0 + torch.arange(padding - size, padding, device=kernel.device, dtype=kernel.dtype)
    p0, p1 = torch.meshgrid(pos, pos, indexing="ij")
    phase = (f0.reshape(-1, 1) * p0.reshape(1, -1) + f1.reshape(-1, 1) * p1.reshape(1, -1)).T
    re = torch.matmul(taps, torch.cos(phase))                    # [A, B, sample^2]
    im = torch.matmul(taps, torch.sin(phase))
    power = re.square().sum(1) + im.square().sum(1)              # [A, sample^2]
    bound = power.max(-1)[0].sum()
    return bound.sqrt() if sqrt else bound
