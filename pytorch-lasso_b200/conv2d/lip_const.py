"""Upper bound of the Lipschitz constant of a stride-1 convolution (Araujo et al.), the quantity the
reference's ``lip_bound_conv2d`` returns (lasso/conv2d/lip_const.py:96-135): the maximum over a
frequency grid of the per-frequency Gram trace of the kernel's Fourier symbol.  Host-side helper on a
tiny tensor; kept in torch like ``linear/utils.ridge``."""
from __future__ import annotations

import math

import torch

__all__ = ["lip_bound_conv2d"]


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
