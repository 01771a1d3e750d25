"""Convolutional ISTA / FISTA -- mirrors ``lasso.conv2d.ista.ista_conv2d`` (lasso/conv2d/ista.py:7-49).

Same signature, defaults and error behaviour.  The loop runs in ``liblasso_b200.so`` as
im2col -> linear on the k-blocked tcgen05 kernel (``lasso_b200_conv2d_fista_f32``):
``conv_transpose2d(z, W)`` is the overlap-add of ``Z W_lin^T`` and ``conv2d(r, W)`` is
``unfold(r) W_lin``, with the residual formed in image space between the two halves of an
iteration.  Codes keep the reference's layout ``[n, filters, oh, ow]`` at the boundary; inside
they are patch-major rows ``[n*oh*ow, filters]`` (two permutes per call).

``lr='exact'`` (an extension) takes the step from the exact Lipschitz constant of the convolutional
dictionary (power iteration on the device, any kernel size / stride / padding); ``lr='auto'`` keeps the
reference's Fourier bound and its restrictions (odd kernels, stride 1).

Built: any integer ``stride`` / ``padding`` (the same for both axes), ``cin*kh*kw <= 128``,
``filters <= 1024`` (multiples of 4), one image's patch matrix within 200 KB of shared memory.
Everything else raises -- there is no PyTorch fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi
from ..linear import utils as _utils
from .lip_const import _one_int, lip_bound_conv2d, lip_constant

__all__ = ["ista_conv2d"]


def ista_conv2d(x, z0, weight, alpha=1.0, stride=1, padding=0, fast=True,
                maxiter=10, lr='auto', tol=1e-5, verbose=False):
    stride, padding = _one_int(stride), _one_int(padding)      # (s, s) / (p, p) as torch's conv2d takes them
    if lr == 'exact':
        # Extension (the reference has no such option): the exact constant lambda_max(conv2d^T conv2d) for
        # this image size, computed on the device -- any kernel size / stride / padding, e.g. BASELINE
        # config 5's 8x8 filters, for which lr='auto' raises in the reference (lip_const.py:101-102).
        # The value is a Rayleigh quotient (never above the eigenvalue, ~1e-5 below at worst when the top
        # of the spectrum is a cluster), hence the 1e-4 margin on the step.
        lr = 1 / (1.0001 * lip_constant(weight, x.shape[-2:], stride=stride, padding=padding))
    if lr == 'auto':
        if stride != 1:
            raise NotImplementedError("auto lr is only implemented for "
                                      "stride == 1.")          # ista.py:10-12
        lr = 1 / float(lip_bound_conv2d(weight, padding))      # ista.py:13-15 (odd kernels only)
    if not (isinstance(stride, int) and isinstance(padding, int)) or stride < 1 or padding < 0:
        raise NotImplementedError("lasso_b200.conv2d takes one integer stride >= 1 and padding >= 0 for both axes")
    for name, t in (("x", x), ("z0", z0), ("weight", weight)):
        if t.dtype != torch.float32:
            raise NotImplementedError("lasso_b200 computes in float32 only; {} has dtype {}".format(name, t.dtype))
        if t.requires_grad:
            raise NotImplementedError("lasso_b200 does not record an autograd graph; detach {} first".format(name))
    if maxiter == 0:
        return z0
    filters, cin, kh, kw = weight.shape
    n, cx, h, w = x.shape
    if (h + 2 * padding - kh) % stride or (w + 2 * padding - kw) % stride:
        # conv_transpose2d(z) would come out smaller than x: the reference fails on `x_hat - x` (ista.py:18-19)
        raise RuntimeError("image size + 2*padding - kernel size must be a multiple of stride={}".format(stride))
    oh, ow = (h + 2 * padding - kh) // stride + 1, (w + 2 * padding - kw) // stride + 1
    if cx != cin or tuple(z0.shape) != (n, filters, oh, ow):
        raise ValueError("expected x[n,{},h,w] and z0[n,{},{},{}]; got {} and {}".format(
            cin, filters, oh, ow, tuple(x.shape), tuple(z0.shape)))
    dev = x.device if x.is_cuda else _utils.default_device()
    tol_abs = float(np.float32(z0.numel() * tol))              # ista.py:16
    # weight_lin[c*kh*kw + a*kw + b, f] = W[f, c, a, b];  codes as patch-major rows
    w_lin = weight.to(dev).reshape(filters, cin * kh * kw).T.contiguous()
    z_rows = z0.to(dev).permute(0, 2, 3, 1).reshape(n * oh * ow, filters).contiguous()
    if not bool(z_rows.any()):
        z_rows = None                                          # zero start: the kernel clears its own buffer
    xd = x.to(dev).contiguous()
    out_rows, done = _cabi.conv2d_fista_device(xd, w_lin, z_rows, kh, kw, alpha, float(lr), maxiter, fast, tol_abs,
                                               want_iters=verbose, stride=stride, padding=padding)
    z = out_rows.reshape(n, oh, ow, filters).permute(0, 3, 1, 2).contiguous()
    if verbose:
        # 'loss: %0.4f' of the iterate each executed iteration starts from (ista.py:21-24, 36-38).  The kernel keeps
        # all iterations on the device, so the i-th iterate is produced by a run of i iterations (deterministic
        # kernels: the same iterate the long run passed through) -- a debugging aid, quadratic in maxiter.
        import torch.nn.functional as F
        wd = weight.to(dev)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            for i in range(done):
                if i == 0:
                    zi = z0.to(dev)
                else:
                    rows, _ = _cabi.conv2d_fista_device(xd, w_lin, z_rows, kh, kw, alpha, float(lr), i, fast, -1.0,
                                                        stride=stride, padding=padding)
                    zi = rows.reshape(n, oh, ow, filters).permute(0, 3, 1, 2)
                x_hat = F.conv_transpose2d(zi, wd, stride=stride, padding=padding)
                loss = (0.5 * (xd - x_hat).pow(2).sum() + alpha * zi.abs().sum()) / n
                print('loss: %0.4f' % float(loss))
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
    return z if x.is_cuda else z.cpu()
