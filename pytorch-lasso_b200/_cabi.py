"""ctypes binding of ``csrc/liblasso_b200.so`` (the C ABI in ``include/lasso_b200.h``).

The library is the product: there is no PyTorch or CPU fallback behind it.  If
the shared object is missing or a call fails, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LASSO_B200_LIB: load another build of the library (kernel variants, tools/variants.sh)
LIB_PATH = os.environ.get("LASSO_B200_LIB") or os.path.join(_HERE, "csrc", "liblasso_b200.so")

PATH_AUTO, PATH_FFMA, PATH_TCGEN05, PATH_RESIDENT, PATH_BLOCKED, PATH_GRAM = 0, 1, 2, 3, 4, 5
_PATH_NAMES = {"auto": PATH_AUTO, "ffma": PATH_FFMA, "tcgen05": PATH_TCGEN05,
               "resident": PATH_RESIDENT, "blocked": PATH_BLOCKED, "gram": PATH_GRAM}

EXPORTS = (
    "lasso_b200_version",
    "lasso_b200_last_error",
    "lasso_b200_select_path",
    "lasso_b200_launch_count",
    "lasso_b200_resident_fallbacks",
    "lasso_b200_fista_f32",
    "lasso_b200_fista_f32_host",
    "lasso_b200_conv2d_fista_f32",
    "lasso_b200_lipschitz_f32",
    "lasso_b200_conv2d_lipschitz_f32",
    "lasso_b200_ridge_init_f32",
    "lasso_b200_matmul_f32",
    "lasso_b200_loss_terms_f32",
    "lasso_b200_gram_f32",
    "lasso_b200_dict_update_gram_f32",
    "lasso_b200_zero_columns_f32",
    "lasso_b200_gradient_f32",
    "lasso_b200_linesearch_trial_f32",
    "lasso_b200_momentum_f32",
    "lasso_b200_release_workspace",
)


class LassoB200Error(RuntimeError):
    """A call into liblasso_b200.so returned a negative status."""


_lib = None
_lock = threading.Lock()


def _declare(lib):
    c = ctypes
    vp, i32, i64, f64 = c.c_void_p, c.c_int32, c.c_int64, c.c_double
    lib.lasso_b200_version.restype = i32
    lib.lasso_b200_version.argtypes = []
    lib.lasso_b200_last_error.restype = c.c_char_p
    lib.lasso_b200_last_error.argtypes = []
    lib.lasso_b200_select_path.restype = i32
    lib.lasso_b200_select_path.argtypes = [i64, i32, i32]
    lib.lasso_b200_launch_count.restype = i64
    lib.lasso_b200_launch_count.argtypes = []
    lib.lasso_b200_resident_fallbacks.restype = i64
    lib.lasso_b200_resident_fallbacks.argtypes = []
    lib.lasso_b200_fista_f32.restype = i32
    lib.lasso_b200_fista_f32.argtypes = [vp, vp, vp, vp, i64, i32, i32, f64, f64, i32, i32,
                                         f64, c.POINTER(i32), vp, i32, vp]
    lib.lasso_b200_fista_f32_host.restype = i32
    lib.lasso_b200_fista_f32_host.argtypes = [vp, vp, vp, vp, i64, i32, i32, f64, f64, i32,
                                              i32, f64, c.POINTER(i32), vp, i32, vp]
    lib.lasso_b200_conv2d_fista_f32.restype = i32
    lib.lasso_b200_conv2d_fista_f32.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, i32,
                                                f64, f64, i32, i32, f64, c.POINTER(i32), vp, vp]
    lib.lasso_b200_lipschitz_f32.restype = i32
    lib.lasso_b200_lipschitz_f32.argtypes = [vp, i32, i32, i32, c.POINTER(f64), vp]
    lib.lasso_b200_conv2d_lipschitz_f32.restype = i32
    lib.lasso_b200_conv2d_lipschitz_f32.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                                    c.POINTER(f64), vp]
    lib.lasso_b200_ridge_init_f32.restype = i32
    lib.lasso_b200_ridge_init_f32.argtypes = [vp, vp, i64, i32, i32, f64, vp, c.POINTER(i32), vp]
    lib.lasso_b200_matmul_f32.restype = i32
    lib.lasso_b200_matmul_f32.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    lib.lasso_b200_loss_terms_f32.restype = i32
    lib.lasso_b200_loss_terms_f32.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    lib.lasso_b200_gram_f32.restype = i32
    lib.lasso_b200_gram_f32.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    lib.lasso_b200_dict_update_gram_f32.restype = i32
    lib.lasso_b200_dict_update_gram_f32.argtypes = [vp, vp, vp, i32, i32, f64, vp, vp, i32, vp]
    lib.lasso_b200_zero_columns_f32.restype = i32
    lib.lasso_b200_zero_columns_f32.argtypes = [vp, i64, i32, vp, vp]
    lib.lasso_b200_gradient_f32.restype = i32
    lib.lasso_b200_gradient_f32.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.lasso_b200_linesearch_trial_f32.restype = i32
    lib.lasso_b200_linesearch_trial_f32.argtypes = [vp, vp, vp, vp, i64, i32, i32, f64, f64, vp, vp, vp]
    lib.lasso_b200_momentum_f32.restype = i32
    lib.lasso_b200_momentum_f32.argtypes = [vp, vp, f64, vp, i64, vp, vp]
    lib.lasso_b200_release_workspace.restype = i32
    lib.lasso_b200_release_workspace.argtypes = []


def load():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise LassoB200Error(
                        "CUDA extension not built: {} is missing. Run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` in the repo "
                        "root (nvcc, sm_100a). There is no CPU fallback.".format(LIB_PATH))
                lib = ctypes.CDLL(LIB_PATH)
                _declare(lib)
                _lib = lib
    return _lib


def _check(status: int):
    if status != 0:
        msg = load().lasso_b200_last_error().decode("utf-8", "replace")
        raise LassoB200Error("liblasso_b200 status {}: {}".format(status, msg))


def path_code(path) -> int:
    if isinstance(path, int):
        return path
    try:
        return _PATH_NAMES[path]
    except KeyError:
        raise ValueError("unknown kernel path '{}'".format(path)) from None


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise LassoB200Error("{} must be a CUDA tensor".format(name))
    if t.dtype != torch.float32:
        raise LassoB200Error("{} must be float32, got {}".format(name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def launch_count() -> int:
    return int(load().lasso_b200_launch_count())


def resident_fallbacks() -> int:
    return int(load().lasso_b200_resident_fallbacks())


def select_path(n: int, d: int, k: int) -> int:
    return int(load().lasso_b200_select_path(n, d, k))


def _check_out(out, n, k, device):
    """A caller-provided result buffer is written through its raw pointer: it has to be exactly the
    [n,k] float32 contiguous tensor on the right device."""
    if tuple(out.shape) != (n, k) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 tensor of shape ({}, {})".format(n, k))
    if out.device.type != device.type or (device.type == "cuda" and out.device != device):
        raise ValueError("out must live on {}".format(device))
    return out


def fista_device(x, weight, z0, alpha, lr, maxiter, fast, tol_abs, path=PATH_AUTO,
                 want_iters=False, want_hist=False, out=None):
    """Run the FISTA loop on device tensors; returns (z, iters_done|None, hist|None)."""
    lib = load()
    x = _dev_f32(x, "x")
    weight = _dev_f32(weight, "weight")
    n, d = x.shape
    k = weight.shape[1]
    if weight.shape[0] != d:
        raise LassoB200Error("weight must be [d,k] with d == x.shape[1]")
    if z0 is not None:
        z0 = _dev_f32(z0, "z0")
    z = _check_out(out, n, k, x.device) if out is not None else \
        torch.empty((n, k), dtype=torch.float32, device=x.device)
    hist = torch.zeros(max(maxiter, 1), dtype=torch.float64, device=x.device) if want_hist else None
    iters = ctypes.c_int32(0)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_fista_f32(
            x.data_ptr(), weight.data_ptr(), z0.data_ptr() if z0 is not None else None,
            z.data_ptr(), n, d, k, float(alpha), float(lr), int(maxiter), int(bool(fast)),
            float(tol_abs), ctypes.byref(iters) if want_iters else None,
            hist.data_ptr() if want_hist else None, path_code(path), _stream_ptr(x.device)))
    return z, (iters.value if want_iters else None), (hist[:maxiter] if want_hist else None)


def conv2d_fista_device(x, weight_lin, z0_rows, kh, kw, alpha, lr, maxiter, fast, tol_abs, want_iters=False,
                        stride=1, padding=0):
    """Convolutional FISTA on device tensors: x [n,cin,h,w], weight_lin [cin*kh*kw, k], codes as rows
    [n*oh*ow, k] with oh = (h + 2*padding - kh) // stride + 1.  Returns (z_rows, iters_done|None)."""
    lib = load()
    x = _dev_f32(x, "x")
    weight_lin = _dev_f32(weight_lin, "weight_lin")
    n_img, cin, h, w = x.shape
    k = weight_lin.shape[1]
    rows = n_img * ((h + 2 * padding - kh) // stride + 1) * ((w + 2 * padding - kw) // stride + 1)
    if z0_rows is not None:
        z0_rows = _dev_f32(z0_rows, "z0")
    z = torch.empty((rows, k), dtype=torch.float32, device=x.device)
    iters = ctypes.c_int32(0)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_conv2d_fista_f32(
            x.data_ptr(), weight_lin.data_ptr(), z0_rows.data_ptr() if z0_rows is not None else None,
            z.data_ptr(), n_img, cin, h, w, int(kh), int(kw), int(stride), int(padding), k, float(alpha),
            float(lr), int(maxiter),
            int(bool(fast)), float(tol_abs), ctypes.byref(iters) if want_iters else None, None,
            _stream_ptr(x.device)))
    return z, (iters.value if want_iters else None)


def fista_host(x, weight, z0, alpha, lr, maxiter, fast, tol_abs, path=PATH_AUTO,
               want_iters=False, out=None):
    """Same solve on HOST (CPU) tensors through the C ABI's host entry point."""
    lib = load()
    for name, t in (("x", x), ("weight", weight)):
        if t.is_cuda or t.dtype != torch.float32:
            raise LassoB200Error("{} must be a float32 CPU tensor".format(name))
    x = x.contiguous()
    weight = weight.contiguous()
    n, d = x.shape
    k = weight.shape[1]
    if z0 is not None:
        z0 = z0.contiguous()
    z = _check_out(out, n, k, torch.device("cpu")) if out is not None else \
        torch.empty((n, k), dtype=torch.float32)
    iters = ctypes.c_int32(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    _check(lib.lasso_b200_fista_f32_host(
        x.data_ptr(), weight.data_ptr(), z0.data_ptr() if z0 is not None else None,
        z.data_ptr(), n, d, k, float(alpha), float(lr), int(maxiter), int(bool(fast)),
        float(tol_abs), ctypes.byref(iters) if want_iters else None, None, path_code(path),
        _stream_ptr(dev)))
    return z, (iters.value if want_iters else None)


def lipschitz(weight, iters=2000) -> float:
    lib = load()
    weight = _dev_f32(weight, "weight")
    d, k = weight.shape
    out = ctypes.c_double(0.0)
    with torch.cuda.device(weight.device):
        _check(lib.lasso_b200_lipschitz_f32(weight.data_ptr(), d, k, int(iters),
                                            ctypes.byref(out), _stream_ptr(weight.device)))
    return out.value


def _f64_out(t, numel, name):
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == numel):
        raise ValueError("{} must be a contiguous float64 CUDA tensor of {} elements".format(name, numel))
    return t


def conv2d_lipschitz(weight, imsize, stride=1, padding=0, iters=2000) -> float:
    """lambda_max of conv2d^T conv2d on [cin, h, w] images (dense operator + the dictionary's eigen-solver)."""
    lib = load()
    weight = _dev_f32(weight, "weight")
    filters, cin, kh, kw = weight.shape
    h, w = imsize
    out = ctypes.c_double(0.0)
    with torch.cuda.device(weight.device):
        _check(lib.lasso_b200_conv2d_lipschitz_f32(
            weight.data_ptr(), filters, cin, kh, kw, int(h), int(w), int(stride), int(padding), int(iters),
            ctypes.byref(out), _stream_ptr(weight.device)))
    return out.value


def ridge_init(x, weight, alpha):
    """z0[n,k] = ((W^T W + alpha I)^-1 W^T x^T)^T on the device; RuntimeError if the Gram is not PD."""
    lib = load()
    x, weight = _dev_f32(x, "x"), _dev_f32(weight, "weight")
    n, d = x.shape
    k = weight.shape[1]
    z = torch.empty((n, k), dtype=torch.float32, device=x.device)
    flag = ctypes.c_int32(0)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_ridge_init_f32(x.data_ptr(), weight.data_ptr(), n, d, k, float(alpha),
                                             z.data_ptr(), ctypes.byref(flag), _stream_ptr(x.device)))
    if flag.value:
        raise RuntimeError("The Gram matrix is not positive definite. "
                           "Try increasing 'alpha'.")          # utils.py:36-38
    return z


def matmul(x, t):
    """x[n,d] @ t[d,k] in float32 on the library's FFMA kernel."""
    lib = load()
    x, t = _dev_f32(x, "x"), _dev_f32(t, "t")
    n, d = x.shape
    k = t.shape[1]
    if t.shape[0] != d:
        raise LassoB200Error("t must be [d,k] with d == x.shape[1]")
    z = torch.empty((n, k), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_matmul_f32(x.data_ptr(), t.data_ptr(), n, d, k, z.data_ptr(), _stream_ptr(x.device)))
    return z


def loss_terms(x, z, weight, out=None) -> torch.Tensor:
    """Device tensor [2] float64: (sum (x - z W^T)^2, sum |z|)."""
    lib = load()
    x, z, weight = _dev_f32(x, "x"), _dev_f32(z, "z"), _dev_f32(weight, "weight")
    n, d = x.shape
    k = weight.shape[1]
    out = _f64_out(out, 2, "out") if out is not None else torch.empty(2, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_loss_terms_f32(x.data_ptr(), z.data_ptr(), weight.data_ptr(),
                                             n, d, k, out.data_ptr(), _stream_ptr(x.device)))
    return out


def gram(z, x, out_zz=None, out_zx=None):
    """float64 device tensors (Z^T Z [k,k], Z^T X [k,d])."""
    lib = load()
    z, x = _dev_f32(z, "z"), _dev_f32(x, "x")
    n, k = z.shape
    d = x.shape[1]
    gzz = _f64_out(out_zz, k * k, "out_zz") if out_zz is not None else \
        torch.empty((k, k), dtype=torch.float64, device=z.device)
    gzx = _f64_out(out_zx, k * d, "out_zx") if out_zx is not None else \
        torch.empty((k, d), dtype=torch.float64, device=z.device)
    with torch.cuda.device(z.device):
        _check(lib.lasso_b200_gram_f32(z.data_ptr(), x.data_ptr(), n, d, k, gzz.data_ptr(),
                                       gzx.data_ptr(), _stream_ptr(z.device)))
    return gzz, gzx


def dict_update_gram(dictionary, gzz, gzx, eps=1e-10, redraw=None, positive=False):
    """In-place Gram-space atom sweep; returns int32 device mask of re-drawn atoms."""
    lib = load()
    if not (dictionary.is_cuda and dictionary.dtype == torch.float32 and dictionary.is_contiguous()):
        raise LassoB200Error("dictionary must be a contiguous float32 CUDA tensor")
    d, k = dictionary.shape
    zeroed = torch.zeros(k, dtype=torch.int32, device=dictionary.device)
    if redraw is not None:
        redraw = _dev_f32(redraw, "redraw")
    with torch.cuda.device(dictionary.device):
        _check(lib.lasso_b200_dict_update_gram_f32(
            dictionary.data_ptr(), gzz.data_ptr(), gzx.data_ptr(), d, k, float(eps),
            redraw.data_ptr() if redraw is not None else None, zeroed.data_ptr(),
            1 if positive else 0, _stream_ptr(dictionary.device)))
    return zeroed


def zero_columns(z, mask):
    """z[:, j] = 0 in place for every j with mask[j] != 0 (mask: int32 device tensor [k])."""
    lib = load()
    if not (z.is_cuda and z.dtype == torch.float32 and z.is_contiguous()):
        raise LassoB200Error("z must be a contiguous float32 CUDA tensor")
    n, k = z.shape
    if not (mask.is_cuda and mask.dtype == torch.int32 and mask.numel() == k):
        raise LassoB200Error("mask must be an int32 CUDA tensor of k entries")
    with torch.cuda.device(z.device):
        _check(lib.lasso_b200_zero_columns_f32(z.data_ptr(), n, k, mask.data_ptr(), _stream_ptr(z.device)))
    return z


def gradient(x, point, weight, grad_out=None):
    """(grad[n,k], f_sum device [1] float64) at ``point`` (ista.py:22-24)."""
    lib = load()
    x, point, weight = _dev_f32(x, "x"), _dev_f32(point, "point"), _dev_f32(weight, "weight")
    n, d = x.shape
    k = weight.shape[1]
    grad = grad_out if grad_out is not None else torch.empty_like(point)
    f_sum = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_gradient_f32(x.data_ptr(), point.data_ptr(), weight.data_ptr(), n, d, k,
                                           grad.data_ptr(), f_sum.data_ptr(), _stream_ptr(x.device)))
    return grad, f_sum


def linesearch_trial(x, point, grad, weight, step, alpha, cand_out=None):
    """(cand[n,k], sums device [4] float64) for one trial step (ista.py:26-40)."""
    lib = load()
    x, point, grad, weight = (_dev_f32(x, "x"), _dev_f32(point, "point"), _dev_f32(grad, "grad"),
                              _dev_f32(weight, "weight"))
    n, d = x.shape
    k = weight.shape[1]
    cand = cand_out if cand_out is not None else torch.empty_like(point)
    sums = torch.empty(4, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib.lasso_b200_linesearch_trial_f32(
            x.data_ptr(), point.data_ptr(), grad.data_ptr(), weight.data_ptr(), n, d, k,
            float(step), float(alpha), cand.data_ptr(), sums.data_ptr(), _stream_ptr(x.device)))
    return cand, sums


def momentum(z_next, z, beta, want_y=True):
    """(y = z_next + beta (z_next - z) or None, delta device [1] float64 = sum |z - z_next|)."""
    lib = load()
    z_next, z = _dev_f32(z_next, "z_next"), _dev_f32(z, "z")
    y = torch.empty_like(z_next) if want_y else None
    delta = torch.empty(1, dtype=torch.float64, device=z.device)
    with torch.cuda.device(z.device):
        _check(lib.lasso_b200_momentum_f32(z_next.data_ptr(), z.data_ptr(), float(beta),
                                           y.data_ptr() if want_y else None, z.numel(),
                                           delta.data_ptr(), _stream_ptr(z.device)))
    return y, delta


def release_workspace():
    _check(load().lasso_b200_release_workspace())
