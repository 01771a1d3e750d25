"""Small dense helpers on the sparse_encode boundary (mirrors lasso/linear/utils.py).

``init='ridge'`` (sparse_encode.py:28-29) runs in the library (``lasso_b200_ridge_init_f32``); the
generic ``ridge`` / ``lstsq`` helpers below mirror lasso/linear/utils.py for callers that used them
directly and stay in torch.
"""
from __future__ import annotations

import torch

__all__ = ["ridge", "lstsq", "default_device"]


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("lasso_b200 needs a CUDA device (sm_100a); none is visible. "
                           "There is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def ridge(b: torch.Tensor, A: torch.Tensor, alpha: float = 1e-4) -> torch.Tensor:
    """argmin_x |A x - b|^2 + alpha |x|^2 via Cholesky of the regularised Gram (utils.py:28-40)."""
    gram = A.T @ A
    gram.diagonal().add_(alpha)
    chol, info = torch.linalg.cholesky_ex(gram)
    if info != 0:
        raise RuntimeError("The Gram matrix is not positive definite. "
                           "Try increasing 'alpha'.")
    return torch.cholesky_solve(A.T @ b, chol)


def lstsq(b: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """Least-squares / least-norm solution of A x = b (utils.py:13-25)."""
    return torch.linalg.lstsq(A, b).solution
