"""Seeded synthetic lasso problems.

Generated with a CPU ``torch.Generator`` so that CPU and GPU runs, the golden
fixtures and the benchmark all see identical bits.  No oracle code here.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["make_dictionary", "make_problem", "rel_fro", "support_mismatch"]


def make_dictionary(d: int, k: int, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """Unit-norm random atoms as columns, the shape dict_learning produces (dict_learning.py:31)."""
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(d, k, generator=g, dtype=dtype), dim=0)


def make_problem(n: int, d: int, k: int, seed: int = 0, kind: str = "planted",
                 density: float = 0.05, noise: float = 0.01, dtype=torch.float32):
    """Return ``(x[n,d], weight[d,k])``.

    kind='planted': x = (randn(n,k) * Bernoulli(density)) @ W^T + noise * randn(n,d)
    kind='randn'  : x = randn(n,d) (stress case, denser codes)
    """
    weight = make_dictionary(d, k, seed, dtype)
    g = torch.Generator().manual_seed(seed + 1000003)
    if kind == "planted":
        code = torch.randn(n, k, generator=g, dtype=dtype)
        code = code * (torch.rand(n, k, generator=g, dtype=dtype) < density)
        x = code @ weight.T + noise * torch.randn(n, d, generator=g, dtype=dtype)
    elif kind == "randn":
        x = torch.randn(n, d, generator=g, dtype=dtype)
    else:
        raise ValueError("unknown problem kind '{}'".format(kind))
    return x.contiguous(), weight.contiguous()


def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_F / ||b||_F in float64 (0 when both are all-zero)."""
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    den = float(b64.norm())
    num = float((a64 - b64).norm())
    return num / den if den > 0 else num


def support_mismatch(a: torch.Tensor, b: torch.Tensor) -> float:
    """Fraction of entries whose zero / non-zero status differs."""
    return float(((a != 0) != (b != 0)).double().mean())
