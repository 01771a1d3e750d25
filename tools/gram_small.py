"""A small solve on the Gram-form kernel and a small tensor-core Gram (for compute-sanitizer racecheck, which is slow)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro
dev = torch.device("cuda", 0)
x, w = make_problem(300, 200, 132, seed=1, kind="randn")
x, w = x.to(dev), w.to(dev)
lr = 1.0 / _cabi.lipschitz(w)
zg, it, _ = _cabi.fista_device(x, w, None, 0.2, lr, 6, True, 0.0, path="gram", want_iters=True)
zf, _, _ = _cabi.fista_device(x, w, None, 0.2, lr, 6, True, 0.0, path="ffma")
print("gram vs ffma", rel_fro(zg.cpu(), zf.cpu()), "iters", it)
