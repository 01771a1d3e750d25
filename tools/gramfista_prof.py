"""One solve at the notebook's shape on the Gram-form kernel (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem
dev = torch.device("cuda", 0)
n = int(os.environ.get("N", 10000))
x, w = make_problem(n, 289, 300, seed=5, kind="planted")
x, w = x.to(dev), w.to(dev)
lr = 1.0 / _cabi.lipschitz(w)
tol = float(os.environ.get("TOL", -1.0))
for _ in range(int(os.environ.get("REPS", 2))):
    _cabi.fista_device(x, w, None, 0.5, lr, 20, True, tol, path="gram")
torch.cuda.synchronize()
