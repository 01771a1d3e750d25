import os
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem
dev = torch.device("cuda", 0)
n, d, k = 65536, 64, 256
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
out = torch.empty(n, k, device=dev)
for name, tol_abs in (("mode0 (tol<0)", -1.0), ("mode2 (tol=0)", 0.0), ("mode1 (tol=1e-12)", 1e-12 * n * k)):
    for _ in range(2):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, tol_abs, path="resident", out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, tol_abs, path="resident", out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("%s: %.2f us/iter, %.0f it/s" % (name, dt / 200 * 1e6, 200 / dt))
