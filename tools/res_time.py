"""Time the resident kernel alone at config 2 (device-resident inputs), for variant builds
(LASSO_B200_LIB=<variant .so>).  Prints it/s; results of experiment builds are NOT checked."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem
n, d, k, iters = int(os.environ.get("N", 65536)), 64, 256, 200
tol = float(os.environ.get("TOL", -1.0))   # -1: no stop-test record, 0: "moved" flags, > 0: sums
x, w = make_problem(n, d, k, seed=0)
dev = torch.device("cuda", 0)
xd, wd = x.to(dev), w.to(dev)
lr = 1.0 / float(torch.linalg.eigvalsh((w.T @ w).double())[-1])
out = torch.empty(n, k, device=dev)
for _ in range(3):
    _cabi.fista_device(xd, wd, None, 0.1, lr, iters, True, tol, path="resident", out=out)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _cabi.fista_device(xd, wd, None, 0.1, lr, iters, True, tol, path="resident", out=out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("tol=%g %s: %.3f ms / %d it = %.0f it/s  fallbacks=%d" % (tol, os.path.basename(os.environ.get("LASSO_B200_LIB", "default")), best, iters, iters / best * 1e3, _cabi.resident_fallbacks()))
