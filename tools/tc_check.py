"""GPU diagnostic: tcgen05 path vs FFMA path vs CPU oracle vs float64 gold, per iteration count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro, support_mismatch

dev = torch.device("cuda", 0)
cases = [(256, 64, 256, "planted", 0.1), (256, 64, 256, "randn", 0.1), (384, 16, 32, "planted", 0.1),
         (100, 20, 60, "planted", 0.1), (1000, 64, 128, "randn", 0.05), (130, 8, 24, "randn", 0.2)]
for n, d, k, kind, alpha in cases:
    x, w = make_problem(n, d, k, seed=0, kind=kind)
    lr = 1.0 / oracle.lipschitz_constant(w)
    xd, wd = x.to(dev), w.to(dev)
    for iters in (1, 2, 3, 10, 50, 200):
        try:
            ztc, _, _ = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="tcgen05")
            torch.cuda.synchronize()
        except Exception as e:
            print("TC FAILED", n, d, k, kind, iters, e); sys.exit(1)
        zff, _, _ = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="ffma")
        z32 = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
        z64 = torch.from_numpy(oracle.ista_f64(x.numpy(), torch.zeros(n, k).numpy(), w.numpy(), alpha, lr, iters))
        print("n=%4d d=%2d k=%3d %-7s it=%3d | tc~ffma %.2e tc~ref32 %.2e tc~f64 %.2e ffma~f64 %.2e ref32~f64 %.2e supp %.1e" % (
            n, d, k, kind, iters, rel_fro(ztc, zff), rel_fro(ztc, z32), rel_fro(ztc, z64), rel_fro(zff, z64),
            rel_fro(z32, z64), support_mismatch(ztc.cpu(), z32)))
# timing at C2
n, d, k = 65536, 64, 256
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
out = torch.empty(n, k, device=dev)
for path in ("tcgen05", "ffma"):
    for _ in range(2):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, 0.0, path=path, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, 0.0, path=path, out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("C2 %s: %.1f us/iter, %.0f it/s" % (path, dt / 200 * 1e6, 200 / dt))
