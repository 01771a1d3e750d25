"""GPU diagnostic: which batch sizes / options make the tcgen05 path fail (run with LASSO_B200_DEBUG=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro
dev = torch.device("cuda", 0)
d, k = 64, 256
w = lasso_b200.testing.make_dictionary(d, k)
lr = 1.0 / oracle.lipschitz_constant(w)
wd = w.to(dev)
for n, iters, tol in [(128 * 8, 5, -1.0), (128 * 149, 1, -1.0), (128 * 149, 5, -1.0), (128 * 296, 5, -1.0),
                      (128 * 300, 5, -1.0), (65536, 3, -1.0), (65536, 20, -1.0), (65536, 20, 0.0), (65536, 200, 0.0), (65536, 200, 0.0), (65536, 200, 0.0), (65536, 200, -1.0), (65536, 200, -1.0), (40000, 200, 0.0), (128*148*2+77, 199, 0.0)]:
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(1)).to(dev)
    try:
        ztc, _, _ = _cabi.fista_device(x, wd, None, 0.1, lr, iters, True, tol, path="tcgen05")
        torch.cuda.synchronize()
        zff, _, _ = _cabi.fista_device(x, wd, None, 0.1, lr, iters, True, tol, path="ffma")
        torch.cuda.synchronize()
        print("n=%6d iters=%3d tol=%4.1f OK  tc~ffma %.2e" % (n, iters, tol, rel_fro(ztc, zff)), flush=True)
    except Exception as e:
        print("n=%6d iters=%3d tol=%4.1f FAILED: %s" % (n, iters, tol, e), flush=True)
        break
