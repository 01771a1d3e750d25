"""GPU timing split of one EM step of dict_learning at C4-per-GPU scale."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.linear import sparse_encode, lasso_loss, update_dict, update_dict_ridge
from lasso_b200.testing import make_problem

dev = torch.device("cuda", 0)
n, d, k = int(os.environ.get("N", 131072)), 64, 256
x, w = make_problem(n, d, k, seed=0)
x, w = x.to(dev), w.to(dev).clone()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out

ms, z = timed(lambda: sparse_encode(x, w, 0.1, maxiter=100, tol=0.0))
print("E-step  sparse_encode 100 it (incl. lr=auto): %8.2f ms" % ms)
ms, _ = timed(lambda: lasso_b200.linear.solvers.lipschitz_constant(w))
print("        lipschitz_constant:                  %8.2f ms" % ms)
ms, _ = timed(lambda: lasso_loss(x, z, w, 0.1))
print("loss                                         %8.2f ms" % ms)
ms, (gzz, gzx) = timed(lambda: _cabi.gram(z, x))
print("M-step  gram statistics:                     %8.2f ms" % ms)
ms, _ = timed(lambda: _cabi.dict_update_gram(w, gzz.clone(), gzx.clone()))
print("        atom sweep (one CTA):                %8.2f ms" % ms)
ms, _ = timed(lambda: update_dict(w, x, z))
print("        update_dict total:                   %8.2f ms" % ms)
ms, _ = timed(lambda: update_dict_ridge(x, z, 1e-2))
print("        update_dict_ridge total:             %8.2f ms" % ms)
print("nnz(z) = %.3f" % float((z != 0).float().mean()))
