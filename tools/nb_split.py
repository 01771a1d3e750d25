import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lasso_b200
from lasso_b200 import _cabi
from lasso_b200.linear import sparse_encode, lasso_loss, update_dict, update_dict_ridge, initialize_code
from lasso_b200.linear.solvers import lipschitz_constant
from lasso_b200.testing import make_problem
dev = torch.device("cuda", 0)
n, d, k = 10000, 289, 300
x, w = make_problem(n, d, k, seed=0)
x, w = (x * 3).to(dev), w.to(dev).clone()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, out
ms, L = timed(lambda: lipschitz_constant(w)); print("lipschitz %.2f ms" % ms)
ms, z0 = timed(lambda: initialize_code(x, w, 0.5, 'ridge')); print("ridge init %.2f ms" % ms)
ms, z = timed(lambda: sparse_encode(x, w, 0.5, z0=z0, maxiter=20, lr=1.0 / L)); print("E-step 20 it (lr pinned) %.2f ms" % ms)
ms, _ = timed(lambda: lasso_loss(x, z, w, 0.5)); print("loss %.2f ms" % ms)
ms, (gzz, gzx) = timed(lambda: _cabi.gram(z, x)); print("gram %.2f ms" % ms)
ms, _ = timed(lambda: _cabi.dict_update_gram(w, gzz.clone(), gzx.clone())); print("sweep %.2f ms" % ms)
ms, _ = timed(lambda: update_dict(w, x, z)); print("update_dict %.2f ms" % ms)
ms, _ = timed(lambda: update_dict_ridge(x, z, 2e-2)); print("update_dict_ridge %.2f ms" % ms)
# ridge-init variants (k x k system, n right-hand sides)
def ridge_solve():
    gram = w.T @ w; gram.diagonal().add_(0.5)
    chol, info = torch.linalg.cholesky_ex(gram)
    assert info == 0
    return torch.cholesky_solve(w.T @ x.T, chol).T.contiguous()
def ridge_inverse():
    gram = w.T @ w; gram.diagonal().add_(0.5)
    chol, info = torch.linalg.cholesky_ex(gram)
    assert info == 0
    return (x @ w) @ torch.cholesky_inverse(chol)
def ridge_inverse_nosync():
    gram = w.T @ w; gram.diagonal().add_(0.5)
    chol = torch.linalg.cholesky(gram)
    return (x @ w) @ torch.cholesky_inverse(chol)
def ridge_f64inv():
    gram = (w.double().T @ w.double()); gram.diagonal().add_(0.5)
    return (x @ w) @ torch.linalg.inv(gram).float()
for name, fn in (("solve", ridge_solve), ("inverse", ridge_inverse), ("inverse_nosync", ridge_inverse_nosync), ("f64 inv", ridge_f64inv)):
    ms, zz = timed(fn); print("ridge %-15s %.2f ms   vs solve %.2e" % (name, ms, float((zz - ridge_solve()).norm() / ridge_solve().norm())))
