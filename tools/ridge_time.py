"""Time initialize_code(mode='ridge') (lasso_b200_ridge_init_f32) against the reference's formula in stock torch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200.linear import initialize_code
from lasso_b200.testing import make_problem
dev = torch.device("cuda", 0)
for (n, d, k) in [(65536, 64, 256), (10000, 289, 300)]:
    x, w = make_problem(n, d, k, seed=0, kind="randn")
    xd, wd = x.to(dev), w.to(dev)
    def torch_ridge():
        g = wd.T @ wd; g.diagonal().add_(0.5); c, info = torch.linalg.cholesky_ex(g); assert info == 0
        return torch.cholesky_solve(wd.T @ xd.T, c).T.contiguous()
    res = {}
    for name, fn in (("fused", lambda: initialize_code(xd, wd, 0.5, "ridge")), ("torch", torch_ridge)):
        for _ in range(3): out = fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): out = fn()
        torch.cuda.synchronize(); res[name] = ((time.perf_counter() - t0) * 50, out)
    err = float((res["fused"][1] - res["torch"][1]).norm() / res["torch"][1].norm())
    print("ridge init n=%d d=%d k=%d: fused %.3f ms, torch %.3f ms, rel diff %.2e" % (n, d, k, res["fused"][0], res["torch"][0], err))
