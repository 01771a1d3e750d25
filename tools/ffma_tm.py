"""E-step of the notebook shape (n=10000, d=289, k=300, 20 FISTA iterations) on the FFMA path for each
row-tile height (LASSO_B200_FFMA_TM is read at every launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lasso_b200
from lasso_b200.linear import sparse_encode
from lasso_b200.testing import make_problem
dev = torch.device("cuda", 0)
shapes = ((10000, 289, 300), (65536, 289, 300), (10000, 200, 512))
if os.environ.get("FFMA_ONE"):   # one shape, default tile (for an ncu capture)
    shapes = ((65536, 289, 300),)
for n, d, k in shapes:
    x, w = make_problem(n, d, k, seed=0)
    x, w = x.to(dev), w.to(dev)
    for tm in (("64", "32", "16") if not os.environ.get("FFMA_ONE") else ("64",)):
        os.environ["LASSO_B200_FFMA_TM"] = tm
        f = lambda: sparse_encode(x, w, 0.5, maxiter=20, lr=0.01)
        f(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): f()
        torch.cuda.synchronize()
        print("n=%d d=%d k=%d TM=%s  %.2f ms / 20 it" % (n, d, k, tm, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
