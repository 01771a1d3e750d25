"""Time the atom sweep (K4): blocked (sweep_blk.cu) against the atom-by-atom kernels (LASSO_B200_SWEEP=legacy)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.linear import sparse_encode
from lasso_b200.testing import make_problem, rel_fro
dev = torch.device("cuda", 0)
for n, d, k in [(131072, 64, 256), (10000, 289, 300), (20000, 128, 1024)]:
    x, w = make_problem(n, d, k, seed=1)
    x, w = x.to(dev), w.to(dev)
    z = sparse_encode(x, w, alpha=0.1, maxiter=20, tol=0.0)
    gzz, gzx = _cabi.gram(z, x)
    res = {}
    for mode in ("blocked", "legacy"):
        if mode == "legacy": os.environ["LASSO_B200_SWEEP"] = "legacy"
        else: os.environ.pop("LASSO_B200_SWEEP", None)
        for _ in range(3):
            wm, a, b = w.clone(), gzz.clone(), gzx.clone()
            _cabi.dict_update_gram(wm, a, b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            wm, a, b = w.clone(), gzz.clone(), gzx.clone()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _cabi.dict_update_gram(wm, a, b)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res[mode] = (wm, sorted(ts)[len(ts) // 2])
    os.environ.pop("LASSO_B200_SWEEP", None)
    print("sweep d=%d k=%d: blocked %.3f ms, atom-by-atom %.3f ms, difference %.2e" % (
        d, k, res["blocked"][1] * 1e3, res["legacy"][1] * 1e3, rel_fro(res["blocked"][0].cpu(), res["legacy"][0].cpu())))
