"""Gram-form tcgen05 kernel (fista_gram.cu) against the FFMA kernel, the float64 solution and the
reference fixture at the notebook's shape; timing of both paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro

dev = torch.device("cuda", 0)


def f64_solution(x, w, z0, alpha, lr, maxiter, fast=True):
    x, w, z = x.double(), w.double(), z0.double()
    lr = float(np.float32(lr)); lam = float(np.float32(alpha * lr))
    y, t = z.clone(), 1.0
    for _ in range(maxiter):
        g = (y @ w.T - x) @ w
        zn = torch.nn.functional.softshrink(y - lr * g, lam)
        if fast:
            tn = (1 + (1 + 4 * t * t) ** 0.5) / 2
            y = zn + float(np.float32((t - 1) / tn)) * (zn - z); t = tn
        else:
            y = zn
        z = zn
    return z


from conftest import load_golden
g = load_golden("r2_notebook_289x300")
x, w, z0 = g["x"].to(dev), g["weight"].to(dev), g["z0"].to(dev)
for path in ("ffma", "gram"):
    z, it, _ = _cabi.fista_device(x, w, z0, g["alpha"], g["lr"], int(g["maxiter"]), bool(g["fast"]),
                                  g["z0"].numel() * g["tol"], path=path, want_iters=True)
    print("fixture r2_notebook_289x300 path=%s: rel %.3e iters %s fallbacks %d" % (
        path, rel_fro(z.cpu(), g["z"]), it, _cabi.load().lasso_b200_resident_fallbacks()))

for (n, d, k, alpha, maxiter, kind) in [(10000, 289, 300, 0.5, 20, "planted"), (10000, 289, 300, 0.1, 100, "planted"),
                                        (4099, 200, 132, 0.2, 50, "randn"), (777, 512, 320, 0.1, 60, "planted")]:
    x, w = make_problem(n, d, k, seed=5, kind=kind)
    x, w = x.to(dev), w.to(dev)
    lr = 1.0 / _cabi.lipschitz(w)
    z0 = torch.zeros(n, k, device=dev)
    want = f64_solution(x, w, z0, alpha, lr, maxiter)
    res = {}
    for path in ("ffma", "gram"):
        z, _, hist = _cabi.fista_device(x, w, None, alpha, lr, maxiter, True, -1.0, path=path, want_hist=True)
        res[path] = (z, hist)
        for _ in range(2):
            _cabi.fista_device(x, w, None, alpha, lr, maxiter, True, -1.0, path=path)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            _cabi.fista_device(x, w, None, alpha, lr, maxiter, True, -1.0, path=path)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print("n=%d d=%d k=%d it=%d %s: vs f64 %.3e, %.3f ms (%.1f us/iter), nnz %.3f" % (
            n, d, k, maxiter, path, rel_fro(z.double().cpu(), want.cpu()), dt * 1e3, dt * 1e6 / maxiter,
            float((z != 0).float().mean())))
    print("   gram vs ffma %.3e; hist rel %.3e; fallbacks %d" % (
        rel_fro(res["gram"][0].cpu(), res["ffma"][0].cpu()),
        float(((res["gram"][1] - res["ffma"][1]).abs() / res["ffma"][1].abs().clamp_min(1e-30)).max()),
        _cabi.load().lasso_b200_resident_fallbacks()))
