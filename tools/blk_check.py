"""GPU diagnostic: k-blocked tcgen05 path vs FFMA path vs CPU oracle (config-3-like shapes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro, support_mismatch

dev = torch.device("cuda", 0)
cases = [(300, 128, 1024, "planted", 0.05, (1, 2, 3, 10, 40)), (200, 128, 1024, "randn", 0.05, (1, 2, 30)),
         (1000, 64, 512, "planted", 0.1, (1, 2, 25)), (130, 100, 300, "randn", 0.1, (1, 3, 20)),
         (257, 72, 320, "planted", 0.1, (2, 15))]
bad = 0
for n, d, k, kind, alpha, its in cases:
    x, w = make_problem(n, d, k, seed=0, kind=kind)
    lr = 1.0 / oracle.lipschitz_constant(w)
    xd, wd = x.to(dev), w.to(dev)
    for iters in its:
        try:
            zb, _, hb = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="blocked", want_hist=True)
            torch.cuda.synchronize()
        except Exception as e:
            print("BLOCKED FAILED", n, d, k, kind, iters, e); sys.exit(1)
        zf, _, hf = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="ffma", want_hist=True)
        z32 = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
        e = rel_fro(zb, z32)
        hd = float(((hb - hf).abs() / hf.abs().clamp_min(1e-30)).max())
        bad += e > 1e-5
        print("n=%5d d=%3d k=%4d %-7s it=%3d | blk~ref32 %.2e blk~ffma %.2e supp %.1e hist %.1e fb=%d%s" % (
            n, d, k, kind, iters, e, rel_fro(zb, zf), support_mismatch(zb.cpu(), z32), hd, _cabi.resident_fallbacks(),
            "" if e <= 1e-5 else "  <-- FAIL"), flush=True)
# warm start, plain ISTA, stop test
n, d, k = 400, 128, 768
x, w = make_problem(n, d, k, seed=3)
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
z0 = 0.05 * torch.randn(n, k)
for fast in (True, False):
    zb, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 20, fast, -1.0, path="blocked")
    z32 = oracle.ista(x, z0, w, alpha=0.1, fast=fast, lr=lr, maxiter=20, tol=0.0)
    print("warm start fast=%s: %.2e" % (fast, rel_fro(zb, z32)))
    bad += rel_fro(zb, z32) > 1e-5
tol_abs = n * k * 1e-3
zb, it_b, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="blocked", want_iters=True)
zf, it_f, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="ffma", want_iters=True)
print("stop test: blocked %d it, ffma %d it; blk~ffma %.2e" % (it_b, it_f, rel_fro(zb, zf)))
if "--quick" in sys.argv:
    sys.exit(1 if bad else 0)
# timing at C3
n, d, k = 262144, 128, 1024
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
out = torch.empty(n, k, device=dev)
for path, iters in (("blocked", 40), ("ffma", 5)):
    _cabi.fista_device(xd, wd, None, 0.05, lr, iters, True, -1.0, path=path, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _cabi.fista_device(xd, wd, None, 0.05, lr, iters, True, -1.0, path=path, out=out)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("C3 %s: %.1f us/iter, %.0f it/s" % (path, dt / iters * 1e6, iters / dt))
sys.exit(1 if bad else 0)
