"""GPU diagnostic: resident tcgen05 path vs streaming tcgen05 path vs CPU oracle vs float64 gold."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem, rel_fro, support_mismatch

dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv
cases = [(256, 64, 256, "planted", 0.1, 1.0), (256, 64, 256, "randn", 0.1, 1.0), (384, 16, 32, "planted", 0.1, 1.0),
         (100, 20, 60, "planted", 0.1, 1.0), (1000, 64, 128, "randn", 0.05, 1.0), (130, 8, 24, "randn", 0.2, 1.0),
         (300, 64, 256, "planted", 0.1, 1e-3), (300, 64, 256, "randn", 0.1, 1e3), (20000, 64, 256, "planted", 0.1, 1.0)]
bad = 0
for n, d, k, kind, alpha, scale in cases:
    x, w = make_problem(n, d, k, seed=0, kind=kind)
    x = x * scale
    alpha = alpha * scale
    lr = 1.0 / oracle.lipschitz_constant(w)
    xd, wd = x.to(dev), w.to(dev)
    for iters in ((1, 2, 3, 10, 50, 200) if n < 5000 else (50,)):
        try:
            zr, _, hist = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="resident", want_hist=True)
            torch.cuda.synchronize()
        except Exception as e:
            print("RESIDENT FAILED", n, d, k, kind, iters, e); sys.exit(1)
        ztc, _, htc = _cabi.fista_device(xd, wd, None, alpha, lr, iters, True, -1.0, path="tcgen05", want_hist=True)
        z32 = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
        if n < 5000:
            z64 = torch.from_numpy(oracle.ista_f64(x.numpy(), torch.zeros(n, k).numpy(), w.numpy(), alpha, lr, iters))
        else:
            z64 = z32
        e = rel_fro(zr, z32)
        hd = float(((hist - htc).abs() / htc.abs().clamp_min(1e-30)).max())
        flagged = "" if e <= 1e-5 else "  <-- FAIL"
        bad += e > 1e-5
        print("n=%5d d=%2d k=%3d %-7s s=%g it=%3d | res~ref32 %.2e res~tc %.2e res~f64 %.2e ref32~f64 %.2e supp %.1e hist %.1e fb=%d%s" % (
            n, d, k, kind, scale, iters, e, rel_fro(zr, ztc), rel_fro(zr, z64), rel_fro(z32, z64),
            support_mismatch(zr.cpu(), z32), hd, _cabi.resident_fallbacks(), flagged), flush=True)
# warm start + ISTA (fast=False) + stop test
n, d, k = 500, 64, 256
x, w = make_problem(n, d, k, seed=3, kind="planted")
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
z0 = 0.1 * torch.randn(n, k)
for fast in (True, False):
    zr, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 30, fast, -1.0, path="resident")
    z32 = oracle.ista(x, z0, w, alpha=0.1, fast=fast, lr=lr, maxiter=30, tol=0.0)
    print("warm start fast=%s: %.2e" % (fast, rel_fro(zr, z32)))
    bad += rel_fro(zr, z32) > 1e-5
for tol in (1e-3, 1e-4):
    tol_abs = n * k * tol
    zr, it_r, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 500, True, tol_abs, path="resident", want_iters=True)
    zt, it_t, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 500, True, tol_abs, path="tcgen05", want_iters=True)
    z32 = oracle.ista(x, torch.zeros(n, k), w, alpha=0.1, lr=lr, maxiter=500, tol=tol)
    print("stop test tol=%g: resident %d it, streaming %d it; res~ref32 %.2e tc~ref32 %.2e" % (
        tol, it_r, it_t, rel_fro(zr, z32), rel_fro(zt, z32)))
# in-place (z0 aliases out)
zbuf = z0.to(dev).clone()
zr, _, _ = _cabi.fista_device(xd, wd, zbuf, 0.1, lr, 30, True, -1.0, path="resident", out=zbuf)
z32 = oracle.ista(x, z0, w, alpha=0.1, lr=lr, maxiter=30, tol=0.0)
print("aliased z0/out: %.2e" % rel_fro(zr, z32))
# overflow fallback: huge dynamic range inside one batch
xb = x.clone(); xb[0] *= 1e6
zr, _, _ = _cabi.fista_device(xb.to(dev), wd, None, 0.1, lr, 20, True, -1.0, path="resident")
z32 = oracle.ista(xb, torch.zeros(n, k), w, alpha=0.1, lr=lr, maxiter=20, tol=0.0)
print("wide-range batch: %.2e, fallbacks=%d" % (rel_fro(zr, z32), _cabi.resident_fallbacks()))
if quick:
    sys.exit(1 if bad else 0)
# timing at C2
n, d, k = 65536, 64, 256
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / oracle.lipschitz_constant(w)
xd, wd = x.to(dev), w.to(dev)
out = torch.empty(n, k, device=dev)
for path in ("resident", "tcgen05"):
    for _ in range(2):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, -1.0, path=path, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        _cabi.fista_device(xd, wd, None, 0.1, lr, 200, True, -1.0, path=path, out=out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("C2 %s: %.1f us/iter, %.0f it/s" % (path, dt / 200 * 1e6, 200 / dt))
z32 = None
sys.exit(1 if bad else 0)
