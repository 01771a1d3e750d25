"""Time the host entry point (pinned host tensors -> codes in pinned host memory) at config 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem
n, d, k, iters = 65536, 64, 256, 200
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / float(torch.linalg.eigvalsh((w.T @ w).double())[-1])
x, w = x.pin_memory(), w.pin_memory()
out = torch.empty(n, k).pin_memory()
torch.cuda.set_device(0)
for _ in range(3): _cabi.fista_host(x, w, None, 0.1, lr, iters, True, 0.0, out=out)
ts = []
for _ in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _cabi.fista_host(x, w, None, 0.1, lr, iters, True, 0.0, out=out)
    ts.append(time.perf_counter() - t0)
ref, _, _ = _cabi.fista_device(x.cuda(), w.cuda(), None, 0.1, lr, iters, True, 0.0)
print("e2e %s: median %.3f ms = %.0f it/s, min %.3f ms; equal to the device entry: %s" % (
    os.environ.get("LASSO_B200_PIPE", "default"), sorted(ts)[5] * 1e3, iters / sorted(ts)[5], min(ts) * 1e3,
    bool(torch.equal(out, ref.cpu()))))
