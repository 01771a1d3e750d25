""""also time reference-on-GPU (torch eager, one B200) as the stronger same-box baseline".
The reference's ista loop (linear/solvers/ista.py:57-104: two matmuls, softshrink, the stop-test sum with its
host sync, the momentum combine) written out with stock torch ops on cuda:0 -- cuBLAS fp32, no TF32 -- on the
C2 workload, next to this engine's device-resident number.  Not part of the product; prints one JSON line."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
n, d, k, alpha, iters = 65536, 64, 256, 0.1, 200
x, w = make_problem(n, d, k, seed=0)
x, w = x.to(dev), w.to(dev)
lr = 1.0 / float(torch.linalg.eigvalsh((w.T @ w).double())[-1])


def eager(tol):
    z = x.new_zeros(n, k)
    y, t = z, 1.0
    thresh = z.numel() * tol
    for _ in range(iters):
        g = torch.matmul(torch.matmul(y, w.T) - x, w)
        z_next = F.softshrink(y - lr * g, alpha * lr)
        if (z - z_next).abs().sum() <= thresh:      # host sync per iteration, as the reference has
            z = z_next
            break
        t_next = (1 + math.sqrt(1 + 4 * t * t)) / 2
        y = z_next + ((t - 1) / t_next) * (z_next - z)
        z, t = z_next, t_next
    return z


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out

t_e, z_e = timed(lambda: eager(0.0))
t_o, z_o = timed(lambda: _cabi.fista_device(x, w, None, alpha, lr, iters, True, -1.0), reps=10)
z_o = z_o[0] if isinstance(z_o, tuple) else z_o
print(json.dumps({"workload": "C2 n=65536 d=64 k=256 alpha=0.1 200 it fp32, inputs resident on cuda:0",
                  "torch_eager_iters_per_s": iters / t_e, "this_engine_iters_per_s": iters / t_o,
                  "rel_fro_between_them": float((z_o - z_e).norm() / z_e.norm())}), flush=True)
