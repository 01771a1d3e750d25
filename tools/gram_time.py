"""Time the M-step statistics (K3) at config 4's per-GPU share, the notebook's shapes and a larger dictionary: tcgen05 kernel vs the FFMA kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
dev = torch.device("cuda", 0)
for n, d, k, dens in [(131072, 64, 256, 0.08), (10000, 289, 300, 0.05), (65536, 128, 512, 0.05)]:
    g = torch.Generator(device=dev).manual_seed(0)
    z = torch.randn(n, k, generator=g, device=dev) * (torch.rand(n, k, generator=g, device=dev) < dens)
    x = torch.randn(n, d, generator=g, device=dev)
    for mode in ("tcgen05", "ffma"):
        if mode == "ffma": os.environ["LASSO_B200_GRAM"] = "ffma"
        else: os.environ.pop("LASSO_B200_GRAM", None)
        for _ in range(3): out = _cabi.gram(z, x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): out = _cabi.gram(z, x)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        want = z.double().T @ z.double()
        print("gram n=%d d=%d k=%d %s: %.3f ms, rel err %.2e" % (n, d, k, mode, dt * 1e3, float((out[0] - want).norm() / want.norm())))
    os.environ.pop("LASSO_B200_GRAM", None)
