import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
dev = torch.device("cuda", 0)
for (n, d, k, dens) in [(131072, 64, 256, 0.08), (20000, 64, 256, 1.0), (9000, 24, 70, 0.1), (5000, 128, 200, 0.3), (4100, 10, 128, 0.5)]:
    g = torch.Generator().manual_seed(n + k)
    z = torch.randn(n, k, generator=g) * (torch.rand(n, k, generator=g) < dens) * 3.0
    x = torch.randn(n, d, generator=g) * 0.25
    zd, xd = z.to(dev), x.to(dev)
    print("case", n, d, k, "launching", flush=True)
    t0 = time.time()
    gzz, gzx = _cabi.gram(zd, xd)
    torch.cuda.synchronize()
    print("  tc done in %.3f s" % (time.time() - t0), flush=True)
    os.environ["LASSO_B200_GRAM"] = "ffma"
    fzz, fzx = _cabi.gram(zd, xd)
    torch.cuda.synchronize()
    del os.environ["LASSO_B200_GRAM"]
    w = z.double().T @ z.double()
    print("  rel err tc %.2e ffma %.2e" % (float((gzz.cpu() - w).norm() / w.norm()), float((fzz.cpu() - w).norm() / w.norm())), flush=True)
