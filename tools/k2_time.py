"""Time the Lipschitz-constant kernels (K2) at the dictionary shapes of the BASELINE configs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200
from lasso_b200 import _cabi
from lasso_b200.testing import make_dictionary
dev = torch.device("cuda", 0)
for d, k in [(64, 256), (10, 50), (128, 1024), (289, 300)]:
    w = make_dictionary(d, k, seed=1).to(dev)
    want = float(torch.linalg.eigvalsh((w.double() @ w.double().T))[-1])
    for _ in range(3): got = _cabi.lipschitz(w)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): got = _cabi.lipschitz(w)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("K2 d=%d k=%d: %.3f ms, rel err %.2e (%s)" % (d, k, dt * 1e3, abs(got - want) / want,
          "generic" if os.environ.get("LASSO_B200_K2_GENERIC") else "default"))
