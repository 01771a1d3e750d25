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

# a dictionary whose largest singular values nearly coincide (what dictionary learning produces: on the
# notebook workload the iteration on G^16 ran into its cap of 2000 iterations, 3.4 ms)
g = torch.Generator().manual_seed(3)
d, k = 289, 300
u, _ = torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.float64))
v, _ = torch.linalg.qr(torch.randn(k, d, generator=g, dtype=torch.float64))
for gap in (1e-2, 1e-3, 1e-4, 1e-6):
    sv = torch.linspace(1.0, 0.2, d, dtype=torch.float64)
    sv[0], sv[1], sv[2] = 1.0 + gap, 1.0, 1.0 - gap
    w = ((u * sv) @ v.T).float().to(dev)
    want = float(torch.linalg.eigvalsh((w.double() @ w.double().T))[-1])
    for _ in range(3): got = _cabi.lipschitz(w)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): got = _cabi.lipschitz(w)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("K2 d=%d k=%d, singular-value gap %.0e: %.3f ms, rel err %.2e" % (d, k, gap, dt * 1e3, abs(got - want) / want))
