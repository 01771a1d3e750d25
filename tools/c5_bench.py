"""BASELINE config 5: convolutional lasso, 28x28 images, 512 filters of 8x8, n=16384 -- im2col -> linear
on the k-blocked tcgen05 kernel.  Prints one JSON line (iterations/s through the C ABI entry point on
device tensors, HBM roofline of an iteration, parity on an image subset against the CPU oracle).

    python tools/c5_bench.py [--n 16384] [--iters 20]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import rel_fro

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
n, filters, size, ks = args.n, 512, 28, 8
o = size - ks + 1
P, d = o * o, ks * ks
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
w = torch.randn(filters, 1, ks, ks, generator=g, device=dev)
w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
x = torch.empty(n, 1, size, size, device=dev)
for lo in range(0, n, 1024):        # planted sparse codes, generated in slabs (the dense code tensor is 14.8 GB)
    hi = min(n, lo + 1024)
    code = torch.randn(hi - lo, filters, o, o, generator=g, device=dev) * (torch.rand(hi - lo, filters, o, o, generator=g, device=dev) < 0.002)
    x[lo:hi] = torch.nn.functional.conv_transpose2d(code, w) + 0.01 * torch.randn(hi - lo, 1, size, size, generator=g, device=dev)
del code
lr, alpha = 2e-3, 0.05
w_lin = w.reshape(filters, d).T.contiguous()
_cabi.conv2d_fista_device(x, w_lin, None, ks, ks, alpha, lr, 2, True, -1.0)     # warm-up (workspace)
torch.cuda.synchronize()
l0 = _cabi.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
z_rows, _ = _cabi.conv2d_fista_device(x, w_lin, None, ks, ks, alpha, lr, args.iters, True, -1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
launches = _cabi.launch_count() - l0
sub = 8
got = z_rows[: sub * P].reshape(sub, o, o, filters).permute(0, 3, 1, 2).cpu()
want = oracle.conv2d_ista(x[:sub].cpu(), torch.zeros(sub, filters, o, o), w.cpu(), alpha=alpha, fast=True,
                          maxiter=args.iters, lr=lr, tol=0.0)
rows = n * P
peaks = {"hbm_gbs": 6650.0}
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peaks = json.load(open(pk))
it_us = ms * 1e3 / args.iters
alg_bytes = rows * (5 * filters + 4 * d) * 4     # codes: 2 reads x 2 passes + 1 write; R / r: write, read, write, read
print(json.dumps({
    "workload": "configs[4]: conv2d lasso, %d images 28x28, 512 filters 8x8 (im2col rows %d x 64, codes %.1f GB), %d FISTA iterations" % (n, rows, rows * filters * 4 / 1e9, args.iters),
    "value": args.iters / (ms * 1e-3), "unit": "iters/s", "ms_per_iter": it_us / 1e3, "gpu_launches": int(launches),
    "rel_err_vs_oracle_on_%d_images" % sub: rel_fro(got, want),
    "roofline": {"bound": "hbm", "achieved": alg_bytes / (it_us * 1e-6) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": alg_bytes / (it_us * 1e-6) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_iteration": alg_bytes}}), flush=True)
