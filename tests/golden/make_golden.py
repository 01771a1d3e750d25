"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference package cannot be imported as shipped on scipy >= 1.12
(lasso/linear/solvers/iterative_ridge.py:5 imports a private name that moved);
the shim below restores that one name and nothing else.  Every case pins ``lr``
to a float because the reference's lr='auto' (ARPACK on a float32 Gram) is not
reproducible run to run.

Fixtures are small compressed .npz files: inputs, the options used and the
reference's outputs.  tests/test_oracle_golden.py checks the oracle against
them on CPU; tests/test_gpu_parity.py checks the CUDA path against them.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def import_reference():
    import scipy.optimize.optimize as legacy  # noqa: F401  (deprecated alias module)
    from scipy.optimize._optimize import _status_message
    legacy._status_message = _status_message
    sys.path.insert(0, REFERENCE)
    import lasso.linear as ref_linear
    from lasso.linear.solvers.ista import ista as ref_ista
    return ref_linear, ref_ista


def lipschitz64(w):
    w64 = w.double()
    return float(torch.linalg.eigvalsh(w64 @ w64.T)[-1])


def main():
    warnings.simplefilter("ignore")
    import lasso_b200  # only for the seeded problem generator
    from lasso_b200.testing import make_problem

    torch.set_num_threads(1)  # the pinned-lr reference is thread-count independent; be safe
    ref, ref_ista = import_reference()
    out = {}

    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # fixture-name prefixes to (re)write

    def save(name, **arrays):
        if only and not any(name.startswith(o) for o in only):
            return
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                                     for k, v in arrays.items()})
        out[name] = os.path.getsize(path)

    # ---- ISTA / FISTA solver cases -------------------------------------------------
    cases = [
        # name, n, d, k, kind, alpha, opts
        ("ista_readme_fista", 128, 10, 50, "randn", 0.5, dict(fast=True, maxiter=10, tol=1e-5)),
        ("ista_readme_plain", 128, 10, 50, "randn", 0.5, dict(fast=False, maxiter=25, tol=1e-5)),
        ("ista_planted_200", 192, 64, 256, "planted", 0.1, dict(fast=True, maxiter=200, tol=0.0)),
        ("ista_randn_200", 160, 64, 256, "randn", 0.1, dict(fast=True, maxiter=200, tol=0.0)),
        ("ista_ragged", 77, 13, 37, "planted", 0.05, dict(fast=True, maxiter=60, tol=0.0)),
        ("ista_wide_d", 50, 150, 90, "randn", 0.3, dict(fast=True, maxiter=40, tol=0.0)),
        ("ista_earlystop", 64, 16, 32, "planted", 0.1, dict(fast=True, maxiter=500, tol=1e-3)),
        ("ista_earlystop_plain", 64, 16, 32, "planted", 0.1, dict(fast=False, maxiter=500, tol=1e-3)),
        ("ista_one_iter", 40, 8, 24, "randn", 0.2, dict(fast=True, maxiter=1, tol=1e-5)),
        ("ista_big_alpha", 32, 8, 16, "randn", 50.0, dict(fast=True, maxiter=5, tol=0.0)),
    ]
    for name, n, d, k, kind, alpha, opts in cases:
        x, w = make_problem(n, d, k, seed=len(name), kind=kind)
        lr = 1.0 / lipschitz64(w)
        z0 = torch.zeros(n, k)
        z = ref_ista(x, z0, w, alpha=alpha, lr=lr, **opts)
        save(name, x=x, weight=w, z0=z0, z=z, alpha=alpha, lr=lr,
             fast=int(opts["fast"]), maxiter=opts["maxiter"], tol=opts["tol"])

    # warm start from a non-zero code
    x, w = make_problem(96, 20, 60, seed=7, kind="planted")
    g = torch.Generator().manual_seed(11)
    z0 = (torch.rand(96, 60, generator=g) - 0.5) * 0.2
    lr = 1.0 / lipschitz64(w)
    z = ref_ista(x, z0, w, alpha=0.1, lr=lr, fast=True, maxiter=15, tol=0.0)
    save("ista_warmstart", x=x, weight=w, z0=z0, z=z, alpha=0.1, lr=lr, fast=1, maxiter=15, tol=0.0)

    # backtracking line search (ista.py:17-54)
    x, w = make_problem(48, 12, 30, seed=3, kind="randn")
    lr = 4.0 / lipschitz64(w)  # deliberately too large so that the search shrinks it
    z0 = torch.zeros(48, 30)
    z = ref_ista(x, z0, w, alpha=0.2, lr=lr, fast=True, maxiter=20, tol=0.0, backtrack=True)
    save("ista_backtrack", x=x, weight=w, z0=z0, z=z, alpha=0.2, lr=lr, fast=1, maxiter=20,
         tol=0.0, backtrack=1, eta_backtrack=1.5)

    # ---- sparse_encode boundary ------------------------------------------------------
    x, w = make_problem(64, 10, 50, seed=5, kind="randn")
    lr = 1.0 / lipschitz64(w)
    for init in ("zero", "ridge", "transpose"):
        z = ref.sparse_encode(x, w, alpha=0.5, algorithm="ista", init=init, lr=lr, maxiter=12,
                              tol=0.0)
        z0 = ref.initialize_code(x, w, 0.5, init)
        save("encode_init_" + init, x=x, weight=w, z0=z0, z=z, alpha=0.5, lr=lr, fast=1,
             maxiter=12, tol=0.0)

    # ---- dictionary learning ---------------------------------------------------------
    x, w = make_problem(128, 10, 50, seed=9, kind="planted")
    z = ref.sparse_encode(x, w, alpha=0.2, algorithm="ista", lr=1.0 / lipschitz64(w), maxiter=30,
                          tol=0.0)
    ref_dl = sys.modules["lasso.linear.dict_learning"]  # the attribute is shadowed by the function
    loss = ref_dl.lasso_loss(x, z, w, 0.2)
    w_upd = ref_dl.update_dict(w.clone(), x, z.clone())
    w_ridge = ref_dl.update_dict_ridge(x, z, lambd=1e-2)
    save("mstep", x=x, weight=w, z=z, alpha=0.2, loss=loss, weight_update=w_upd,
         weight_ridge=w_ridge, lambd=1e-2)

    # degenerate atom: a code column that is all zero makes |u_j| < eps (dict_learning.py:91-98)
    z_deg = z.clone()
    z_deg[:, 3] = 0
    z_deg[:, 17] = 0
    torch.manual_seed(1234)
    w_deg = ref_dl.update_dict(w.clone(), x, z_deg)
    save("mstep_degenerate", x=x, weight=w, z=z_deg, weight_update=w_deg, zero_atoms=[3, 17])

    # positive=True: the atom is clamped at zero before it is normalised (dict_learning.py:87-88)
    w_pos = ref_dl.update_dict(w.clone(), x, z.clone(), positive=True)
    save("mstep_positive", x=x, weight=w, z=z, weight_update=w_pos)

    # ---- convolutional ISTA (lasso/conv2d/ista.py) -------------------------------------
    from lasso.conv2d.ista import ista_conv2d as ref_conv
    g = torch.Generator().manual_seed(77)

    def conv_problem(n, cin, size, filters, ksize, density=0.03):
        w = torch.randn(filters, cin, ksize, ksize, generator=g)
        w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
        o = size - ksize + 1
        code = torch.randn(n, filters, o, o, generator=g) * (torch.rand(n, filters, o, o, generator=g) < density)
        x = torch.nn.functional.conv_transpose2d(code, w) + 0.01 * torch.randn(n, cin, size, size, generator=g)
        return x, w, o

    def conv_lr(w):   # a safe step for any kernel size: 1 / (sum over filters of squared l1 norms) bound
        return 1.0 / float(w.flatten(1).abs().sum(1).square().sum())

    x, w, o = conv_problem(5, 1, 20, 16, 8)
    lr = conv_lr(w) * 4
    z0 = torch.zeros(5, 16, o, o)
    z = ref_conv(x, z0, w, alpha=0.05, fast=True, maxiter=25, lr=lr, tol=0.0)
    save("conv_8x8_fista", x=x, weight=w, z0=z0, z=z, alpha=0.05, lr=lr, fast=1, maxiter=25, tol=0.0)
    z = ref_conv(x, z0, w, alpha=0.05, fast=False, maxiter=12, lr=lr, tol=0.0)
    save("conv_8x8_plain", x=x, weight=w, z0=z0, z=z, alpha=0.05, lr=lr, fast=0, maxiter=12, tol=0.0)
    z0w = (torch.rand(5, 16, o, o, generator=g) - 0.5) * 0.1
    z = ref_conv(x, z0w, w, alpha=0.05, fast=True, maxiter=9, lr=lr, tol=0.0)
    save("conv_8x8_warmstart", x=x, weight=w, z0=z0w, z=z, alpha=0.05, lr=lr, fast=1, maxiter=9, tol=0.0)
    # odd kernel, several input channels, lr='auto' (Fourier bound, lip_const.py:96-135), early stop
    x, w, o = conv_problem(4, 4, 12, 12, 3, density=0.05)
    z0 = torch.zeros(4, 12, o, o)
    z = ref_conv(x, z0, w, alpha=0.1, fast=True, maxiter=300, lr='auto', tol=1e-3)
    from lasso.conv2d.lip_const import lip_bound_conv2d as ref_bound
    save("conv_3x3_auto_earlystop", x=x, weight=w, z0=z0, z=z, alpha=0.1, lr=-1.0, fast=1, maxiter=300,
         tol=1e-3, lip_bound=float(ref_bound(w, 0)))

    # stride / padding (ista.py:7, 18-19): patches of the zero-padded image on a stride grid
    def conv_problem_sp(n, cin, size, filters, ksize, stride, padding, density=0.05):
        w = torch.randn(filters, cin, ksize, ksize, generator=g)
        w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
        o = (size + 2 * padding - ksize) // stride + 1
        code = torch.randn(n, filters, o, o, generator=g) * (torch.rand(n, filters, o, o, generator=g) < density)
        x = torch.nn.functional.conv_transpose2d(code, w, stride=stride, padding=padding)
        assert x.shape[-1] == size
        return x + 0.01 * torch.randn(n, cin, size, size, generator=g), w, o

    for name, (cin, size, filters, ksize, stride, padding, fast, iters) in {
            "conv_8x8_pad3": (1, 20, 16, 8, 1, 3, True, 20),
            "conv_4x4_stride2": (4, 20, 16, 4, 2, 0, True, 20),
            "conv_4x4_stride2_pad1": (4, 20, 16, 4, 2, 1, False, 12)}.items():
        x, w, o = conv_problem_sp(4, cin, size, filters, ksize, stride, padding)
        lr = conv_lr(w) * 4
        z0 = torch.zeros(4, filters, o, o)
        z = ref_conv(x, z0, w, alpha=0.05, stride=stride, padding=padding, fast=fast, maxiter=iters, lr=lr, tol=0.0)
        save(name, x=x, weight=w, z0=z0, z=z, alpha=0.05, lr=lr, fast=int(fast), maxiter=iters, tol=0.0,
             stride=stride, padding=padding)
    x, w, o = conv_problem_sp(4, 4, 12, 12, 3, 1, 1)
    z0 = torch.zeros(4, 12, o, o)
    z = ref_conv(x, z0, w, alpha=0.1, stride=1, padding=1, fast=True, maxiter=40, lr='auto', tol=0.0)
    save("conv_3x3_pad1_auto", x=x, weight=w, z0=z0, z=z, alpha=0.1, lr=-1.0, fast=1, maxiter=40, tol=0.0,
         stride=1, padding=1, lip_bound=float(ref_bound(w, 1)))

    for constrained in (True, False):
        x, _ = make_problem(128, 10, 50, seed=21, kind="randn")
        torch.manual_seed(0)
        w0 = torch.empty(10, 50)
        torch.nn.init.orthogonal_(w0)
        if constrained:
            w0 = torch.nn.functional.normalize(w0, dim=0)
        torch.manual_seed(0)
        # lr='auto' is what dict_learning users get; its ARPACK value wobbles ~1e-6, which
        # the comparison tolerance of the dict_learning tests accounts for
        w_fin, losses = ref.dict_learning(x, 50, alpha=0.5, constrained=constrained, steps=8,
                                          lambd=1e-2, progbar=False, algorithm="ista", maxiter=10)
        save("dict_learning_" + ("constrained" if constrained else "ridge"), x=x, weight0=w0,
             weight=w_fin, losses=losses, alpha=0.5, steps=8, lambd=1e-2, maxiter=10)

    # ---- round 2 additions -----------------------------------------------------------------
    # shapes that reach the k-blocked tcgen05 kernel (d <= 128, k <= 1024 beyond the resident shape) and
    # the notebook's dictionary (d = 289, k = 300: examples/dict_learning_omniglot.ipynb:638-640)
    for name, n, d, k, kind, alpha, opts in [
            ("r2_blocked_128x512", 160, 128, 512, "planted", 0.05, dict(fast=True, maxiter=100, tol=0.0)),
            ("r2_blocked_64x512_plain", 130, 64, 512, "randn", 0.1, dict(fast=False, maxiter=40, tol=0.0)),
            ("r2_notebook_289x300", 96, 289, 300, "planted", 0.5, dict(fast=True, maxiter=20, tol=0.0))]:
        x, w = make_problem(n, d, k, seed=len(name), kind=kind)
        lr = 1.0 / lipschitz64(w)
        z0 = torch.zeros(n, k)
        z = ref_ista(x, z0, w, alpha=alpha, lr=lr, **opts)
        save(name, x=x, weight=w, z0=z0, z=z, alpha=alpha, lr=lr,
             fast=int(opts["fast"]), maxiter=opts["maxiter"], tol=opts["tol"])

    # init='lstsq' (sparse_encode.py:26-27, utils.py:13-25: least-norm branch since d < k)
    x, w = make_problem(64, 10, 50, seed=5, kind="randn")
    lr = 1.0 / lipschitz64(w)
    z = ref.sparse_encode(x, w, alpha=0.5, algorithm="ista", init="lstsq", lr=lr, maxiter=12, tol=0.0)
    z0 = ref.initialize_code(x, w, 0.5, "lstsq")
    save("r2_encode_init_lstsq", x=x, weight=w, z0=z0, z=z, alpha=0.5, lr=lr, fast=1, maxiter=12, tol=0.0)
    # init='unif' (sparse_encode.py:24-25) draws from the global CPU generator
    torch.manual_seed(4321)
    z = ref.sparse_encode(x, w, alpha=0.5, algorithm="ista", init="unif", lr=lr, maxiter=12, tol=0.0)
    torch.manual_seed(4321)
    z0 = ref.initialize_code(x, w, 0.5, "unif")
    save("r2_encode_init_unif", x=x, weight=w, z0=z0, z=z, alpha=0.5, lr=lr, fast=1, maxiter=12, tol=0.0,
         seed=4321)

    # backtracking that FAILS (ista.py:39-52): eta so close to 1 that 1000 shrinks of a far too large
    # step never satisfy F <= Q -> warning, revert to the initial step for this iteration
    x, w = make_problem(24, 6, 12, seed=13, kind="randn")
    lr = 1e6 / lipschitz64(w)
    z0 = torch.zeros(24, 12)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        z = ref_ista(x, z0, w, alpha=0.2, lr=lr, fast=True, maxiter=2, tol=0.0, backtrack=True,
                     eta_backtrack=1.0 + 1e-6)
    assert any("backtracking line search failed" in str(c.message) for c in caught)
    save("r2_backtrack_failure", x=x, weight=w, z0=z0, z=z, alpha=0.2, lr=lr, fast=1, maxiter=2, tol=0.0,
         backtrack=1, eta_backtrack=1.0 + 1e-6)

    # dict_learning with the step PINNED (lr travels through **solver_kwargs, dict_learning.py:25,38):
    # bit-reproducible reference, so the CUDA path is held to 1e-5 on W and 1e-6 on the losses
    for constrained in (True, False):
        x, _ = make_problem(128, 10, 50, seed=21, kind="randn")
        torch.manual_seed(0)
        w0 = torch.empty(10, 50)
        torch.nn.init.orthogonal_(w0)
        if constrained:
            w0 = torch.nn.functional.normalize(w0, dim=0)
        torch.manual_seed(0)
        w_fin, losses = ref.dict_learning(x, 50, alpha=0.5, constrained=constrained, steps=8,
                                          lambd=1e-2, progbar=False, algorithm="ista", maxiter=10, lr=0.05)
        save("r2_dict_learning_pinned_" + ("constrained" if constrained else "ridge"), x=x, weight0=w0,
             weight=w_fin, losses=losses, alpha=0.5, steps=8, lambd=1e-2, maxiter=10, lr=0.05)
    # ... and with persist=True + init='ridge' (what the notebook runs, ipynb:638-640)
    x, _ = make_problem(128, 10, 50, seed=22, kind="planted")
    torch.manual_seed(0)
    w0 = torch.nn.functional.normalize(torch.nn.init.orthogonal_(torch.empty(10, 50)), dim=0)
    torch.manual_seed(0)
    w_fin, losses = ref.dict_learning(x, 50, alpha=0.2, constrained=True, persist=True, steps=6,
                                      progbar=False, algorithm="ista", init="ridge", maxiter=8, lr=0.05)
    save("r2_dict_learning_persist_ridge_init", x=x, weight0=w0, weight=w_fin, losses=losses, alpha=0.2,
         steps=6, maxiter=8, lr=0.05)

    for name, size in sorted(out.items()):
        print("{:32s} {:8d} bytes".format(name, size))


if __name__ == "__main__":
    main()
