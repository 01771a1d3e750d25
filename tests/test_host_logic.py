"""Host-side logic of the drop-in boundary: argument handling, error behaviour of the
reference API, the deferred global stop rule under a 2-rank gloo group.  CPU only:
the CUDA calls are replaced by an oracle-backed stub where a result is needed."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lasso_b200
import oracle
from lasso_b200.linear import initialize_code, sparse_encode
from lasso_b200.linear.solvers import ista
from lasso_b200.linear.solvers import ista as ista_fn
from lasso_b200.testing import make_problem, rel_fro

ista_mod = __import__("sys").modules["lasso_b200.linear.solvers.ista"]


def test_api_surface_matches_reference_names():
    lin = lasso_b200.linear
    for name in ("sparse_encode", "initialize_code", "dict_learning", "dict_evaluate",
                 "update_dict", "update_dict_ridge", "solvers", "utils"):
        assert hasattr(lin, name)
    import inspect
    sig = inspect.signature(ista_fn)
    want = [("alpha", 1.0), ("fast", True), ("lr", 'auto'), ("maxiter", 10), ("tol", 1e-5),
            ("backtrack", False), ("eta_backtrack", 1.5), ("verbose", False)]
    names = list(sig.parameters)
    assert names[:3] == ["x", "z0", "weight"]
    for key, default in want:
        assert sig.parameters[key].default == default
    sig = inspect.signature(lin.sparse_encode)
    assert list(sig.parameters)[:6] == ["x", "weight", "alpha", "z0", "algorithm", "init"]
    sig = inspect.signature(lin.dict_learning)
    assert list(sig.parameters)[:9] == ["X", "n_components", "alpha", "constrained", "persist",
                                        "lambd", "steps", "device", "progbar"]
    assert sig.parameters["steps"].default == 60 and sig.parameters["lambd"].default == 1e-2


def test_error_behaviour_of_the_reference_is_kept():
    x, w = torch.randn(6, 4), torch.randn(4, 5)
    with pytest.raises(ValueError, match="invalid algorithm parameter 'nope'"):
        sparse_encode(x, w, algorithm="nope")
    with pytest.raises(ValueError, match="invalid init parameter 'bogus'"):
        sparse_encode(x, w, init="bogus")
    with pytest.raises(AssertionError):
        sparse_encode(x, w, z0=torch.zeros(6, 4))
    with pytest.raises(ValueError, match="eta must be > 1"):
        ista(x, torch.zeros(6, 5), w, lr=0.1, backtrack=True, eta_backtrack=1.0)
    with pytest.raises(NotImplementedError):
        sparse_encode(x, w, algorithm="cd")
    with pytest.raises(NotImplementedError):
        sparse_encode(x.double(), w.double(), lr=0.1)
    z0 = torch.zeros(6, 5)
    assert ista(x, z0, w, lr=0.1, maxiter=0) is z0      # the reference returns z0 itself


def test_initialize_code_modes():
    x, w = make_problem(12, 5, 9, seed=1)
    assert torch.equal(initialize_code(x, w, 0.3, "zero"), torch.zeros(12, 9))
    u = initialize_code(x, w, 0.3, "unif")
    assert u.shape == (12, 9) and float(u.abs().max()) <= 0.1
    assert rel_fro(initialize_code(x, w, 0.3, "transpose"), x @ w) == 0
    # 'ridge' runs in the library (lasso_b200_ridge_init_f32): no CUDA device, no result -- never a CPU fallback
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            initialize_code(x, w, 0.3, "ridge")


def test_momentum_schedule_and_tolerance_helpers():
    betas = oracle.beta_schedule(5)
    assert betas[0] == 0.0
    t1 = (1 + math.sqrt(5)) / 2
    assert betas[1] == pytest.approx((t1 - 1) / ((1 + math.sqrt(1 + 4 * t1 * t1)) / 2))
    assert ista_mod._abs_tolerance(65536 * 256, 1e-5) == pytest.approx(167.77216, rel=1e-6)
    hist = torch.tensor([5.0, 3.0, 0.5, 0.4, 0.1], dtype=torch.float64)
    assert ista_mod._first_stop(hist, 1.0) == 3
    assert ista_mod._first_stop(hist, 0.01) == 5
    assert ista_mod._first_stop(hist, 0.1) == 5   # the last iteration never "stops early"


# ---------------------------------------------------------------------------------------
# 2-rank gloo: the deferred, all-reduced stop rule must reproduce the unsharded result
# ---------------------------------------------------------------------------------------

def _oracle_backed_fista_device(x, weight, z0, alpha, lr, maxiter, fast, tol_abs, path=0,
                                want_iters=False, want_hist=False, out=None):
    n, k = x.shape[0], weight.shape[1]
    start = z0 if z0 is not None else torch.zeros(n, k)
    tol = -1.0 if tol_abs < 0 else tol_abs / max(start.numel(), 1)
    z, done, deltas = oracle.ista(x, start, weight, alpha=alpha, fast=bool(fast), lr=lr,
                                  maxiter=maxiter, tol=tol, return_info=True)
    hist = torch.zeros(maxiter, dtype=torch.float64)
    hist[:len(deltas)] = torch.tensor(deltas, dtype=torch.float64)
    return z, (done if want_iters else None), (hist if want_hist else None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sharded_worker(rank, world, port, tol, maxiter, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lasso_b200 as pkg
        mod = __import__("sys").modules["lasso_b200.linear.solvers.ista"]
        pkg._cabi.fista_device = _oracle_backed_fista_device
        x, w = make_problem(64, 16, 32, seed=len("ista_earlystop"), kind="planted")
        lr = 1.0 / oracle.lipschitz_constant(w)
        rows = slice(rank * 32, (rank + 1) * 32)
        xs = x[rows].clone()
        # pretend the shard is a CUDA tensor for the host logic only
        class Shard(torch.Tensor):
            @property
            def is_cuda(self):
                return True
        xs = xs.as_subclass(Shard)
        z, done = mod.solve(xs, None, w, alpha=0.1, fast=True, lr=lr, maxiter=maxiter, tol=tol,
                            group=dist.group.WORLD, return_iters=True)
        torch.save({"z": torch.Tensor(z), "done": done}, os.path.join(result_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tol,maxiter", [(1e-3, 500), (0.0, 40)])
def test_sharded_stop_rule_two_ranks(tmp_path, tol, maxiter):
    port = _free_port()
    mp.spawn(_sharded_worker, args=(2, port, tol, maxiter, str(tmp_path)), nprocs=2, join=True)
    x, w = make_problem(64, 16, 32, seed=len("ista_earlystop"), kind="planted")
    lr = 1.0 / oracle.lipschitz_constant(w)
    want, want_done, _ = oracle.ista(x, torch.zeros(64, 32), w, alpha=0.1, fast=True, lr=lr,
                                     maxiter=maxiter, tol=tol, return_info=True)
    parts = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(2)]
    got = torch.cat([p["z"] for p in parts])
    assert parts[0]["done"] == parts[1]["done"] == want_done
    if tol > 0:
        assert want_done < maxiter
    # rows are independent: the sharded run equals the unsharded one bit for bit
    assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------
# 2-rank gloo: the M-step all-reduces the packed statistics [Z^T Z | Z^T X] (and the two loss sums)
# once, then every rank runs the same atom sweep: dictionaries identical on all ranks and equal to the
# unsharded update
# ---------------------------------------------------------------------------------------

class _FakeCuda(torch.Tensor):
    """A CPU tensor that answers is_cuda = True (host logic only; the CUDA calls are stubbed)."""
    @property
    def is_cuda(self):
        return True


def _stub_gram(z, x, out_zz=None, out_zx=None):
    z64, x64 = torch.Tensor(z).double(), torch.Tensor(x).double()
    gzz, gzx = z64.T @ z64, z64.T @ x64
    if out_zz is not None:
        out_zz.copy_(gzz)
        out_zx.copy_(gzx)
        return out_zz, out_zx
    return gzz, gzx


def _stub_zero_columns(z, mask):
    torch.Tensor(z)[:, mask.bool()] = 0
    return z


def _stub_dict_update_gram(dictionary, gzz, gzx, eps=1e-10, redraw=None, positive=False):
    new, zeroed = oracle.update_dict_gram(torch.Tensor(dictionary), gzz, gzx, eps=eps, positive=positive)
    keep = [j for j in range(new.size(1)) if j not in zeroed]
    torch.Tensor(dictionary)[:, keep] = new[:, keep]
    mask = torch.zeros(new.size(1), dtype=torch.int32)
    mask[zeroed] = 1
    return mask


def _stub_loss_terms(x, z, w, out=None):
    x, z, w = torch.Tensor(x).double(), torch.Tensor(z).double(), torch.Tensor(w).double()
    terms = torch.stack([(z @ w.T - x).square().sum(), z.abs().sum()])
    if out is not None:
        out.copy_(terms)
        return out
    return terms


def _mstep_problem():
    x, w = make_problem(96, 12, 24, seed=31, kind="planted")
    z = oracle.ista(x, torch.zeros(96, 24), w, alpha=0.05, lr=1.0 / oracle.lipschitz_constant(w), maxiter=40, tol=0.0)
    z[:, 5] = 0                       # an unused atom: flagged on every rank, re-drawn identically (broadcast)
    return x, w, z


def _mstep_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lasso_b200 as pkg
        from lasso_b200.linear import lasso_loss, update_dict, update_dict_ridge
        pkg._cabi.gram = _stub_gram
        pkg._cabi.dict_update_gram = _stub_dict_update_gram
        pkg._cabi.loss_terms = _stub_loss_terms
        pkg._cabi.zero_columns = _stub_zero_columns
        x, w, z = _mstep_problem()
        rows = slice(0, 40) if rank == 0 else slice(40, 96)          # ragged shards
        xs, zs = x[rows].clone().as_subclass(_FakeCuda), z[rows].clone().as_subclass(_FakeCuda)
        w = w.as_subclass(_FakeCuda)
        torch.manual_seed(100 + rank)                                 # ranks draw differently: rank 0's draw wins
        d_new = update_dict(torch.Tensor(w).clone().as_subclass(_FakeCuda), xs, zs, group=dist.group.WORLD)
        v_new = update_dict_ridge(xs, zs, lambd=1e-2, group=dist.group.WORLD)
        loss = lasso_loss(xs, zs, w, 0.05, group=dist.group.WORLD)
        torch.save({"d": torch.Tensor(d_new).clone(), "v": torch.Tensor(v_new).clone(), "loss": float(loss),
                    "z5": float(torch.Tensor(zs)[:, 5].abs().max())},
                   os.path.join(result_dir, "m%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_sharded_mstep_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_mstep_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    x, w, z = _mstep_problem()
    parts = [torch.load(os.path.join(str(tmp_path), "m%d.pt" % r)) for r in range(2)]
    # replicated dictionary without a broadcast of the result: identical bits on both ranks
    assert torch.equal(parts[0]["d"], parts[1]["d"]) and torch.equal(parts[0]["v"], parts[1]["v"])
    assert parts[0]["loss"] == parts[1]["loss"]
    want = oracle.update_dict(w.clone(), x, z.clone())
    keep = [j for j in range(24) if j != 5]
    assert rel_fro(parts[0]["d"][:, keep], want[:, keep]) <= 1e-5
    assert abs(float(parts[0]["d"][:, 5].norm()) - 1.0) <= 1e-6 and parts[0]["z5"] == 0.0
    assert rel_fro(parts[0]["v"], oracle.update_dict_ridge(x, z, lambd=1e-2)) <= 1e-5
    assert abs(parts[0]["loss"] - float(oracle.lasso_loss(x, z, w, 0.05))) <= 1e-6 * abs(parts[0]["loss"])


# ---------------------------------------------------------------------------------------
# 2-rank gloo: a whole sharded dict_learning step is ONE all-reduce (statistics, loss sums and the
# E-step's stop-test sums in one buffer) and reproduces the unsharded run
# ---------------------------------------------------------------------------------------

def _dl_worker(rank, world, port, tol, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lasso_b200 as pkg
        dl = __import__("sys").modules["lasso_b200.linear.dict_learning"]
        pkg._cabi.fista_device = _oracle_backed_fista_device
        pkg._cabi.gram = _stub_gram
        pkg._cabi.dict_update_gram = _stub_dict_update_gram
        pkg._cabi.loss_terms = _stub_loss_terms
        pkg._cabi.zero_columns = _stub_zero_columns
        dl.default_device = lambda: torch.device("cpu")
        calls = {"all_reduce": 0, "broadcast": 0}
        real_ar, real_bc = dist.all_reduce, dist.broadcast

        def counted_ar(*a, **k):
            calls["all_reduce"] += 1
            return real_ar(*a, **k)

        def counted_bc(*a, **k):
            calls["broadcast"] += 1
            return real_bc(*a, **k)
        dist.all_reduce, dist.broadcast = counted_ar, counted_bc
        x, _ = make_problem(96, 12, 24, seed=5, kind="planted")
        rows = slice(0, 40) if rank == 0 else slice(40, 96)
        torch.manual_seed(0)
        steps = 4
        w, losses = dl.dict_learning(x[rows].clone(), 24, alpha=0.05, steps=steps, progbar=False, lr=0.1,
                                     maxiter=30, tol=tol, group=dist.group.WORLD)
        torch.save({"w": w, "losses": losses, "calls": calls, "steps": steps},
                   os.path.join(result_dir, "d%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tol", [0.0, 3e-3])
def test_sharded_dict_learning_one_all_reduce_per_step(tmp_path, tol):
    port = _free_port()
    mp.spawn(_dl_worker, args=(2, port, tol, str(tmp_path)), nprocs=2, join=True)
    parts = [torch.load(os.path.join(str(tmp_path), "d%d.pt" % r)) for r in range(2)]
    assert torch.equal(parts[0]["w"], parts[1]["w"]) and torch.equal(parts[0]["losses"], parts[1]["losses"])
    steps = parts[0]["steps"]
    x, _ = make_problem(96, 12, 24, seed=5, kind="planted")
    torch.manual_seed(0)
    want_w, want_losses = oracle.dict_learning(x, 24, alpha=0.05, steps=steps, lr=0.1, maxiter=30, tol=tol)
    assert rel_fro(parts[0]["w"], want_w) <= 1e-4
    assert torch.allclose(parts[0]["losses"], want_losses, rtol=1e-4)
    # one all-reduce for the global row count + one per EM step; with tol > 0 a step whose global stop
    # test fired early is redone once (a second all-reduce for that step); one broadcast of the
    # initial dictionary, none afterwards (no degenerate atoms here)
    n_ar = parts[0]["calls"]["all_reduce"]
    if tol == 0.0:
        assert n_ar == 1 + steps
    else:
        assert 1 + steps <= n_ar <= 1 + 2 * steps
    assert parts[0]["calls"]["broadcast"] == 1


@pytest.mark.parametrize("block", [1, 7, 64])
def test_blocked_gauss_seidel_is_the_sequential_sweep(block):
    """The algebra behind csrc/sweep_blk.cu, restated in float64 on the CPU: taking the atoms in blocks -- the part
    of u_j that does not depend on the block's own updates as ONE product for the whole block, then per atom only the
    corrections (d_l_new - d_l_start) A_jl of the block's earlier atoms -- is the sequential Gauss-Seidel sweep of
    update_dict in Gram space (oracle.update_dict_gram, dict_learning.py:82-101), unused atoms included: a degenerate
    atom's difference is -d_l_start, later blocks see its column of A as zero."""
    g = torch.Generator().manual_seed(4)
    n, d, k = 400, 9, 23
    z = torch.randn(n, k, generator=g, dtype=torch.float64) * (torch.rand(n, k, generator=g) < 0.3)
    unused = [0, 8, 22]
    z[:, unused] = 0
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    w0 = torch.nn.functional.normalize(torch.randn(d, k, generator=g, dtype=torch.float64), dim=0)
    a, b = z.T @ z, z.T @ x
    draws = torch.randn(d, k, generator=g, dtype=torch.float64)
    it = iter(unused)
    want, zeroed = oracle.update_dict_gram(w0, a, b, redraw=lambda m: draws[:, next(it)])
    assert zeroed == unused

    dmat, dead = w0.clone(), torch.zeros(k, dtype=torch.bool)
    for j0 in range(0, k, block):
        blk = range(j0, min(j0 + block, k))
        amask = a.clone()
        amask[:, dead] = 0                      # later blocks see a dead atom's column as zero
        u0 = {j: b[j] - dmat @ amask[j] + amask[j, j] * dmat[:, j] for j in blk}      # one product per block
        start, delta = dmat.clone(), {}
        for j in blk:
            u = u0[j] - sum(delta[l] * a[j, l] for l in delta)
            nrm = u.norm()
            if nrm < 1e-10:
                new = draws[:, j] / draws[:, j].norm()
                dead[j] = True
                delta[j] = -start[:, j]         # takes the atom's start value out of the later atoms' U0
            else:
                new = u / nrm
                delta[j] = new - start[:, j]
            dmat[:, j] = new
    assert sorted(torch.nonzero(dead).flatten().tolist()) == unused
    assert rel_fro(dmat, want) <= 1e-12


@pytest.mark.parametrize("gap", [1e-2, 1e-4, 1.2e-7, 1e-9])
def test_lipschitz_scheme_error_bound(gap):
    """The scheme of K2 (csrc/aux_kernels.cu), restated on the CPU: float32 power iteration on the 4096th power of the
    Gram (twelve trace-normalised squarings in float32), at most 256 steps, then ONE float64 Rayleigh quotient on the
    original Gram.  For any relative gap g between the two largest eigenvalues the quotient is off by at most
    g exp(-2 * 4096 * 256 g) <= 1.8e-7 -- including the gap that maximises that expression (1.2e-7)."""
    g = torch.Generator().manual_seed(11)
    m = 60
    q, _ = torch.linalg.qr(torch.randn(m, m, generator=g, dtype=torch.float64))
    ev = torch.linspace(1.0, 0.05, m, dtype=torch.float64)
    ev[0], ev[1] = 1.0, 1.0 - gap
    gram = (q * ev) @ q.T
    gram = 0.5 * (gram + gram.T)
    b = gram.to(torch.float32)
    for _ in range(12):
        b = b / b.diagonal().sum()
        b = b @ b
    v = torch.ones(m, dtype=torch.float32) + 0.25 * torch.rand(m, generator=g)
    for _ in range(256):
        u = b @ v
        v = u / u.norm()
    v64 = v.double()
    lam = float(v64 @ gram @ v64 / (v64 @ v64))
    assert 0.0 <= 1.0 - lam <= 2.5e-7


def test_gram_form_restatement_and_deferred_stop_test():
    """The algorithm of csrc/fista_gram.cu restated on the CPU in float64: the gradient y (W^T W) - x W is the
    reference's (y W^T - x) W (ista.py:71-73); the stop-test record of iteration i is taken while iteration i + 1 forms
    its extrapolation (|z_i - z_{i+1}| of the two buffers it reads anyway), one closing pass takes the last one; and
    running all iterations, reading the records and replaying with the count they give returns what the reference's
    in-place stop returns (ista.py:93-95)."""
    n, d, k, alpha, maxiter, tol = 60, 40, 24, 0.2, 200, 1e-4
    x, w = make_problem(n, d, k, seed=9, kind="planted")
    lr = 1.0 / oracle.lipschitz_constant(w)
    want, done, deltas = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, fast=True, lr=lr, maxiter=maxiter, tol=tol,
                                     return_info=True)
    assert 1 < done < maxiter
    x64, w64 = x.double(), w.double()
    gw, bxw = w64.T @ w64, x64 @ w64
    lr32, lam = float(torch.tensor(lr, dtype=torch.float32)), float(torch.tensor(alpha * lr, dtype=torch.float32))

    def run(iters):
        bufs = [torch.zeros(n, k, dtype=torch.float64), torch.zeros(n, k, dtype=torch.float64)]
        hist, t = [0.0] * iters, 1.0
        for it in range(iters + 1):
            zc, zp = bufs[it & 1], bufs[(it & 1) ^ 1]
            if it > 0:
                hist[it - 1] = float((zp - zc).abs().sum())       # the previous iteration's record
            if it == iters:
                break
            beta = 0.0
            if it > 0:
                t_next = (1 + math.sqrt(1 + 4 * t * t)) / 2
                beta, t = (t - 1) / t_next, t_next
            y = zc + beta * (zc - zp)
            grad = y @ gw - bxw
            assert rel_fro(grad, (y @ w64.T - x64) @ w64) <= 1e-12
            bufs[(it & 1) ^ 1] = torch.nn.functional.softshrink(y - lr32 * grad, lam)
        return bufs[iters & 1], hist

    _, hist = run(maxiter)
    stop = next(i + 1 for i, h in enumerate(hist) if h <= n * k * tol)
    assert stop == done
    z, hist2 = run(stop)
    assert rel_fro(z.float(), want) <= 2e-6
    assert torch.allclose(torch.tensor(hist2[:stop - 1], dtype=torch.float64),
                          torch.tensor([float(v) for v in deltas[:stop - 1]], dtype=torch.float64), rtol=1e-3)
