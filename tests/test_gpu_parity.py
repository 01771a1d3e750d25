"""Parity of the CUDA path (through the C ABI) with the reference's golden vectors and the
CPU oracle.  Tolerance: 1e-5 relative Frobenius on the returned coefficients
(BASELINE.json north_star); the float32 reference itself sits 5e-7..3e-6 from float64."""
import numpy as np
import pytest
import torch

import lasso_b200
import oracle
from conftest import CONV_CASES, SOLVER_CASES, load_golden
from lasso_b200 import _cabi
from lasso_b200.linear import (dict_evaluate, dict_learning, lasso_loss, sparse_encode,
                               update_dict, update_dict_ridge)
from lasso_b200.linear.solvers import ista, lipschitz_constant
from lasso_b200.testing import make_problem, rel_fro, support_mismatch

pytestmark = pytest.mark.gpu
TOL = 1e-5
PATHS = ["ffma", "auto", "tcgen05"]   # auto = resident tcgen05 kernel where the shape fits


def _tc_shape(d, k):
    return d % 4 == 0 and k % 4 == 0 and 4 <= d <= 64 and 4 <= k <= 256


def _skip_unless_supported(path, d, k):
    # the streaming tcgen05 kernel moves rows by TMA (16-byte aligned rows: d, k multiples of 4);
    # the resident kernel takes any d <= 64, k <= 256, larger shapes go to FFMA
    if path == "tcgen05" and not _tc_shape(d, k):
        pytest.skip("the streaming tcgen05 kernel does not take d={} k={}".format(d, k))
    if path == "resident" and not (d <= 64 and k <= 256):
        pytest.skip("the resident kernel does not take d={} k={}".format(d, k))


@pytest.fixture(scope="module")
def dev():
    lasso_b200._cabi.load()  # fail loudly if the extension is missing
    return torch.device("cuda", 0)


def run_case(g, dev, path, **extra):
    z0 = g["z0"].to(dev)
    return ista(g["x"].to(dev), z0, g["weight"].to(dev), alpha=g["alpha"], fast=bool(g["fast"]),
                lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"], path=path, **extra)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", SOLVER_CASES)
def test_golden_solver_cases(dev, name, path):
    g = load_golden(name)
    _skip_unless_supported(path, g["weight"].shape[0], g["weight"].shape[1])
    z = run_case(g, dev, path)
    assert z.shape == g["z"].shape and z.dtype == torch.float32 and z.is_cuda
    assert rel_fro(z, g["z"]) <= TOL, name
    assert support_mismatch(z.cpu(), g["z"]) <= 2e-3


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("init", ["zero", "ridge", "transpose", "lstsq"])
def test_sparse_encode_inits(dev, init, path):
    g = load_golden(("r2_encode_init_" if init == "lstsq" else "encode_init_") + init)
    _skip_unless_supported(path, g["weight"].shape[0], g["weight"].shape[1])
    z0 = lasso_b200.linear.initialize_code(g["x"].to(dev), g["weight"].to(dev), g["alpha"], init)
    assert rel_fro(z0, g["z0"]) <= TOL          # the start code itself (sparse_encode.py:19-35)
    z = sparse_encode(g["x"].to(dev), g["weight"].to(dev), alpha=g["alpha"], algorithm="ista",
                      init=init, lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"], path=path)
    assert rel_fro(z, g["z"]) <= TOL


def test_sparse_encode_init_unif(dev):
    # init='unif' (sparse_encode.py:24-25) draws from the generator of x's device.  CPU tensors: the
    # reference's own draw (fixture, seeded CPU generator), solved on the GPU through the host entry.
    g = load_golden("r2_encode_init_unif")
    torch.manual_seed(int(g["seed"]))
    z = sparse_encode(g["x"], g["weight"], alpha=g["alpha"], algorithm="ista", init="unif", lr=g["lr"],
                      maxiter=int(g["maxiter"]), tol=g["tol"])
    assert not z.is_cuda and rel_fro(z, g["z"]) <= TOL
    # CUDA tensors: the CUDA generator's stream (what the reference consumes on a GPU), then the same solve
    xd, wd = g["x"].to(dev), g["weight"].to(dev)
    torch.manual_seed(77)
    want_z0 = xd.new_empty(xd.size(0), wd.size(1)).uniform_(-0.1, 0.1)
    torch.manual_seed(77)
    z0 = lasso_b200.linear.initialize_code(xd, wd, g["alpha"], "unif")
    assert torch.equal(z0, want_z0) and float(z0.abs().max()) <= 0.1
    torch.manual_seed(77)
    z = sparse_encode(xd, wd, alpha=g["alpha"], algorithm="ista", init="unif", lr=g["lr"],
                      maxiter=int(g["maxiter"]), tol=g["tol"])
    want = oracle.ista(g["x"], want_z0.cpu(), g["weight"], alpha=g["alpha"], lr=g["lr"],
                       maxiter=int(g["maxiter"]), tol=g["tol"])
    assert rel_fro(z, want) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_early_stop_count_and_history(dev, path):
    g = load_golden("ista_earlystop")
    _, want_done, want_deltas = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"],
                                            fast=True, lr=g["lr"], maxiter=int(g["maxiter"]),
                                            tol=g["tol"], return_info=True)
    tol_abs = float(np.float32(g["z0"].numel() * g["tol"]))
    z, done, hist = _cabi.fista_device(g["x"].to(dev), g["weight"].to(dev), None, g["alpha"],
                                       g["lr"], int(g["maxiter"]), True, tol_abs, path=path,
                                       want_iters=True, want_hist=True)
    assert done == want_done
    assert rel_fro(z, g["z"]) <= TOL
    got = hist.cpu().numpy()[:want_done - 1]
    np.testing.assert_allclose(got, np.array(want_deltas[:want_done - 1]), rtol=1e-4)


@pytest.mark.parametrize("path", PATHS)
def test_host_entry_point_equals_device_entry_point(dev, path):
    g = load_golden("ista_warmstart" if path == "tcgen05" else "ista_ragged")
    zd = run_case(g, dev, path)
    zh = ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True, lr=g["lr"],
              maxiter=int(g["maxiter"]), tol=g["tol"], path=path)
    assert not zh.is_cuda
    assert torch.equal(zh, zd.cpu())
    pinned = g["x"].pin_memory()
    zp = ista(pinned, g["z0"], g["weight"], alpha=g["alpha"], fast=True, lr=g["lr"],
              maxiter=int(g["maxiter"]), tol=g["tol"], path=path)
    assert torch.equal(zp, zh)


@pytest.mark.parametrize("path", PATHS)
def test_edge_cases(dev, path):
    # ragged shape for the FFMA kernel, the nearest shape the tcgen05 kernels take otherwise
    d, k = (8, 20) if path == "tcgen05" else (7, 19)
    x, w = make_problem(33, d, k, seed=2)
    xd, wd = x.to(dev), w.to(dev)
    lr = 1.0 / oracle.lipschitz_constant(w)
    # maxiter = 0 through the C ABI copies the start
    z0 = torch.rand(33, k, device=dev)
    z, done, _ = _cabi.fista_device(xd, wd, z0, 0.1, lr, 0, True, 0.0, path=path, want_iters=True)
    assert done == 0 and torch.equal(z, z0)
    z, _, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 0, True, 0.0, path=path)
    assert float(z.abs().max()) == 0.0
    # empty batch
    ze = sparse_encode(xd[:0], wd, alpha=0.1, lr=lr, maxiter=5, path=path)
    assert ze.shape == (0, k)
    # single row, single iteration, odd/even iteration counts land in the caller's buffer
    for iters in (1, 2, 3, 4):
        want = oracle.ista(x[:1], torch.zeros(1, k), w, alpha=0.1, lr=lr, maxiter=iters, tol=0.0)
        got = sparse_encode(xd[:1], wd, alpha=0.1, lr=lr, maxiter=iters, tol=0.0, path=path)
        assert rel_fro(got, want) <= TOL
    # result written in place over the start buffer (z_out aliases z0)
    z0 = torch.zeros(33, k, device=dev)
    out, _, _ = _cabi.fista_device(xd, wd, z0, 0.1, lr, 7, True, 0.0, path=path, out=z0)
    want = oracle.ista(x, torch.zeros(33, k), w, alpha=0.1, lr=lr, maxiter=7, tol=0.0)
    assert out.data_ptr() == z0.data_ptr() and rel_fro(z0, want) <= TOL
    # all codes shrink to zero -> delta == 0 -> the stop test fires at the first iteration
    z, done, _ = _cabi.fista_device(xd, wd, None, 1e3, lr, 9, True, 0.0, path=path, want_iters=True)
    assert done == 1 and float(z.abs().max()) == 0.0
    # non-contiguous inputs are accepted
    xt = torch.randn(d, 33, device=dev).T
    z = sparse_encode(xt, wd, alpha=0.1, lr=lr, maxiter=3, tol=0.0, path=path)
    want = oracle.ista(xt.cpu().contiguous(), torch.zeros(33, k), w, alpha=0.1, lr=lr, maxiter=3,
                       tol=0.0)
    assert rel_fro(z, want) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_deterministic_and_shard_invariant(dev, path):
    x, w = make_problem(4096, 64, 256, seed=0)
    xd, wd = x.to(dev), w.to(dev)
    lr = 1.0 / oracle.lipschitz_constant(w)
    kw = dict(alpha=0.1, lr=lr, maxiter=40, tol=0.0, path=path)
    full = sparse_encode(xd, wd, **kw)
    again = sparse_encode(xd, wd, **kw)
    assert torch.equal(full, again)
    # rows are independent lasso problems: any row split gives the same bits
    for shards in (2, 8, 3):
        parts = [sparse_encode(part.contiguous(), wd, **kw) for part in xd.chunk(shards)]
        assert torch.equal(torch.cat(parts), full)


def test_auto_takes_the_resident_kernel(dev):
    assert _cabi.select_path(65536, 64, 256) == _cabi.PATH_RESIDENT
    assert _cabi.select_path(128, 10, 50) == _cabi.PATH_RESIDENT    # unaligned rows are fine
    assert _cabi.select_path(128, 65, 50) == _cabi.PATH_FFMA        # d > 64
    assert _cabi.select_path(1000, 128, 1024) == _cabi.PATH_BLOCKED  # C3: dictionary exceeds one SM
    assert _cabi.select_path(1000, 150, 90) == _cabi.PATH_FFMA


@pytest.mark.parametrize("n,d,k,kind,alpha,iters", [
    (300, 128, 1024, "planted", 0.05, 40),      # BASELINE config 3 shape
    (200, 128, 1024, "randn", 0.05, 30),
    (1000, 64, 512, "planted", 0.1, 25),        # the conv2d config's im2col shape (64-dim patches, 512 filters)
    (130, 100, 300, "randn", 0.1, 20),          # ragged: last chunk and feature block partly empty
    (257, 72, 320, "planted", 0.1, 1),
])
def test_blocked_kernel_matches_oracle(dev, n, d, k, kind, alpha, iters):
    assert _cabi.select_path(n, d, k) == _cabi.PATH_BLOCKED
    fallbacks = _cabi.resident_fallbacks()
    x, w = make_problem(n, d, k, seed=0, kind=kind)
    lr = 1.0 / oracle.lipschitz_constant(w)
    got = sparse_encode(x.to(dev), w.to(dev), alpha=alpha, lr=lr, maxiter=iters, tol=0.0)     # auto
    want = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
    assert rel_fro(got, want) <= TOL
    assert support_mismatch(got.cpu(), want) <= 2e-3
    assert _cabi.resident_fallbacks() == fallbacks      # solved by the tensor-core kernel itself


def test_blocked_kernel_warm_start_stop_test_and_fallback(dev, monkeypatch):
    n, d, k = 400, 128, 768
    x, w = make_problem(n, d, k, seed=3)
    lr = 1.0 / oracle.lipschitz_constant(w)
    xd, wd = x.to(dev), w.to(dev)
    g = torch.Generator().manual_seed(5)
    z0 = 0.05 * torch.randn(n, k, generator=g)
    for fast in (True, False):
        got, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 20, fast, -1.0, path="blocked")
        want = oracle.ista(x, z0, w, alpha=0.1, fast=fast, lr=lr, maxiter=20, tol=0.0)
        assert rel_fro(got, want) <= TOL
    # batch-global stop test, evaluated on the device between launches: same count and history as FFMA
    tol_abs = float(np.float32(n * k * 1e-3))
    zb, it_b, hb = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="blocked",
                                      want_iters=True, want_hist=True)
    zf, it_f, hf = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="ffma",
                                      want_iters=True, want_hist=True)
    assert 1 < it_b < 400 and it_b == it_f and rel_fro(zb, zf) <= TOL
    np.testing.assert_allclose(hb.cpu().numpy()[:it_b - 1], hf.cpu().numpy()[:it_b - 1], rtol=1e-4)
    # in place over the start buffer, and the hand-over to the FFMA kernel from the intact start
    before = _cabi.resident_fallbacks()
    monkeypatch.setenv("LASSO_B200_RES_LIMIT", "1e-3")
    buf = z0.to(dev).clone()
    got, _, _ = _cabi.fista_device(xd, wd, buf, 0.1, lr, 9, True, -1.0, path="blocked", out=buf)
    monkeypatch.delenv("LASSO_B200_RES_LIMIT")
    assert _cabi.resident_fallbacks() == before + 1
    ffma, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 9, True, -1.0, path="ffma")
    assert torch.equal(got, ffma)


@pytest.mark.parametrize("n,d,k,kind,alpha,iters,fast", [
    (1000, 289, 300, "planted", 0.5, 20, True),   # the reference notebook's dictionary and solver settings
    (700, 289, 300, "randn", 0.1, 60, True),
    (513, 200, 132, "randn", 0.2, 30, False),     # plain ISTA, last slab and last k-step partly empty
    (300, 512, 320, "planted", 0.1, 40, True),    # the largest dictionary whose pieces fit TMEM
    (130, 129, 4, "randn", 0.1, 25, True),        # one k-step, one slab
    (20000, 160, 64, "planted", 0.1, 3, True),    # more tiles than SMs: several tiles per CTA
])
def test_gram_form_kernel_matches_oracle(dev, n, d, k, kind, alpha, iters, fast):
    """fista_gram.cu (d > 128, k <= 320): g = y (W^T W) - x W on tcgen05 against the oracle's two-GEMM float32 loop."""
    assert _cabi.select_path(n, d, k) == _cabi.PATH_GRAM
    fallbacks = _cabi.resident_fallbacks()
    x, w = make_problem(n, d, k, seed=2, kind=kind)
    lr = 1.0 / oracle.lipschitz_constant(w)
    rows = slice(0, n) if n <= 2000 else slice(n - 700, n)        # the oracle on a row subset of the big case
    xd, wd = x.to(dev), w.to(dev)
    got = ista(xd, torch.zeros(n, k, device=dev), wd, alpha=alpha, fast=fast, lr=lr, maxiter=iters, tol=0.0)   # auto
    want = oracle.ista(x[rows], torch.zeros(x[rows].size(0), k), w, alpha=alpha, fast=fast, lr=lr,
                       maxiter=iters, tol=0.0)
    assert rel_fro(got[rows], want) <= TOL
    assert support_mismatch(got[rows].cpu(), want) <= 2e-3
    assert _cabi.resident_fallbacks() == fallbacks      # solved by the tensor-core kernel itself
    ffma, _, _ = _cabi.fista_device(xd, wd, None, alpha, lr, iters, fast, -1.0, path="ffma")
    assert rel_fro(got, ffma) <= TOL
    # shapes it must leave to the FFMA kernel: k beyond the TMEM piece columns, k not a multiple of 4
    assert _cabi.select_path(n, 289, 324) == _cabi.PATH_FFMA
    assert _cabi.select_path(n, 289, 298) == _cabi.PATH_FFMA


def test_gram_form_kernel_warm_start_stop_test_and_fallback(dev, monkeypatch):
    g = load_golden("r2_notebook_289x300")
    z = run_case(g, dev, "gram")
    assert rel_fro(z, g["z"]) <= TOL
    n, d, k = 900, 289, 300
    x, w = make_problem(n, d, k, seed=3)
    lr = 1.0 / oracle.lipschitz_constant(w)
    xd, wd = x.to(dev), w.to(dev)
    gen = torch.Generator().manual_seed(5)
    z0 = 0.05 * torch.randn(n, k, generator=gen)
    for fast in (True, False):
        got, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 21, fast, -1.0, path="gram")
        want = oracle.ista(x, z0, w, alpha=0.1, fast=fast, lr=lr, maxiter=21, tol=0.0)
        assert rel_fro(got, want) <= TOL
    # batch-global stop test: all iterations run in one launch, the count comes from the recorded sums and the
    # run is replayed with it -- same count, history and codes as the FFMA kernel, which stops in place
    tol_abs = float(np.float32(n * k * 1e-3))
    zb, it_b, hb = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="gram",
                                      want_iters=True, want_hist=True)
    zf, it_f, hf = _cabi.fista_device(xd, wd, None, 0.1, lr, 400, True, tol_abs, path="ffma",
                                      want_iters=True, want_hist=True)
    assert 1 < it_b < 400 and it_b == it_f and rel_fro(zb, zf) <= TOL
    np.testing.assert_allclose(hb.cpu().numpy()[:it_b - 1], hf.cpu().numpy()[:it_b - 1], rtol=1e-4)
    assert float(hb[it_b - 1]) <= tol_abs < float(hb[it_b - 2])
    want, done, _ = oracle.ista(x, torch.zeros(n, k), w, alpha=0.1, lr=lr, maxiter=400, tol=1e-3, return_info=True)
    assert done == it_b and rel_fro(zb, want) <= TOL
    # in place over the start buffer (odd and even iteration counts put z_out on either side of the ping-pong)
    for iters in (8, 9):
        buf = z0.to(dev).clone()
        got, _, _ = _cabi.fista_device(xd, wd, buf, 0.1, lr, iters, True, -1.0, path="gram", out=buf)
        assert got.data_ptr() == buf.data_ptr()
        want = oracle.ista(x, z0, w, alpha=0.1, lr=lr, maxiter=iters, tol=0.0)
        assert rel_fro(got, want) <= TOL
    # the hand-over to the FFMA kernel from the intact start codes
    before = _cabi.resident_fallbacks()
    monkeypatch.setenv("LASSO_B200_RES_LIMIT", "1e-3")
    buf = z0.to(dev).clone()
    got, _, _ = _cabi.fista_device(xd, wd, buf, 0.1, lr, 9, True, -1.0, path="gram", out=buf)
    monkeypatch.delenv("LASSO_B200_RES_LIMIT")
    assert _cabi.resident_fallbacks() == before + 1
    ffma, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 9, True, -1.0, path="ffma")
    assert torch.equal(got, ffma)
    # CPU tensors: the host entry point (H2D, solve on the Gram-form kernel, D2H), same codes as the device entry
    z_host = ista(x, z0.clone(), w, alpha=0.1, lr=lr, maxiter=12, tol=0.0)
    z_dev = ista(xd, z0.to(dev), wd, alpha=0.1, lr=lr, maxiter=12, tol=0.0)
    assert not z_host.is_cuda and torch.equal(z_host, z_dev.cpu())
    # rows at wildly different scales: every row is rescaled on its own
    scale = 10.0 ** (8 * torch.rand(n, 1, generator=gen) - 4)
    got = ista(xd * scale.to(dev), torch.zeros(n, k, device=dev), wd, alpha=0.1, lr=lr, maxiter=30, tol=0.0)
    want = oracle.ista_f64((x * scale).double().numpy(), np.zeros((n, k)), w.double().numpy(), 0.1, lr, 30)
    err = (got.double().cpu() - torch.from_numpy(want)).norm(dim=1) / torch.from_numpy(want).norm(dim=1).clamp_min(1e-30)
    assert float(err.max()) <= 5e-5 and float(err.median()) <= TOL


def test_resident_rows_at_wildly_different_scales(dev):
    # every row is its own lasso problem and is rescaled on its own: the relative error of EACH
    # row stays at the float32 level although the batch spans 12 orders of magnitude
    n, d, k, iters = 600, 64, 256, 60
    fallbacks = _cabi.resident_fallbacks()
    x, w = make_problem(n, d, k, seed=5, kind="planted")
    g = torch.Generator().manual_seed(11)
    scale = 10.0 ** (12 * torch.rand(n, 1, generator=g) - 6)
    lr = 1.0 / oracle.lipschitz_constant(w)
    got = torch.empty(n, k)
    want = torch.empty(n, k)
    # alpha scales with the row (alpha is a scalar of the API: solve in groups of equal scale)
    for lo in range(0, n, 100):
        sl = slice(lo, lo + 100)
        s = float(scale[lo])
        xs = x[sl] * s
        want[sl] = oracle.ista(xs, torch.zeros(100, k), w, alpha=0.1 * s, lr=lr, maxiter=iters, tol=0.0)
        got[sl] = sparse_encode(xs.to(dev), w.to(dev), alpha=0.1 * s, lr=lr, maxiter=iters, tol=0.0,
                                path="resident").cpu()
    row_err = (got - want).double().norm(dim=1) / want.double().norm(dim=1).clamp_min(1e-300)
    assert float(row_err.max()) <= TOL
    # one batch that mixes the scales (a single alpha): still per-row accurate
    xm = x * scale
    wantm = oracle.ista(xm, torch.zeros(n, k), w, alpha=0.05, lr=lr, maxiter=iters, tol=0.0)
    gotm = sparse_encode(xm.to(dev), w.to(dev), alpha=0.05, lr=lr, maxiter=iters, tol=0.0, path="resident").cpu()
    live = wantm.double().norm(dim=1) > 0
    row_err = (gotm - wantm).double().norm(dim=1)[live] / wantm.double().norm(dim=1)[live]
    assert float(row_err.max()) <= TOL
    assert float(gotm[~live].abs().max() if (~live).any() else 0.0) == 0.0
    assert _cabi.resident_fallbacks() == fallbacks


def test_resident_falls_back_to_the_streaming_kernel(dev, monkeypatch):
    # LASSO_B200_RES_LIMIT lowers the operand bound at which the resident kernel gives up, so
    # that the hand-over (flag, stream sync, streaming bf16x3 solve from the intact z0) is exercised
    n, d, k = 300, 32, 128
    x, w = make_problem(n, d, k, seed=6)
    xd, wd = x.to(dev), w.to(dev)
    lr = 1.0 / oracle.lipschitz_constant(w)
    z0 = (0.05 * torch.randn(n, k)).to(dev)
    before = _cabi.resident_fallbacks()
    monkeypatch.setenv("LASSO_B200_RES_LIMIT", "1e-3")
    buf = z0.clone()
    got, done, _ = _cabi.fista_device(xd, wd, buf, 0.1, lr, 25, True, -1.0, path="resident",
                                      want_iters=True, out=buf)      # z0 aliases the output
    monkeypatch.delenv("LASSO_B200_RES_LIMIT")
    assert _cabi.resident_fallbacks() == before + 1 and done == 25
    stream, _, _ = _cabi.fista_device(xd, wd, z0, 0.1, lr, 25, True, -1.0, path="tcgen05")
    assert torch.equal(got, stream)
    # non-finite input: the resident kernel hands over as well instead of returning garbage silently
    xb = xd.clone()
    xb[3, 5] = float("inf")
    got, _, _ = _cabi.fista_device(xb, wd, None, 0.1, lr, 5, True, -1.0, path="resident")
    assert _cabi.resident_fallbacks() == before + 2
    assert bool(torch.isfinite(got[:3]).all()) and bool(torch.isfinite(got[4:]).all())


def test_host_pipeline_streams_waves_through_a_ring(dev, monkeypatch):
    # more waves than ring slots: chunk buffers are reused while earlier downloads are in flight;
    # the result must equal the device entry point bit for bit, with a warm start and with a stop
    # test that fires early (second pipelined pass with exactly that many iterations)
    n, d, k = 90000, 16, 32
    x, w = make_problem(n, d, k, seed=9)
    lr = 1.0 / oracle.lipschitz_constant(w)
    g = torch.Generator().manual_seed(3)
    z0 = 0.05 * torch.randn(n, k, generator=g)
    xd, wd = x.to(dev), w.to(dev)
    want, _, _ = _cabi.fista_device(xd, wd, z0.to(dev), 0.1, lr, 12, True, -1.0, path="resident")
    # pageable x / z0 / out: the staged three-stream pipeline
    got, _ = _cabi.fista_host(x, w, z0, 0.1, lr, 12, True, -1.0, path="resident")
    assert torch.equal(got, want.cpu())
    # pinned x / z0 / out: same pipeline without the driver's staging copy; with LASSO_B200_ZERO_COPY=1 the
    # kernel reads and writes the host buffers directly (one launch, no staging at all)
    out_pin = torch.empty(n, k).pin_memory()
    for zero_copy in (False, True):
        if zero_copy:
            monkeypatch.setenv("LASSO_B200_ZERO_COPY", "1")
        got, _ = _cabi.fista_host(x.pin_memory(), w, z0.pin_memory(), 0.1, lr, 12, True, -1.0, path="resident",
                                  out=out_pin)
        assert got is out_pin and torch.equal(got, want.cpu())
    monkeypatch.delenv("LASSO_B200_ZERO_COPY")
    tol_abs = float(np.float32(n * k * 1e-3))
    want, done_d, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 300, True, tol_abs, path="resident",
                                         want_iters=True)
    got, done_h = _cabi.fista_host(x, w, None, 0.1, lr, 300, True, tol_abs, path="resident", want_iters=True)
    assert 1 < done_d < 300 and done_h == done_d
    assert torch.equal(got, want.cpu())
    got, done_p = _cabi.fista_host(x.pin_memory(), w, None, 0.1, lr, 300, True, tol_abs, path="resident",
                                   want_iters=True, out=out_pin)
    assert done_p == done_d and torch.equal(got, want.cpu())
    # probe runs (maxiter >= 256 with a real tolerance: 64, 256, ... iterations before the full length) find the
    # same stopping iteration as one full-length run would
    want_it = oracle.ista(x[:4096], torch.zeros(4096, k), w, alpha=0.1, lr=lr, maxiter=2000,
                          tol=1e-3 * n / 4096, return_info=True)[1]
    assert done_d <= 300 and want_it >= 1


def test_resident_zero_threshold_stop_test(dev):
    # tol = 0 keeps the reference's stop test armed: it fires when an iteration changes nothing
    # (ista.py:93 with a threshold of 0).  The resident kernel records only "moved / did not move".
    n, d, k = 200, 16, 64
    x, w = make_problem(n, d, k, seed=7)
    xd, wd = x.to(dev), w.to(dev)
    lr = 1.0 / oracle.lipschitz_constant(w)
    z, done, _ = _cabi.fista_device(xd, wd, None, 1e3, lr, 9, True, 0.0, path="resident", want_iters=True)
    assert done == 1 and float(z.abs().max()) == 0.0
    z, done, _ = _cabi.fista_device(xd, wd, None, 0.1, lr, 9, True, 0.0, path="resident", want_iters=True)
    want = oracle.ista(x, torch.zeros(n, k), w, alpha=0.1, lr=lr, maxiter=9, tol=0.0)
    assert done == 9 and rel_fro(z, want) <= TOL


@pytest.mark.parametrize("kind", ["planted", "randn"])
def test_full_size_c2_against_row_subset_oracle(dev, kind):
    # BASELINE config 2: n=65536, d=64, k=256, alpha=0.1, 200 FISTA iterations, fp32
    n, d, k, alpha, iters = 65536, 64, 256, 0.1, 200
    x, w = make_problem(n, d, k, seed=0, kind=kind)
    lr = 1.0 / oracle.lipschitz_constant(w)
    z = sparse_encode(x.to(dev), w.to(dev), alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
    rows = torch.cat([torch.arange(0, 256), torch.arange(n // 2 - 100, n // 2 + 100),
                      torch.arange(n - 256, n)])
    want = oracle.ista(x[rows], torch.zeros(len(rows), k), w, alpha=alpha, lr=lr, maxiter=iters,
                       tol=0.0)
    got = z[rows.to(dev)].cpu()
    assert rel_fro(got, want) <= TOL
    assert support_mismatch(got, want) <= 2e-3
    # size-independent property: one more ISTA step from the result barely moves it, and the
    # objective is no worse than the oracle's on the subset
    obj = lambda zz: float(oracle.lasso_loss(x[rows], zz, w, alpha))
    assert obj(got) <= obj(want) * (1 + 1e-5)


def test_kkt_conditions_after_convergence(dev):
    x, w = make_problem(512, 32, 96, seed=4, kind="planted")
    alpha = 0.05
    z = sparse_encode(x.to(dev), w.to(dev), alpha=alpha, maxiter=3000, tol=0.0).cpu().double()
    grad = (z @ w.double().T - x.double()) @ w.double()
    on = z != 0
    assert float((grad[on] + alpha * torch.sign(z[on])).abs().max()) <= 2e-5
    assert float(grad[~on].abs().max()) <= alpha * (1 + 1e-4)


@pytest.mark.parametrize("d,k", [(64, 256), (10, 50), (300, 40), (128, 1024), (1, 1)])
def test_lipschitz_constant(dev, d, k):
    w = torch.randn(d, k, generator=torch.Generator().manual_seed(d * k))
    want = oracle.lipschitz_constant(w)
    got = lipschitz_constant(w.to(dev))
    assert abs(got - want) <= 1e-9 * want
    assert abs(lipschitz_constant(w) - want) <= 1e-9 * want   # CPU tensor is moved, not solved on CPU


def test_lr_auto_matches_pinned_lr(dev):
    g = load_golden("ista_planted_200")
    z = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"],
             maxiter=int(g["maxiter"]), tol=0.0)
    assert rel_fro(z, g["z"]) <= TOL


def test_loss_and_statistics(dev):
    g = load_golden("mstep")
    x, z, w = g["x"].to(dev), g["z"].to(dev), g["weight"].to(dev)
    loss = lasso_loss(x, z, w, g["alpha"])
    assert loss.dtype == torch.float32 and abs(float(loss) - g["loss"]) <= 2e-6 * abs(g["loss"])
    gzz, gzx = _cabi.gram(z, x)
    z64, x64 = g["z"].double(), g["x"].double()
    assert rel_fro(gzz, z64.T @ z64) <= 1e-6 and rel_fro(gzx, z64.T @ x64) <= 1e-6
    # larger, ragged, sparse codes
    xb, wb = make_problem(5000, 24, 70, seed=8)
    zb = sparse_encode(xb.to(dev), wb.to(dev), alpha=0.1, maxiter=30, tol=0.0)
    gzz, gzx = _cabi.gram(zb, xb.to(dev))
    z64 = zb.cpu().double()
    assert rel_fro(gzz, z64.T @ z64) <= 1e-6 and rel_fro(gzx, z64.T @ xb.double()) <= 1e-6
    want = oracle.lasso_loss(xb, zb.cpu(), wb, 0.1)
    assert abs(float(lasso_loss(xb.to(dev), zb, wb.to(dev), 0.1)) - float(want)) <= 1e-5 * float(want)


def test_update_dict_matches_reference(dev):
    g = load_golden("mstep")
    w = g["weight"].to(dev).clone()
    z = g["z"].to(dev).clone()
    out = update_dict(w, g["x"].to(dev), z)
    assert out.data_ptr() == w.data_ptr()          # in place, like the reference
    assert rel_fro(w, g["weight_update"]) <= TOL
    assert torch.equal(z.cpu(), g["z"])
    v = update_dict_ridge(g["x"].to(dev), g["z"].to(dev), lambd=g["lambd"])
    assert v.shape == g["weight_ridge"].shape and rel_fro(v, g["weight_ridge"]) <= TOL


@pytest.mark.parametrize("n,d,k", [(3000, 150, 300), (1500, 289, 300), (2000, 9, 600), (1500, 31, 77)])
def test_update_dict_large_dictionary(dev, n, d, k):
    # dictionaries beyond the single-SM sweep kernel (d * k floats do not fit its shared memory, k > 256,
    # or odd d / k): the atom sweep runs on a cluster of CTAs, each holding a slice of the rows
    x, w = make_problem(n, d, k, seed=12)
    z = sparse_encode(x.to(dev), w.to(dev), alpha=0.1, maxiter=25, tol=0.0)
    want = oracle.update_dict(w.clone(), x, z.cpu().clone())
    got = update_dict(w.to(dev).clone(), x.to(dev), z.clone())
    assert rel_fro(got, want) <= TOL


def test_update_dict_large_dictionary_degenerate(dev):
    # unused atoms in the cluster sweep: flagged, re-drawn with unit norm, their codes dropped; the others
    # match the oracle (which drops the same atoms)
    n, d, k = 1500, 150, 300
    x, w = make_problem(n, d, k, seed=13)
    z = sparse_encode(x.to(dev), w.to(dev), alpha=0.1, maxiter=25, tol=0.0)
    unused = [3, 150, 299]
    z[:, unused] = 0
    want = oracle.update_dict(w.clone(), x, z.cpu().clone())
    got = update_dict(w.to(dev).clone(), x.to(dev), z, random_seed=5)
    keep = [j for j in range(k) if j not in unused]
    assert rel_fro(got[:, keep], want[:, keep]) <= TOL
    for j in unused:
        assert abs(float(got[:, j].norm()) - 1.0) <= 1e-6


@pytest.mark.parametrize("n,d,k,positive", [(4096, 64, 256, False), (2000, 289, 300, False), (1500, 150, 300, True),
                                           (600, 7, 19, False), (3000, 700, 96, False)])
def test_blocked_sweep_equals_atom_by_atom_sweep(dev, monkeypatch, n, d, k, positive):
    """sweep_blk.cu (blocks of atoms: one product for the block, then a serial chain of norms) against the
    atom-by-atom kernels of aux_kernels.cu and the oracle's sequential update_dict, with unused atoms in the first,
    a middle and the last block."""
    x, w = make_problem(n, d, k, seed=21)
    xd, wd = x.to(dev), w.to(dev)
    z = sparse_encode(xd, wd, alpha=0.1, maxiter=20, tol=0.0)
    unused = sorted({1, k // 2, k - 1})
    z[:, unused] = 0
    gzz, gzx = _cabi.gram(z, xd)
    redraw = torch.randn(d, k, generator=torch.Generator().manual_seed(3)).to(dev)
    out = {}
    for mode in ("blocked", "legacy"):
        if mode == "legacy":
            monkeypatch.setenv("LASSO_B200_SWEEP", "legacy")
        wm, a, b = wd.clone(), gzz.clone(), gzx.clone()
        flags = _cabi.dict_update_gram(wm, a, b, redraw=redraw, positive=positive)
        out[mode] = (wm, flags, a, b)
    monkeypatch.delenv("LASSO_B200_SWEEP")
    assert torch.equal(out["blocked"][1], out["legacy"][1])
    assert sorted(torch.nonzero(out["blocked"][1]).flatten().tolist()) == unused
    assert rel_fro(out["blocked"][0], out["legacy"][0]) <= 2e-6
    assert torch.equal(out["blocked"][2], out["legacy"][2]) and torch.equal(out["blocked"][3], out["legacy"][3])
    want = oracle.update_dict(w.clone(), x, z.cpu().clone(), positive=positive)
    keep = [j for j in range(k) if j not in unused]
    assert rel_fro(out["blocked"][0][:, keep], want[:, keep]) <= TOL
    for j in unused:      # the replacement: the supplied draw, clamped if positive, unit norm
        r = redraw[:, j].clamp_min(0) if positive else redraw[:, j]
        assert rel_fro(out["blocked"][0][:, j], r / r.norm()) <= 1e-6


def test_update_dict_positive(dev):
    # positive=True (dict_learning.py:87-88) against the reference's own output, on both sweep kernels
    g = load_golden("mstep_positive")
    w = update_dict(g["weight"].to(dev).clone(), g["x"].to(dev), g["z"].to(dev).clone(), positive=True)
    assert rel_fro(w, g["weight_update"]) <= TOL and float(w.min()) >= 0.0
    n, d, k = 1500, 150, 300          # cluster sweep
    x, w0 = make_problem(n, d, k, seed=14)
    z = sparse_encode(x.to(dev), w0.to(dev), alpha=0.1, maxiter=25, tol=0.0)
    want = oracle.update_dict(w0.clone(), x, z.cpu().clone(), positive=True)
    got = update_dict(w0.to(dev).clone(), x.to(dev), z.clone(), positive=True)
    assert rel_fro(got, want) <= TOL and float(got.min()) >= 0.0


def test_update_dict_degenerate_atoms(dev):
    g = load_golden("mstep_degenerate")
    zero_atoms = [int(a) for a in g["zero_atoms"]]
    w = g["weight"].to(dev).clone()
    z = g["z"].to(dev).clone()
    z[:, zero_atoms[0]] = 0
    update_dict(w, g["x"].to(dev), z, random_seed=1234)
    keep = [j for j in range(w.size(1)) if j not in zero_atoms]
    assert rel_fro(w[:, keep], g["weight_update"][:, keep]) <= TOL
    for j in zero_atoms:   # re-drawn atoms: unit norm, codes dropped
        assert abs(float(w[:, j].norm()) - 1.0) <= 1e-6
        assert float(z[:, j].abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["constrained", "ridge"])
def test_dict_learning_matches_reference(dev, kind):
    # lr='auto' fixture: the reference's own step wobbles ~1e-6 run to run (ARPACK on a float32 Gram) and
    # EM amplifies it, hence the loose bound here; the pinned-step fixtures below carry the tight one
    g = load_golden("dict_learning_" + kind)
    torch.manual_seed(0)
    w, losses = dict_learning(g["x"], 50, alpha=g["alpha"], constrained=(kind == "constrained"),
                              steps=int(g["steps"]), lambd=g["lambd"], device="cpu", progbar=False,
                              algorithm="ista", maxiter=int(g["maxiter"]))
    assert w.shape == g["weight"].shape and not w.is_cuda and losses.shape == g["losses"].shape
    assert torch.allclose(losses, g["losses"], rtol=2e-4)
    assert rel_fro(w, g["weight"]) <= 5e-3
    loss = dict_evaluate(g["x"].to(dev), w.to(dev), g["alpha"], maxiter=int(g["maxiter"]))
    assert float(loss) == pytest.approx(float(oracle.lasso_loss(
        g["x"], oracle.sparse_encode(g["x"], w, g["alpha"], maxiter=int(g["maxiter"])), w,
        g["alpha"])), rel=1e-4)


@pytest.mark.parametrize("name", ["r2_dict_learning_pinned_constrained", "r2_dict_learning_pinned_ridge",
                                  "r2_dict_learning_persist_ridge_init"])
def test_dict_learning_pinned_step_matches_reference(dev, name):
    """Step pinned through **solver_kwargs (dict_learning.py:25,38): the reference is bit-reproducible, the
    CUDA path is held to 1e-5 on the dictionary and 1e-6 on the losses over the whole EM run."""
    g = load_golden(name)
    kw = dict(init="ridge", persist=True) if name.endswith("ridge_init") else {}
    torch.manual_seed(0)
    w, losses = dict_learning(g["x"], 50, alpha=g["alpha"], constrained=not name.endswith("pinned_ridge"),
                              steps=int(g["steps"]), lambd=g.get("lambd", 1e-2), device="cpu", progbar=False,
                              algorithm="ista", maxiter=int(g["maxiter"]), lr=g["lr"], **kw)
    assert torch.allclose(losses, g["losses"], rtol=1e-6, atol=0.0)
    assert rel_fro(w, g["weight"]) <= TOL


def test_cpu_tensors_through_the_dictionary_api(dev):
    """The reference workflow on CPU tensors: W, _ = dict_learning(X, k); dict_evaluate(Xtest, W, alpha);
    update_dict / update_dict_ridge / lasso_loss on CPU tensors (results on the inputs' device, in-place
    semantics of update_dict kept)."""
    g = load_golden("mstep")
    loss = lasso_loss(g["x"], g["z"], g["weight"], g["alpha"])
    assert not loss.is_cuda and abs(float(loss) - g["loss"]) <= 2e-6 * abs(g["loss"])
    w = g["weight"].clone()
    ret = update_dict(w, g["x"], g["z"].clone())
    assert ret is w and not w.is_cuda and rel_fro(w, g["weight_update"]) <= TOL
    v = update_dict_ridge(g["x"], g["z"], lambd=g["lambd"])
    assert not v.is_cuda and rel_fro(v, g["weight_ridge"]) <= TOL
    gd = load_golden("mstep_degenerate")
    w, z = gd["weight"].clone(), gd["z"].clone()
    z[:, 3] = 1.0            # stale codes of a degenerate atom must be cleared in the CALLER's tensor
    z_in = gd["z"].clone()
    update_dict(w, gd["x"], z_in)
    for j in [int(a) for a in gd["zero_atoms"]]:
        assert float(z_in[:, j].abs().max()) == 0.0 and abs(float(w[:, j].norm()) - 1.0) <= 1e-6
    torch.manual_seed(0)
    wl, _ = dict_learning(g["x"], 50, alpha=0.2, steps=3, progbar=False, maxiter=10, lr=0.05)
    loss = dict_evaluate(g["x"], wl, 0.2, maxiter=10, lr=0.05)
    want = oracle.lasso_loss(g["x"], oracle.sparse_encode(g["x"], wl, 0.2, maxiter=10, lr=0.05), wl, 0.2)
    assert not loss.is_cuda and float(loss) == pytest.approx(float(want), rel=1e-5)


def test_degenerate_redraw_consumes_the_generator_like_the_reference(dev):
    """dict_learning.py:92-93: one normal_() per degenerate atom, in atom order, on the dictionary's device."""
    gd = load_golden("mstep_degenerate")
    zero_atoms = [int(a) for a in gd["zero_atoms"]]
    w = gd["weight"].to(dev).contiguous()
    torch.manual_seed(99)
    update_dict(w, gd["x"].to(dev), gd["z"].to(dev).clone())
    torch.manual_seed(99)
    ref = gd["weight"].to(dev).clone()
    for j in zero_atoms:
        ref[:, j].normal_()                       # the reference's call, dict_learning.py:93
        assert torch.allclose(w[:, j], ref[:, j] / ref[:, j].norm(), rtol=1e-6, atol=1e-7)
    # on CPU tensors the draws come from the CPU generator, i.e. the reference's own stream: same atoms
    torch.manual_seed(1234)
    wc = update_dict(gd["weight"].clone(), gd["x"], gd["z"].clone())
    assert rel_fro(wc, gd["weight_update"]) <= TOL


def test_backtracking_matches_reference(dev):
    # ista.py:17-54: the step is re-searched from lr0 at every outer iteration
    g = load_golden("ista_backtrack")
    z = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"], fast=True,
             lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"], backtrack=True,
             eta_backtrack=g["eta_backtrack"])
    assert rel_fro(z, g["z"]) <= TOL
    # CPU tensors are moved, solved on the GPU and moved back
    zc = ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True, lr=g["lr"],
              maxiter=int(g["maxiter"]), tol=g["tol"], backtrack=True)
    assert not zc.is_cuda and torch.equal(zc, z.cpu())
    with pytest.raises(ValueError, match="eta must be > 1"):
        ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), lr=g["lr"], backtrack=True,
             eta_backtrack=0.5)


def test_backtracking_failure_warns_and_reverts(dev):
    # ista.py:39-52: eta so close to 1 that 1000 shrinks of a far too large step never reach F <= Q:
    # the reference warns and takes the iteration with the initial step.  Fixture from the reference.
    g = load_golden("r2_backtrack_failure")
    with pytest.warns(UserWarning, match="backtracking line search failed"):
        z = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"], fast=True,
                 lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"], backtrack=True,
                 eta_backtrack=g["eta_backtrack"])
    assert rel_fro(z, g["z"]) <= TOL
    # and the accepted-step path against the constant-step solver: with lr = 1/L the very first trial
    # is accepted, so backtrack=True must equal backtrack=False
    g = load_golden("ista_planted_200")
    kw = dict(alpha=g["alpha"], fast=True, lr=g["lr"], maxiter=12, tol=0.0)
    a = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), backtrack=True, **kw)
    b = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), path="ffma", **kw)
    assert rel_fro(a, b) <= 1e-6


def test_verbose_prints_the_reference_losses(dev, capsys):
    g = load_golden("ista_readme_fista")
    z = ista(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"], fast=True,
             lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"], verbose=True)
    printed = [float(l.split()[1]) for l in capsys.readouterr().out.splitlines() if l.startswith("loss:")]
    assert rel_fro(z, g["z"]) <= TOL
    # the reference prints the loss of the CURRENT iterate before each update (ista.py:80-81)
    zs, want = g["z0"], []
    for i in range(len(printed)):
        want.append(float(oracle.lasso_loss(g["x"], zs, g["weight"], g["alpha"])))
        zs = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], lr=g["lr"], maxiter=i + 1,
                         tol=0.0)
    assert len(printed) >= 1
    assert printed == pytest.approx(want, abs=6e-5)


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv2d_ista_matches_reference(dev, name):
    """lasso.conv2d.ista_conv2d on the k-blocked tcgen05 kernel (im2col -> linear, residual in image
    space) against the reference's outputs."""
    from lasso_b200.conv2d import ista_conv2d
    g = load_golden(name)
    lr = "auto" if g["lr"] < 0 else g["lr"]
    stride, padding = int(g.get("stride", 1)), int(g.get("padding", 0))
    z = ista_conv2d(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"], stride=stride,
                    padding=padding, fast=bool(g["fast"]), maxiter=int(g["maxiter"]), lr=lr, tol=g["tol"])
    assert z.shape == g["z"].shape and z.is_cuda
    assert rel_fro(z, g["z"]) <= TOL
    assert support_mismatch(z.cpu(), g["z"]) <= 2e-3
    # CPU tensors take the same path (H2D, solve, D2H)
    if name == "conv_8x8_plain":
        zc = ista_conv2d(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=False,
                         maxiter=int(g["maxiter"]), lr=lr, tol=g["tol"])
        assert not zc.is_cuda and torch.equal(zc, z.cpu())


def test_conv2d_config5_shape_against_oracle(dev):
    # BASELINE config 5: 28x28 images, 512 filters of 8x8 (batch cut to 24 images for the oracle)
    from lasso_b200.conv2d import ista_conv2d
    g = torch.Generator().manual_seed(5)
    n, filters, size, ks = 24, 512, 28, 8
    w = torch.randn(filters, 1, ks, ks, generator=g)
    w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
    o = size - ks + 1
    code = torch.randn(n, filters, o, o, generator=g) * (torch.rand(n, filters, o, o, generator=g) < 0.002)
    x = torch.nn.functional.conv_transpose2d(code, w) + 0.01 * torch.randn(n, 1, size, size, generator=g)
    lr, alpha, iters = 2e-3, 0.05, 20
    z0 = torch.zeros(n, filters, o, o)
    want = oracle.conv2d_ista(x, z0, w, alpha=alpha, fast=True, maxiter=iters, lr=lr, tol=0.0)
    got = ista_conv2d(x.to(dev), z0.to(dev), w.to(dev), alpha=alpha, fast=True, maxiter=iters, lr=lr, tol=0.0)
    assert rel_fro(got, want) <= TOL


def test_full_size_c3_against_row_subset_oracle(dev):
    """BASELINE config 3 at its full size (n = 262144, d = 128, k = 1024, alpha = 0.05) on the k-blocked
    tcgen05 kernel; rows are independent lasso problems, so a row subset is checked against the oracle."""
    n, d, k, alpha, iters = 262144, 128, 1024, 0.05, 30
    x, w = make_problem(n, d, k, seed=0, kind="planted")
    lr = 1.0 / oracle.lipschitz_constant(w)
    assert _cabi.select_path(n, d, k) == _cabi.PATH_BLOCKED
    z, _, _ = _cabi.fista_device(x.to(dev), w.to(dev), None, alpha, lr, iters, True, -1.0)
    rows = torch.cat([torch.arange(0, 96), torch.arange(n // 2, n // 2 + 64), torch.arange(n - 96, n)])
    want = oracle.ista(x[rows], torch.zeros(len(rows), k), w, alpha=alpha, lr=lr, maxiter=iters, tol=0.0)
    assert rel_fro(z[rows.to(dev)].cpu(), want) <= TOL
    del z


def test_full_size_c5_against_image_subset_oracle(dev):
    """BASELINE config 5 at its full size (16384 images 28x28, 512 filters 8x8 = 7.2 M patch rows, 14.8 GB
    of codes); images are independent, the first and last four are checked against the oracle."""
    from lasso_b200.conv2d import ista_conv2d
    g = torch.Generator().manual_seed(5)
    n, filters, size, ks = 16384, 512, 28, 8
    w = torch.randn(filters, 1, ks, ks, generator=g)
    w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
    x = torch.randn(n, 1, size, size, generator=g)
    o = size - ks + 1
    lr, alpha, iters = 2e-3, 0.05, 6
    xd, wd = x.to(dev), w.to(dev)
    z0 = torch.zeros(n, filters, o, o, device=dev)
    got = ista_conv2d(xd, z0, wd, alpha=alpha, fast=True, maxiter=iters, lr=lr, tol=0.0)
    del z0
    sel = torch.cat([torch.arange(0, 4), torch.arange(n - 4, n)])
    want = oracle.conv2d_ista(x[sel], torch.zeros(len(sel), filters, o, o), w, alpha=alpha, fast=True,
                              maxiter=iters, lr=lr, tol=0.0)
    assert rel_fro(got[sel.to(dev)].cpu(), want) <= TOL
    del got
    torch.cuda.empty_cache()


def test_out_buffer_is_validated(dev):
    g = load_golden("ista_ragged")
    xd, wd = g["x"].to(dev), g["weight"].to(dev)
    n, k = xd.size(0), wd.size(1)
    kw = dict(alpha=g["alpha"], lr=g["lr"], maxiter=5, tol=0.0)
    for bad in (torch.empty(n, k, dtype=torch.float64, device=dev), torch.empty(k, n, device=dev).T,
                torch.empty(n - 1, k, device=dev), torch.empty(n, k)):
        with pytest.raises(ValueError, match="out must"):
            ista(xd, torch.zeros(n, k, device=dev), wd, out=bad, **kw)
    with pytest.raises(ValueError, match="out must"):
        ista(g["x"], torch.zeros(n, k), g["weight"], out=torch.empty(n, k, device=dev), **kw)
    # host entry point, out aliases z0, early stop (second pass needs the start codes again)
    ge = load_golden("ista_warmstart")
    z0 = ge["z0"].clone()
    z = ista(ge["x"], z0, ge["weight"], alpha=ge["alpha"], lr=ge["lr"], maxiter=400, tol=1e-4, out=z0)
    want = oracle.ista(ge["x"], ge["z0"], ge["weight"], alpha=ge["alpha"], lr=ge["lr"], maxiter=400, tol=1e-4)
    assert z is z0 and rel_fro(z, want) <= TOL


def test_two_streams_and_two_threads_share_the_device_workspace(dev):
    """The per-device workspace (second code buffer, stop-test record, dictionary image) is shared: calls
    from different streams / host threads must be serialised by the library, not corrupt each other."""
    import threading
    ga, gb = load_golden("ista_planted_200"), load_golden("ista_randn_200")
    probs = [(g["x"].to(dev), g["weight"].to(dev), g) for g in (ga, gb)]
    streams = [torch.cuda.Stream(device=dev) for _ in probs]
    outs = [None, None]
    for rep in range(4):
        for i, (xd, wd, g) in enumerate(probs):
            with torch.cuda.stream(streams[i]):
                outs[i] = ista(xd, torch.zeros(xd.size(0), wd.size(1), device=dev), wd, alpha=g["alpha"],
                               lr=g["lr"], maxiter=int(g["maxiter"]), tol=0.0,
                               path="tcgen05" if rep % 2 else "auto")
    torch.cuda.synchronize()
    for out, (_, _, g) in zip(outs, probs):
        assert rel_fro(out, g["z"]) <= TOL
    results = {}

    def worker(i):
        xh, wh, g = probs[i][2]["x"], probs[i][2]["weight"], probs[i][2]
        for _ in range(3):
            results[i] = ista(xh, torch.zeros(xh.size(0), wh.size(1)), wh, alpha=g["alpha"], lr=g["lr"],
                              maxiter=int(g["maxiter"]), tol=0.0)
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for i, (_, _, g) in enumerate(probs):
        assert rel_fro(results[i], g["z"]) <= TOL


def _integration_stub_source():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "ctypes.CDLL" in b]
    assert len(stub) == 1
    return root, stub[0]


def test_integration_stub_runs_as_written(dev):
    """the ctypes stub a maintainer of the reference would paste into
    lasso/linear/solvers/ista.py, executed verbatim against the reference's fixtures."""
    import os
    root, src = _integration_stub_source()
    cwd = os.getcwd()
    os.chdir(root)                      # the stub loads the library by its repo-relative path
    try:
        ns = {}
        exec(compile(src, "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    for name in ("ista_planted_200", "ista_warmstart", "ista_earlystop"):
        g = load_golden(name)
        z = ns["ista"](g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"],
                       fast=bool(g["fast"]), lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"])
        assert rel_fro(z, g["z"]) <= TOL
    g = load_golden("ista_planted_200")
    z = ns["ista"](g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), alpha=g["alpha"], lr='auto',
                   maxiter=int(g["maxiter"]), tol=g["tol"])
    assert rel_fro(z, g["z"]) <= 5e-5       # lr='auto': the reference's own ARPACK value wobbles ~1e-6


def _conv_operator_lambda_max(w, imsize, stride, padding):
    """float64 dense reference: the matrix of conv2d on [cin, h, w] images, then eigvalsh of A^T A."""
    import torch.nn.functional as F
    cin, (h, wd) = w.shape[1], imsize
    eye = torch.eye(cin * h * wd, dtype=torch.float64).reshape(-1, cin, h, wd)
    cols = F.conv2d(eye, w.double(), stride=stride, padding=padding).flatten(1)      # row i = A e_i
    return float(torch.linalg.eigvalsh(cols @ cols.T)[-1])


@pytest.mark.parametrize("filters,cin,ks,size,stride,padding", [
    (16, 1, 8, 20, 1, 0), (12, 4, 3, 12, 1, 1), (16, 4, 4, 20, 2, 1), (8, 2, 5, 9, 1, 0)])
def test_conv2d_exact_lipschitz_constant(dev, filters, cin, ks, size, stride, padding):
    """lip_constant (lip_const.py:8-31) as a device power iteration, against the dense operator in float64."""
    from lasso_b200.conv2d import lip_constant
    g = torch.Generator().manual_seed(filters + ks)
    w = torch.randn(filters, cin, ks, ks, generator=g)
    w = w / w.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
    want = _conv_operator_lambda_max(w, (size, size), stride, padding)
    got = lip_constant(w.to(dev), (size, size), stride=stride, padding=padding)
    # a Rayleigh quotient: never above the eigenvalue; the top of a convolution's spectrum is a cluster, so
    # the eigenvector (float32) is the limit: ~1e-5 relative at worst
    assert want * (1 - 1e-4) <= got <= want * (1 + 1e-9)
    o = (size + 2 * padding - ks) // stride + 1
    got_t = lip_constant(w, (o, o), transpose=True, stride=stride, padding=padding, sqrt=True)   # CPU tensor in
    assert abs(got_t ** 2 - want) <= 1e-4 * want


@pytest.mark.parametrize("name", ["conv_4x4_stride2_pad1", "conv_8x8_fista", "conv_3x3_auto_earlystop"])
def test_conv2d_verbose_prints_the_reference_losses(dev, capsys, name):
    """verbose=True (lasso/conv2d/ista.py:21-24, 36-38): 'loss: %0.4f' of the iterate every executed iteration starts
    from -- against the oracle's iterates, one line per executed iteration (the early-stop fixture stops the prints)."""
    import torch.nn.functional as F
    from lasso_b200.conv2d import ista_conv2d, lip_bound_conv2d
    g = load_golden(name)
    stride, padding = int(g.get("stride", 1)), int(g.get("padding", 0))
    lr = g["lr"] if g["lr"] > 0 else 1 / float(lip_bound_conv2d(g["weight"], padding))
    kw = dict(alpha=g["alpha"], stride=stride, padding=padding, fast=bool(g["fast"]), lr=lr, tol=g["tol"])
    z = ista_conv2d(g["x"].to(dev), g["z0"].to(dev), g["weight"].to(dev), maxiter=int(g["maxiter"]), verbose=True, **kw)
    printed = [l for l in capsys.readouterr().out.splitlines() if l.startswith("loss: ")]
    assert rel_fro(z, g["z"]) <= TOL
    _, done = oracle.conv2d_ista(g["x"], g["z0"], g["weight"], maxiter=int(g["maxiter"]), return_iters=True, **kw)
    assert len(printed) == done
    for i, line in enumerate(printed):
        zi = g["z0"] if i == 0 else oracle.conv2d_ista(g["x"], g["z0"], g["weight"], maxiter=i, **dict(kw, tol=-1.0))
        x_hat = F.conv_transpose2d(zi, g["weight"], stride=stride, padding=padding)
        want = float((0.5 * (g["x"] - x_hat).pow(2).sum() + g["alpha"] * zi.abs().sum()) / g["x"].size(0))
        assert abs(float(line.split()[1]) - want) <= 1e-4 + 2e-5 * abs(want), (i, line, want)


def test_conv2d_lr_exact_for_even_kernels(dev):
    """lr='exact': config 5's 8x8 filters get an automatic step (lr='auto' raises for even kernels, like the
    reference: lip_const.py:101-102); the solve equals the oracle run with the same step."""
    from lasso_b200.conv2d import ista_conv2d, lip_constant
    g = load_golden("conv_8x8_fista")
    xd, wd, z0 = g["x"].to(dev), g["weight"].to(dev), g["z0"].to(dev)
    with pytest.raises(ValueError):
        ista_conv2d(xd, z0, wd, alpha=g["alpha"], lr='auto', maxiter=3)
    lr = 1 / (1.0001 * lip_constant(wd, xd.shape[-2:]))
    z = ista_conv2d(xd, z0, wd, alpha=g["alpha"], lr='exact', maxiter=15, tol=0.0)
    want = oracle.conv2d_ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True, maxiter=15, lr=lr, tol=0.0)
    assert rel_fro(z, want) <= TOL
    # the exact constant is what makes the step safe: the objective decreases monotonically under plain ISTA
    losses = []
    for it in (1, 5, 20):
        zi = ista_conv2d(xd, z0, wd, alpha=g["alpha"], lr='exact', fast=False, maxiter=it, tol=0.0).cpu()
        xh = torch.nn.functional.conv_transpose2d(zi, g["weight"])
        losses.append(float(0.5 * (g["x"] - xh).square().sum() + g["alpha"] * zi.abs().sum()))
    assert losses[0] > losses[1] > losses[2]


@pytest.mark.parametrize("n,d,k,alpha", [(300, 64, 256, 0.1), (77, 13, 37, 0.5), (200, 289, 300, 0.5),
                                         (150, 120, 40, 0.2), (64, 300, 90, 1.0), (5, 1, 1, 0.3)])
def test_fused_ridge_start_matches_reference_formula(dev, n, d, k, alpha):
    """initialize_code(mode='ridge') (sparse_encode.py:28-29, utils.py:28-40) in the library: against the
    oracle's restatement in float32 (the reference's arithmetic) and in float64."""
    x, w = make_problem(n, d, k, seed=n + d, kind="randn")
    got = lasso_b200.linear.initialize_code(x.to(dev), w.to(dev), alpha, "ridge")
    assert got.shape == (n, k) and got.is_cuda
    want32 = oracle.initialize_code(x, w, alpha, "ridge")
    gram = w.double().T @ w.double() + alpha * torch.eye(k, dtype=torch.float64)
    want64 = torch.linalg.solve(gram, w.double().T @ x.double().T).T
    # min(d, k) <= 64: float64 factorisation; beyond: float32 like the reference's own Cholesky (utils.py:34)
    assert rel_fro(got.cpu().double(), want64) <= (2e-6 if min(d, k) <= 64 else 1e-5)
    assert rel_fro(got, want32) <= TOL
    cpu = lasso_b200.linear.initialize_code(x, w, alpha, "ridge")           # CPU tensors in, CPU tensor out
    assert not cpu.is_cuda and torch.equal(cpu, got.cpu())
    # the library entry point by itself (what initialize_code uses for min(d, k) <= 64), any size up to 320
    lib = _cabi.ridge_init(x.to(dev), w.to(dev), alpha)
    assert rel_fro(lib.cpu().double(), want64) <= (2e-6 if min(d, k) <= 64 else 1e-5)


def test_fused_ridge_start_not_positive_definite(dev):
    x, w = make_problem(16, 8, 12, seed=2, kind="randn")
    with pytest.raises(RuntimeError, match="not positive definite"):          # utils.py:35-38
        lasso_b200.linear.initialize_code(x.to(dev), w.to(dev), -5.0, "ridge")


@pytest.mark.parametrize("n,d,k,density", [(131072, 64, 256, 0.08), (20000, 64, 256, 1.0), (9000, 24, 70, 0.1),
                                           (5000, 128, 200, 0.3), (4100, 10, 128, 0.5),
                                           (10000, 289, 300, 0.05),     # the notebook's shapes: 3 x 3 tiles of 128
                                           (6000, 130, 512, 0.1), (4096, 512, 129, 0.2)])
def test_tensor_core_gram_statistics(dev, monkeypatch, n, d, k, density):
    """K3 on tcgen05 (gram_tc.cu): Z^T Z and Z^T X against float64, and against the FFMA kernel it replaces for
    k <= 512, d <= 512.  The round-toward-zero accumulate of the tensor core is confined to runs of 16
    accumulations (two-level accumulation), which keeps the statistics within the 1e-6 they are tested to."""
    g = torch.Generator().manual_seed(n + k)
    z = torch.randn(n, k, generator=g) * (torch.rand(n, k, generator=g) < density) * 3.0
    z[:, 3] = z[:, 3].abs()                    # a same-sign column: the case a truncation bias would show in
    x = torch.randn(n, d, generator=g) * 0.25
    zd, xd = z.to(dev), x.to(dev)
    gzz, gzx = _cabi.gram(zd, xd)
    monkeypatch.setenv("LASSO_B200_GRAM", "ffma")
    fzz, fzx = _cabi.gram(zd, xd)
    monkeypatch.delenv("LASSO_B200_GRAM")
    z64, x64 = z.double(), x.double()
    wzz, wzx = z64.T @ z64, z64.T @ x64
    assert rel_fro(fzz, wzz) <= 1e-6 and rel_fro(fzx, wzx) <= 1e-6
    assert rel_fro(gzz, wzz) <= 1e-6 and rel_fro(gzx, wzx) <= 1e-6
    assert float((gzz.cpu() - wzz).diagonal().abs().max() / wzz.diagonal().abs().max()) <= 1e-6
    assert torch.equal(gzz, gzz.T)             # mirrored tiles: exactly symmetric
