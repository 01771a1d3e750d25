"""Multi-GPU (NCCL, one process per GPU) checks of the row-sharded paths.  Skipped on boxes
with fewer than two GPUs; the host-side logic is also covered on CPU by the gloo test in
tests/test_host_logic.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import load_golden
from lasso_b200.testing import make_problem, rel_fro

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from lasso_b200.linear import dict_learning, lasso_loss, sparse_encode
        group = dist.group.WORLD
        res = {}
        # (1) global stop rule on a sharded batch: shards differ in size on purpose
        g = load_golden("ista_earlystop")
        n = g["x"].size(0)
        cut = [0, 20, n] if world == 2 else [round(i * n / world) for i in range(world + 1)]
        rows = slice(cut[rank], cut[rank + 1])
        z = sparse_encode(g["x"][rows].to(dev), g["weight"].to(dev), alpha=g["alpha"], lr=g["lr"],
                          maxiter=int(g["maxiter"]), tol=g["tol"], group=group)
        res["z_stop"] = z.cpu()
        # (2) sharded loss equals the global loss
        loss = lasso_loss(g["x"][rows].to(dev), z, g["weight"].to(dev), g["alpha"], group=group)
        res["loss"] = float(loss)
        # (3) sharded dictionary learning: every rank ends with the same dictionary
        gd = load_golden("dict_learning_constrained")
        nd = gd["x"].size(0)
        rows = slice(rank * nd // world, (rank + 1) * nd // world)
        torch.manual_seed(0)
        w, losses = dict_learning(gd["x"][rows], 50, alpha=gd["alpha"], steps=int(gd["steps"]),
                                  device="cpu", progbar=False, group=group, algorithm="ista",
                                  maxiter=int(gd["maxiter"]))
        res["w"], res["losses"] = w, losses
        # (4) the same with the step pinned (bit-reproducible reference => tight bound) and a tolerance
        # that makes the GLOBAL stop test fire early in some EM steps (the redo path of the packed step)
        for key, name, tol in (("pin", "r2_dict_learning_pinned_constrained", 1e-5),
                               ("pin_ridge", "r2_dict_learning_pinned_ridge", 1e-5)):
            gp = load_golden(name)
            rows = slice(rank * nd // world, (rank + 1) * nd // world)
            torch.manual_seed(0)
            res[key] = dict_learning(gp["x"][rows], 50, alpha=gp["alpha"], constrained=key == "pin",
                                     steps=int(gp["steps"]), lambd=gp["lambd"], device="cpu", progbar=False,
                                     group=group, algorithm="ista", maxiter=int(gp["maxiter"]), lr=gp["lr"],
                                     tol=tol)
        torch.manual_seed(0)
        res["early"] = dict_learning(gd["x"][rows], 50, alpha=gd["alpha"], steps=4, device="cpu",
                                     progbar=False, group=group, maxiter=300, lr=0.05, tol=1e-3)
        torch.save(res, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_encode_and_dict_learning(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    g = load_golden("ista_earlystop")
    z = torch.cat([p["z_stop"] for p in parts])
    assert rel_fro(z, g["z"]) <= 1e-5                       # stopped at the reference's iteration
    want_loss = float(oracle.lasso_loss(g["x"], g["z"], g["weight"], g["alpha"]))
    assert parts[0]["loss"] == pytest.approx(want_loss, rel=1e-5)
    assert parts[0]["loss"] == parts[1]["loss"]
    gd = load_golden("dict_learning_constrained")
    assert torch.equal(parts[0]["w"], parts[1]["w"])         # replicated without a broadcast
    assert torch.allclose(parts[0]["losses"], gd["losses"], rtol=2e-4)
    assert rel_fro(parts[0]["w"], gd["weight"]) <= 5e-3
    for key, name in (("pin", "r2_dict_learning_pinned_constrained"), ("pin_ridge", "r2_dict_learning_pinned_ridge")):
        gp = load_golden(name)
        w, losses = parts[0][key]
        assert torch.equal(w, parts[1][key][0])
        assert torch.allclose(losses, gp["losses"], rtol=1e-6) and rel_fro(w, gp["weight"]) <= 1e-5
    # early global stop inside sharded EM steps: equals the unsharded oracle run with the same options
    torch.manual_seed(0)
    want_w, want_losses = oracle.dict_learning(gd["x"], 50, alpha=gd["alpha"], steps=4, maxiter=300, lr=0.05, tol=1e-3)
    w, losses = parts[0]["early"]
    assert torch.equal(w, parts[1]["early"][0])
    assert torch.allclose(losses, want_losses, rtol=1e-5) and rel_fro(w, want_w) <= 1e-4
