"""BASELINE config 4: dict_learning, batch row-sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 tools/dl_bench.py [--rows-per-gpu 131072] [--steps 50] [--maxiter 100]

Every rank holds 131072 rows (n = 1 048 576 at G = 8), d = 64, k = 256, alpha = 0.1; E-step = 100
FISTA iterations (resident kernel, lr = 'auto' on device), M-step = Gram statistics, ONE NCCL
all-reduce of [Z^T Z | Z^T X | loss sums] (327 688 B) and the replicated atom sweep.
Prints one JSON line (rank 0): seconds per EM step (max over ranks, CUDA events), losses, and that
every rank ended with the same dictionary.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import lasso_b200
from lasso_b200.linear import dict_learning
from lasso_b200.testing import make_dictionary

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=131072)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--maxiter", type=int, default=100)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD if world > 1 else None
d, k, alpha = 64, 256, 0.1
# planted data from one hidden dictionary, a different code draw per rank
w_true = make_dictionary(d, k, seed=123).to(dev)
g = torch.Generator(device=dev).manual_seed(1000 + rank)
n = args.rows_per_gpu
code = torch.randn(n, k, generator=g, device=dev) * (torch.rand(n, k, generator=g, device=dev) < 0.05)
x = code @ w_true.T + 0.01 * torch.randn(n, d, generator=g, device=dev)
del code

def run(steps):
    torch.manual_seed(0)     # same initial dictionary draw on every rank (it is broadcast anyway)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w, losses = dict_learning(x, k, alpha=alpha, steps=steps, device=str(dev), progbar=False, group=group,
                              algorithm="ista", maxiter=args.maxiter, tol=0.0)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return w, losses, float(ms.item())

run(2)                                   # warm-up (workspace, NCCL channels)
w, losses, ms = run(args.steps)
same = True
if world > 1:
    ref = w.clone()
    dist.broadcast(ref, src=0)
    flag = torch.tensor([int(torch.equal(ref, w))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    same = bool(flag.item())
if rank == 0:
    print(json.dumps({
        "workload": "configs[3]: dict_learning n={} (={} x {} GPUs) d=64 k=256 alpha=0.1, {} EM steps x {} FISTA iterations".format(
            n * world, n, world, args.steps, args.maxiter),
        "n_gpus": world, "ms_per_em_step": ms / args.steps, "em_steps_per_s": args.steps / (ms * 1e-3),
        "fista_iters_per_s_aggregate": world * args.steps * args.maxiter / (ms * 1e-3),
        "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
        "dictionary_identical_on_all_ranks": same,
        "atom_recovery": float((w.T @ w_true).abs().max(dim=1).values.mean()),
        "launches": lasso_b200._cabi.launch_count()}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
