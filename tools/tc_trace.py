"""Run one C2-sized solve with LASSO_B200_TRACE and print block 0's per-warp timeline."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("LASSO_B200_TRACE", "/tmp/tc_trace.txt")
import torch
import lasso_b200, oracle
from lasso_b200 import _cabi
from lasso_b200.testing import make_problem
n, d, k = 65536, 64, 256
x, w = make_problem(n, d, k, seed=0)
lr = 1.0 / oracle.lipschitz_constant(w)
dev = torch.device("cuda", 0)
kpath = os.environ.get("TRACE_PATH", "tcgen05")   # "resident": iterations 3 and 4 of block 0's first tile
_cabi.fista_device(x.to(dev), w.to(dev), None, 0.1, lr, 6, True, -1.0, path=kpath)
torch.cuda.synchronize()
ev = collections.defaultdict(list)
for line in open(path):
    wi, t, e = line.split(); ev[int(wi)].append((int(t), int(e)))
t0 = min(t for v in ev.values() for t, _ in v)
names = {1: "P:empty_ok", 10: "M:tile", 11: "M:aready_ok", 12: "M:commit_chunk", 13: "M:commit_rfull", 14: "M:rready_ok",
         15: "M:gfree_ok", 16: "M:commit_g", 20: "C:full_ok", 21: "C:math_done", 22: "C:sfree_ok", 23: "C:st_done",
         30: "C:rfull_ok", 31: "C:phaseB_done", 40: "C:gfull_ok", 41: "C:epi_done", 50: "C:tile_end"}
if kpath == "resident":
    # MMA warp: 18 iteration top, 14 r pieces ready, 15 G buffer free (before a GEMM2 chunk is issued), 16 GEMM2
    # chunk committed, 11 pieces of a chunk ready, 12 GEMM1 slice committed.  Compute warps (0 / 4: set A, 8 / 12:
    # set B): 30 R complete, 31 phase B done, 40 G of the chunk complete, 42 sub-step's G in registers, 41 sub-step's
    # arithmetic done, 22 piece slot free, 23 pieces stored and signalled
    names = {11: "M:aready_ok", 12: "M:commit_chunk", 14: "M:rready_ok", 15: "M:gfree_ok",
             16: "M:commit_g", 22: "A:stage_free", 23: "A:arrived",
             30: "B:rfull_ok", 31: "B:done", 40: "C:gfull_ok", 41: "C:chunk_done", 42: "C:gfree_arrived",
             18: "M:iter_top"}
for wi in [int(v) for v in os.environ.get("TRACE_WARPS", "0,2,4,12" if kpath != "resident" else "16,0,4,8,12").split(",")]:
    print("---- warp", wi)
    prev = None
    for t, e in ev[wi][:int(os.environ.get("TRACE_ROWS", "70"))]:
        print("%8d  +%6d  %s" % (t - t0, (t - prev) if prev else 0, names.get(e, e)))
        prev = t
