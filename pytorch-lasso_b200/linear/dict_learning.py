"""Dictionary learning front end -- mirrors lasso/linear/dict_learning.py.

E-step = ``sparse_encode`` (the FISTA kernels).  M-step in Gram space: one pass
over (Z, X) produces A = Z^T Z and B = Z^T X.  When the batch is row-sharded over
a process group, ONE buffer crosses GPUs per EM step,

    [ A (k*k) | B (k*d) | sum (x - z W^T)^2, sum |z| | per-iteration stop-test sums (maxiter) ]

(float64, one all-reduce), after which every rank runs the same tiny atom sweep,
so the dictionary stays replicated without a broadcast.  The E-step's
batch-global stop test (ista.py:93) rides in the same buffer: the solve runs
with the test disabled, and only if the all-reduced sums show that the test
fired early is the step redone with that many iterations (a second all-reduce,
rare).  The global row count is reduced once per ``dict_learning`` call.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _cabi
from .solvers.ista import _abs_tolerance, _first_stop, lipschitz_constant
from .sparse_encode import initialize_code, sparse_encode
from .utils import default_device

__all__ = ["lasso_loss", "dict_evaluate", "dict_learning", "update_dict", "update_dict_ridge"]


# set to a dict by bench.py to collect CUDA-event pairs around the EM step's collective
PROFILE = None


def _world(group):
    return torch.distributed.get_world_size(group) if group is not None else 1


def _root(group):
    return torch.distributed.get_global_rank(group, 0) \
        if group is not torch.distributed.group.WORLD else 0


def _global_rows(n, device, group):
    if _world(group) == 1:
        return n
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    torch.distributed.all_reduce(cnt, group=group)
    return int(cnt.item())


def _on_device(*tensors):
    """CUDA copies of the tensors (the arithmetic always runs on the GPU) + whether any was moved."""
    if all(t.is_cuda for t in tensors):
        return tensors, False
    dev = next((t.device for t in tensors if t.is_cuda), None) or default_device()
    return tuple(t.to(dev) for t in tensors), True


def lasso_loss(X, Z, weight, alpha=1.0, group=None, n_global=None):
    """(0.5 |X - Z W^T|^2 + alpha |Z|_1) / n as a float32 scalar tensor (dict_learning.py:10-13).
    CPU tensors are moved to the GPU for the arithmetic; the result lives on X's device."""
    (Xd, Zd, Wd), _ = _on_device(X, Z, weight)
    terms = _cabi.loss_terms(Xd, Zd, Wd)
    if _world(group) > 1:
        torch.distributed.all_reduce(terms, group=group)
    n = n_global if n_global is not None else _global_rows(X.size(0), Xd.device, group)
    return ((0.5 * terms[0] + alpha * terms[1]) / n).to(torch.float32).to(X.device)


def dict_evaluate(X, weight, alpha, group=None, **kwargs):
    """Held-out loss: encode then ``lasso_loss`` (dict_learning.py:16-20)."""
    X = X.to(weight.device)
    Z = sparse_encode(X, weight, alpha, group=group, **kwargs) if group is not None \
        else sparse_encode(X, weight, alpha, **kwargs)
    return lasso_loss(X, Z, weight, alpha, group=group)


class _Packed:
    """The one buffer an EM step all-reduces: views into a single float64 device tensor."""

    def __init__(self, k, d, extra, device):
        self.k, self.d = k, d
        self.buf = torch.zeros(k * k + k * d + 2 + extra, dtype=torch.float64, device=device)
        self.gzz = self.buf[:k * k].view(k, k)
        self.gzx = self.buf[k * k:k * k + k * d].view(k, d)
        self.loss = self.buf[k * k + k * d:k * k + k * d + 2]
        self.hist = self.buf[k * k + k * d + 2:]

    def fill(self, X, Z, weight=None):
        """Local statistics of (X, Z) written in place (no concatenation copy); loss terms too when
        ``weight`` (the pre-update dictionary, dict_learning.py:39) is given."""
        _cabi.gram(Z, X, out_zz=self.gzz, out_zx=self.gzx)
        if weight is not None:
            _cabi.loss_terms(X, Z, weight, out=self.loss)

    def all_reduce(self, group):
        if _world(group) > 1:
            if PROFILE is not None:    # bench.py: CUDA events around the step's one collective
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.distributed.all_reduce(self.buf, group=group)
                e1.record()
                PROFILE.setdefault("all_reduce_events", []).append((e0, e1))
                PROFILE["all_reduce_bytes"] = self.buf.numel() * 8
            else:
                torch.distributed.all_reduce(self.buf, group=group)


def _statistics(X, Z, group):
    k, d = Z.shape[1], X.shape[1]
    packed = _Packed(k, d, 0, Z.device)
    packed.fill(X, Z)
    packed.all_reduce(group)
    return packed.gzz, packed.gzx


def _sweep_device(dictionary, gzz, gzx, Z, positive, eps):
    """Atom sweep on replicated statistics; the codes of degenerate atoms are cleared by a masked
    kernel (dict_learning.py:98).  Returns the int32 device mask of degenerate atoms; no read-back."""
    zeroed = _cabi.dict_update_gram(dictionary, gzz, gzx, eps=eps, redraw=None, positive=positive)
    _cabi.zero_columns(Z, zeroed)
    return zeroed


def _redraw(dictionary, zeroed, positive, group, draw_device=None):
    """Host part of the degenerate-atom branch (dict_learning.py:91-95): fresh N(0,1) atoms, one
    `normal_()` per atom, in atom order, on a column view of a [d,k] tensor on the device the caller's
    dictionary lives on -- the very call the reference's loop makes (`dictionary[:, k].normal_()`), so a
    seeded run consumes that device's generator the same way.  Deferring the draws to the end of the
    sweep changes nothing: a re-drawn atom has no codes left, so later atoms never see it."""
    idx = zeroed.nonzero().flatten().tolist()
    if not idx:
        return
    d, k = dictionary.shape
    scratch = torch.empty(d, k, device=draw_device if draw_device is not None else dictionary.device)
    for j in idx:
        scratch[:, j].normal_()
    atoms = scratch[:, idx].T.contiguous().to(dictionary.device)
    if positive:
        atoms.clamp_(0, None)      # dict_learning.py:94-95
    if _world(group) > 1:
        torch.distributed.broadcast(atoms, src=_root(group), group=group)
    dictionary[:, idx] = (atoms / atoms.norm(dim=1, keepdim=True)).T


def _sweep(dictionary, gzz, gzx, Z, positive, eps, group, draw_device=None):
    """Sweep + degenerate-atom branch with ONE flag read per sweep (the reference branches on the
    host once per atom)."""
    zeroed = _sweep_device(dictionary, gzz, gzx, Z, positive, eps)
    if bool(zeroed.any()):
        _redraw(dictionary, zeroed, positive, group, draw_device)
    return zeroed


def update_dict(dictionary, X, Z, random_seed=None, positive=False, eps=1e-10, group=None):
    """Block-coordinate atom update, in place on ``dictionary`` (and on ``Z`` for
    degenerate atoms) like the reference (dict_learning.py:56-103).  CPU tensors are accepted:
    the arithmetic runs on the GPU and the results are copied back into the caller's tensors."""
    if random_seed is not None:
        torch.manual_seed(random_seed)
    if dictionary.dtype != torch.float32:
        raise _cabi.LassoB200Error("dictionary must be a float32 tensor")
    (Dd, Xd, Zd), moved = _on_device(dictionary, X, Z)
    if not Dd.is_contiguous() or Dd is not dictionary:
        Dd = Dd.contiguous()
    Zc = Zd if Zd.is_contiguous() else Zd.contiguous()
    gzz, gzx = _statistics(Xd, Zc, group)
    zeroed = _sweep(Dd, gzz, gzx, Zc, positive, eps, group, draw_device=dictionary.device)
    if Dd is not dictionary:
        dictionary.copy_(Dd)
    if Zc is not Z:   # the caller's Z was moved or made contiguous: clear the columns there as well
        idx = zeroed.nonzero().flatten().to(Z.device)
        if idx.numel():
            Z[:, idx] = 0
    return dictionary


def _ridge_mstep(gzz, gzx, lambd, n):
    gzz.diagonal().add_(lambd * n)
    chol = torch.linalg.cholesky(gzz)
    return torch.cholesky_solve(gzx, chol).T.to(torch.float32).contiguous()


def update_dict_ridge(x, z, lambd=1e-4, group=None):
    """Unconstrained M-step: V = ((Z^T Z + lambd n I)^-1 Z^T X)^T (dict_learning.py:106-123)."""
    (xd, zd), _ = _on_device(x, z)
    gzz, gzx = _statistics(xd, zd.contiguous(), group)
    n = _global_rows(x.size(0), xd.device, group)
    return _ridge_mstep(gzz, gzx, lambd, n).to(x.device)


_FAST_KEYS = {'fast', 'lr', 'maxiter', 'tol', 'path', 'init'}


def dict_learning(X, n_components, alpha=1.0, constrained=True, persist=False,
                  lambd=1e-2, steps=60, device='cpu', progbar=True, group=None,
                  **solver_kwargs):
    """EM dictionary learning (dict_learning.py:23-53); returns ``(weight[d,k], losses[steps])``.

    ``device`` keeps its reference meaning for where the inputs/outputs live and
    which RNG draws the initial dictionary; the arithmetic runs on the current
    CUDA device either way.  With ``group`` each rank passes its row shard of X.
    """
    out_device = torch.device(device)
    compute = out_device if out_device.type == 'cuda' else default_device()
    n_features = X.shape[1]
    X = X.to(compute).contiguous()
    weight = torch.empty(n_features, n_components, device=out_device)
    torch.nn.init.orthogonal_(weight)
    if constrained:
        weight = F.normalize(weight, dim=0)
    weight = weight.to(compute).contiguous()
    sharded = _world(group) > 1
    if sharded:
        torch.distributed.broadcast(weight, src=_root(group), group=group)
    n_global = _global_rows(X.size(0), compute, group)      # once per call, not per step
    k, d = n_components, n_features

    # Constant-step solves (everything but backtrack / verbose) take the packed single-collective
    # step; the host-driven line search keeps its own per-trial reductions (ista.py:26-35).
    opts = dict(solver_kwargs)
    algorithm = opts.pop('algorithm', 'ista')
    packed_step = algorithm == 'ista' and set(opts) <= _FAST_KEYS
    maxiter = int(opts.get('maxiter', 10))
    tol = opts.get('tol', 1e-5)
    tol_abs = _abs_tolerance(n_global * k, tol)
    packed = _Packed(k, d, max(maxiter, 1) if sharded else 0, compute) if packed_step else None
    if group is not None and not packed_step:
        solver_kwargs = dict(solver_kwargs, group=group)

    Z0 = None
    losses = torch.zeros(steps, device=compute)

    try:
        from tqdm import tqdm
        bar = tqdm(total=steps, disable=not progbar)
    except ImportError:  # pragma: no cover
        bar = None
    for i in range(steps):
        if packed_step:
            # ---- E-step + statistics + loss, ONE all-reduce and ONE read-back per EM step ----
            z_start = Z0
            if z_start is None and opts.get('init', None) not in (None, 'zero'):
                z_start = initialize_code(X, weight, alpha, opts['init'])
            lr = opts.get('lr', 'auto')
            if lr == 'auto':
                lr = 1.0 / lipschitz_constant(weight)
            fast, path = opts.get('fast', True), opts.get('path', 'auto')
            w_before = weight.clone() if (sharded and constrained and maxiter > 1) else None
            run = maxiter
            for attempt in range(2):
                if maxiter == 0:
                    Z = z_start if z_start is not None else X.new_zeros(X.size(0), k)
                elif sharded:
                    # stop test disabled locally: its per-iteration sums travel with the statistics
                    spec = attempt == 0 and maxiter > 1
                    Z, _, hist = _cabi.fista_device(X, weight, z_start, alpha, lr, run, fast, -1.0,
                                                    path=path, want_hist=spec)
                    packed.hist.zero_()
                    if spec:
                        packed.hist[:run] = hist
                else:
                    Z, _, _ = _cabi.fista_device(X, weight, z_start, alpha, lr, run, fast, tol_abs, path=path)
                packed.fill(X, Z, weight)
                packed.all_reduce(group)
                losses[i] = ((0.5 * packed.loss[0] + alpha * packed.loss[1]) / n_global).to(torch.float32)
                # M-step on the replicated statistics, speculating that the global stop test did not
                # fire before maxiter; both questions the host has to answer are read back together
                flags = []
                if sharded and attempt == 0 and maxiter > 1:
                    flags.append((packed.hist[:maxiter - 1] <= tol_abs).any())
                if constrained:
                    zeroed = _sweep_device(weight, packed.gzz, packed.gzx, Z, False, 1e-10)
                    flags.append(zeroed.any())
                else:
                    weight_next = _ridge_mstep(packed.gzz, packed.gzx, lambd, n_global)
                flags = torch.stack(flags).tolist() if flags else []     # the step's one read-back
                early = bool(flags[0]) if (sharded and attempt == 0 and maxiter > 1) else False
                if early:
                    # rare: the global test fired at iteration `done` < maxiter (ista.py:93-95): redo the
                    # step with exactly that many iterations from the dictionary the E-step started with
                    run = _first_stop(packed.hist[:maxiter], tol_abs)
                    if constrained:
                        weight.copy_(w_before)
                    continue
                if constrained:
                    if flags and bool(flags[-1]):
                        _redraw(weight, zeroed, False, group, draw_device=out_device)
                else:
                    weight = weight_next
                break
        else:
            Z = sparse_encode(X, weight, alpha, Z0, **solver_kwargs)
            losses[i] = lasso_loss(X, Z, weight, alpha, group=group, n_global=n_global)
            gzz, gzx = _statistics(X, Z, group)
            if constrained:
                _sweep(weight, gzz, gzx, Z, False, 1e-10, group, draw_device=out_device)
            else:
                weight = _ridge_mstep(gzz, gzx, lambd, n_global)
        if persist:
            Z0 = Z
        if bar is not None and progbar:
            bar.set_postfix(loss=losses[i].item())
            bar.update(1)
    if bar is not None:
        bar.close()
    return weight.to(out_device), losses.to(out_device)
