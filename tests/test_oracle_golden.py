"""The oracle (oracle/ista_oracle.py) against the golden vectors produced by the real
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

import oracle
from conftest import CONV_CASES, SOLVER_CASES, load_golden
from lasso_b200.testing import rel_fro

# Same torch build + same machine gives bit-identical results; a different CPU can
# select other MKL kernels, so allow float32 summation-order noise.
TOL = 2e-6


@pytest.mark.parametrize("name", SOLVER_CASES)
def test_ista_matches_reference(name):
    g = load_golden(name)
    z = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=bool(g["fast"]),
                    lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"])
    assert rel_fro(z, g["z"]) <= TOL


def test_backtracking_matches_reference():
    g = load_golden("ista_backtrack")
    z = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True, lr=g["lr"],
                    maxiter=int(g["maxiter"]), tol=g["tol"], backtrack=True,
                    eta_backtrack=g["eta_backtrack"])
    assert rel_fro(z, g["z"]) <= TOL


def test_backtracking_failure_matches_reference():
    """ista.py:39-52: 1000 shrinks that never satisfy F <= Q -> warning + the initial step."""
    g = load_golden("r2_backtrack_failure")
    with pytest.warns(UserWarning, match="backtracking line search failed"):
        z = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True, lr=g["lr"],
                        maxiter=int(g["maxiter"]), tol=g["tol"], backtrack=True,
                        eta_backtrack=g["eta_backtrack"])
    assert rel_fro(z, g["z"]) <= TOL


@pytest.mark.parametrize("init", ["zero", "ridge", "transpose", "lstsq", "unif"])
def test_initialize_code(init):
    g = load_golden(("r2_encode_init_" if init in ("lstsq", "unif") else "encode_init_") + init)
    if init == "unif":
        torch.manual_seed(int(g["seed"]))
    z0 = oracle.initialize_code(g["x"], g["weight"], g["alpha"], init)
    assert rel_fro(z0, g["z0"]) <= TOL
    if init == "unif":
        torch.manual_seed(int(g["seed"]))
    z = oracle.sparse_encode(g["x"], g["weight"], g["alpha"], init=init, lr=g["lr"],
                             maxiter=int(g["maxiter"]), tol=g["tol"])
    assert rel_fro(z, g["z"]) <= TOL


def test_early_stop_iteration_count():
    g = load_golden("ista_earlystop")
    z, done, deltas = oracle.ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], fast=True,
                                  lr=g["lr"], maxiter=int(g["maxiter"]), tol=g["tol"],
                                  return_info=True)
    assert 1 < done < int(g["maxiter"])          # the stop test fired
    assert deltas[-1] <= g["z0"].numel() * g["tol"] < deltas[-2]
    assert rel_fro(z, g["z"]) <= TOL


def test_f64_gold_close_to_reference():
    g = load_golden("ista_planted_200")
    z64 = oracle.ista_f64(g["x"].numpy(), g["z0"].numpy(), g["weight"].numpy(), g["alpha"],
                          g["lr"], int(g["maxiter"]))
    # the float32 reference sits 5e-7..3e-6 from the float64 solution
    assert rel_fro(torch.from_numpy(z64), g["z"]) <= 1e-5


def test_mstep_matches_reference():
    g = load_golden("mstep")
    assert abs(float(oracle.lasso_loss(g["x"], g["z"], g["weight"], g["alpha"])) - g["loss"]) \
        <= 1e-6 * abs(g["loss"])
    w = oracle.update_dict(g["weight"].clone(), g["x"], g["z"].clone())
    assert rel_fro(w, g["weight_update"]) <= TOL
    w = oracle.update_dict_ridge(g["x"], g["z"], lambd=g["lambd"])
    assert rel_fro(w, g["weight_ridge"]) <= 1e-5


def test_gram_space_sweep_equals_sequential_sweep():
    g = load_golden("mstep")
    z64, x64 = g["z"].double(), g["x"].double()
    w, zeroed = oracle.update_dict_gram(g["weight"], z64.T @ z64, z64.T @ x64)
    assert zeroed == []
    assert rel_fro(w, g["weight_update"]) <= 5e-6


def test_positive_atoms_match_reference():
    # update_dict(positive=True): atoms clamped at zero before the norm (dict_learning.py:87-88)
    g = load_golden("mstep_positive")
    w = oracle.update_dict(g["weight"].clone(), g["x"], g["z"].clone(), positive=True)
    assert rel_fro(w, g["weight_update"]) <= TOL
    assert float(g["weight_update"].min()) >= 0.0
    z64, x64 = g["z"].double(), g["x"].double()
    w2, zeroed = oracle.update_dict_gram(g["weight"], z64.T @ z64, z64.T @ x64, positive=True)
    assert zeroed == [] and rel_fro(w2, g["weight_update"]) <= 5e-6


def test_degenerate_atoms():
    g = load_golden("mstep_degenerate")
    zero_atoms = [int(a) for a in g["zero_atoms"]]
    torch.manual_seed(1234)
    w = oracle.update_dict(g["weight"].clone(), g["x"], g["z"].clone())
    assert rel_fro(w, g["weight_update"]) <= TOL
    z64, x64 = g["z"].double(), g["x"].double()
    w2, zeroed = oracle.update_dict_gram(g["weight"], z64.T @ z64, z64.T @ x64)
    assert zeroed == zero_atoms
    keep = [j for j in range(w.size(1)) if j not in zero_atoms]
    assert rel_fro(w2[:, keep], g["weight_update"][:, keep]) <= 5e-6


@pytest.mark.parametrize("kind", ["constrained", "ridge"])
def test_dict_learning_matches_reference(kind):
    g = load_golden("dict_learning_" + kind)
    w, losses = oracle.dict_learning(g["x"], g["weight0"].size(1), alpha=g["alpha"],
                                     constrained=(kind == "constrained"), steps=int(g["steps"]),
                                     lambd=g["lambd"], weight0=g["weight0"],
                                     maxiter=int(g["maxiter"]))
    # lr='auto' differs (ARPACK float32 vs float64 eigvalsh): 1e-6 on lr, amplified by EM
    assert torch.allclose(losses, g["losses"], rtol=2e-4)
    assert rel_fro(w, g["weight"]) <= 5e-3


@pytest.mark.parametrize("name", ["r2_dict_learning_pinned_constrained", "r2_dict_learning_pinned_ridge",
                                  "r2_dict_learning_persist_ridge_init"])
def test_dict_learning_pinned_step_matches_reference(name):
    """With lr pinned (it travels through **solver_kwargs, dict_learning.py:25,38) the reference is
    bit-reproducible, and so must the restatement be."""
    g = load_golden(name)
    kw = dict(init="ridge", persist=True) if name.endswith("ridge_init") else {}
    w, losses = oracle.dict_learning(g["x"], g["weight0"].size(1), alpha=g["alpha"],
                                     constrained=not name.endswith("pinned_ridge"), steps=int(g["steps"]),
                                     lambd=g.get("lambd", 1e-2), weight0=g["weight0"],
                                     maxiter=int(g["maxiter"]), lr=g["lr"], **kw)
    assert torch.allclose(losses, g["losses"], rtol=1e-6)
    assert rel_fro(w, g["weight"]) <= 1e-5


def test_oracle_equals_the_reference_itself():
    """When the unmodified reference travelled with the repo (oracle/_ref, built by oracle/Makefile),
    run it side by side with the restatement on fresh inputs: same bits with the step pinned."""
    from oracle import ref_loader
    ref_ista = ref_loader.ista()
    if ref_ista is None:
        pytest.skip("oracle/_ref absent (run `make -C oracle` where /root/reference exists)")
    from lasso_b200.testing import make_problem
    for seed, (n, d, k, alpha, fast) in enumerate([(48, 12, 40, 0.1, True), (33, 7, 19, 0.3, False)]):
        x, w = make_problem(n, d, k, seed=100 + seed, kind="planted")
        lr = 1.0 / oracle.lipschitz_constant(w)
        want = ref_ista(x, torch.zeros(n, k), w, alpha=alpha, fast=fast, lr=lr, maxiter=30, tol=0.0)
        got = oracle.ista(x, torch.zeros(n, k), w, alpha=alpha, fast=fast, lr=lr, maxiter=30, tol=0.0)
        assert torch.equal(got, want)


def test_dict_learning_init_draw_is_the_references():
    g = load_golden("dict_learning_constrained")
    torch.manual_seed(0)
    w, _ = oracle.dict_learning(g["x"], 50, alpha=g["alpha"], steps=0)
    assert torch.equal(w, g["weight0"])


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv2d_ista_matches_reference(name):
    """lasso/conv2d/ista.py:7-49 -- the oracle's restatement against the reference's own outputs."""
    import lasso_b200
    from lasso_b200.conv2d import lip_bound_conv2d
    g = load_golden(name)
    lr = g["lr"]
    stride, padding = int(g.get("stride", 1)), int(g.get("padding", 0))
    if lr < 0:      # the case was generated with lr='auto': the Fourier bound must match the reference's
        bound = float(lip_bound_conv2d(g["weight"], padding))
        assert abs(bound - g["lip_bound"]) <= 1e-6 * g["lip_bound"]
        lr = 1 / bound
    z = oracle.conv2d_ista(g["x"], g["z0"], g["weight"], alpha=g["alpha"], stride=stride, padding=padding,
                           fast=bool(g["fast"]), maxiter=int(g["maxiter"]), lr=lr, tol=g["tol"])
    assert rel_fro(z, g["z"]) <= TOL


def test_conv2d_lip_bound_error_behaviour():
    import lasso_b200
    from lasso_b200.conv2d import ista_conv2d, lip_bound_conv2d
    with pytest.raises(ValueError):                      # lip_const.py:101-102: odd kernels only
        lip_bound_conv2d(torch.randn(4, 1, 8, 8), 0)
    with pytest.raises(ValueError):
        lip_bound_conv2d(torch.randn(4, 1, 3, 5), 0)
    with pytest.raises(NotImplementedError):             # ista.py:10-12
        ista_conv2d(torch.randn(1, 1, 9, 9), torch.zeros(1, 4, 4, 4), torch.randn(4, 1, 3, 3), stride=2)
    # (h + 2 padding - kh) not a multiple of the stride: conv_transpose2d(z) cannot have x's size (the reference
    # fails on `x_hat - x`, ista.py:18-19); a code tensor of the wrong grid is a ValueError before any GPU work
    with pytest.raises(RuntimeError):
        ista_conv2d(torch.randn(1, 1, 10, 10), torch.zeros(1, 4, 4, 4), torch.randn(4, 1, 3, 3), stride=2, lr=0.1)
    with pytest.raises(ValueError):
        ista_conv2d(torch.randn(1, 1, 9, 9), torch.zeros(1, 4, 7, 7), torch.randn(4, 1, 3, 3), stride=2, lr=0.1)
    z0 = torch.zeros(1, 4, 4, 4)
    assert ista_conv2d(torch.randn(1, 1, 9, 9), z0, torch.randn(4, 1, 3, 3), stride=2, lr=0.1, maxiter=0) is z0
    # a pair of equal integers is the integer (torch's conv2d convention); unequal pairs are not built
    assert ista_conv2d(torch.randn(1, 1, 9, 9), z0, torch.randn(4, 1, 3, 3), stride=(2, 2), padding=(0, 0), lr=0.1,
                       maxiter=0) is z0
    with pytest.raises(NotImplementedError):
        ista_conv2d(torch.randn(1, 1, 9, 9), z0, torch.randn(4, 1, 3, 3), stride=(2, 1), lr=0.1)
