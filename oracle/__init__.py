"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ISTA/FISTA sparse-encode path.

Nothing in the product package (``pytorch-lasso_b200/``) imports this.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may use it, and there only as the checker or
as the timed CPU baseline.

Parity status: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/``.
"""
from .ista_oracle import (  # noqa: F401
    beta_schedule,
    conv2d_ista,
    dict_learning,
    initialize_code,
    ista,
    ista_f64,
    lasso_loss,
    lipschitz_constant,
    sparse_encode,
    update_dict,
    update_dict_gram,
    update_dict_ridge,
)
