"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference package from ``oracle/_ref``.

``oracle/_ref/lasso`` is a verbatim copy of ``/root/reference/lasso`` made by ``oracle/Makefile`` in
the build container (git-ignored; it travels to the GPU box with the repo snapshot).  The package
cannot be imported as shipped on scipy >= 1.12: ``lasso/linear/solvers/iterative_ridge.py:5`` imports
a private name that moved.  The shim below restores that one name and nothing else.

Used by ``bench.py`` (``--impl reference`` / ``cpu_baseline`` / ``gpu_eager_baseline``) and by
``tests/test_oracle_golden.py`` to check the oracle against the reference itself.  Never imported by
the product package.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "lasso", "linear", "solvers", "ista.py"))


def load():
    """Returns the reference's ``lasso`` package, or None when oracle/_ref is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    import scipy.optimize  # noqa: F401
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import scipy.optimize.optimize as legacy  # deprecated alias module
            from scipy.optimize._optimize import _status_message
        legacy._status_message = _status_message
    except Exception:  # pragma: no cover - other scipy layouts: the import below reports the problem
        pass
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import lasso  # noqa: F401  (the reference)
    import lasso.linear  # noqa: F401
    import lasso.linear.solvers.ista  # noqa: F401
    _cached = lasso
    return _cached


def ista():
    """The reference's ``lasso.linear.solvers.ista.ista`` or None."""
    pkg = load()
    if pkg is None:
        return None
    return sys.modules["lasso.linear.solvers.ista"].ista
