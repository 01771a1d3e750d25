"""Mirror of ``lasso.linear.solvers`` (solvers/__init__.py:1-8): only ``ista`` is in scope."""
from .ista import ista, lipschitz_constant  # noqa: F401
