"""Mirror of ``lasso.linear`` (lasso/linear/__init__.py:1-5) for the ISTA path."""
from . import solvers  # noqa: F401
from . import utils  # noqa: F401
from .dict_learning import (dict_evaluate, dict_learning, lasso_loss,  # noqa: F401
                            update_dict, update_dict_ridge)
from .sparse_encode import initialize_code, sparse_encode  # noqa: F401
