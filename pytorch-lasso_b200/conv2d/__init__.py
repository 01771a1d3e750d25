"""Mirror of ``lasso.conv2d`` (lasso/conv2d/ista.py, lip_const.py) on the B200 kernels."""
from .ista import ista_conv2d  # noqa: F401
from .lip_const import lip_bound_conv2d, lip_constant  # noqa: F401
