"""numpy stand-in for the few jax modules the reference's RQS and spline files import (see ../README.md).  Test infrastructure."""
from . import numpy, nn, ops, random, lax, scipy  # noqa: F401
from . import example_libraries  # noqa: F401
from ._core import jit, vmap, grad, custom_jvp, hessian, config, value_and_grad, tree_map  # noqa: F401
