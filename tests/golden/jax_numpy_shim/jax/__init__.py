"""numpy stand-in for the few jax modules the reference's RQS file imports (see ../README.md).  Test infrastructure."""
from . import numpy, nn, ops, random  # noqa: F401
from . import example_libraries  # noqa: F401
