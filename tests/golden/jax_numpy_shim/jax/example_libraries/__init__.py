from . import stax, optimizers  # noqa: F401
