from . import stax  # noqa: F401
