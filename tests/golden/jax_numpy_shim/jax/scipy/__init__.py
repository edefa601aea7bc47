from . import special, linalg, stats  # noqa: F401
