"""names only (distributions.py imports them at module level; the spline models do not use them)"""
from types import SimpleNamespace as _NS

norm = _NS()
multivariate_normal = _NS()
uniform = _NS()
