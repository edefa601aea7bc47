"""jax.numpy on numpy with JAX's default (x64 disabled) dtype canonicalisation: float64 -> float32 on the way in and out."""
import numpy as _np

from ._core import JArr as _JArr, jarr as _jarr, X64 as _X64

float32 = _np.float32
int32 = _np.int32
newaxis = None
pi = _np.pi
ndarray = _np.ndarray


def _canon(x):
    if _X64:
        return _jarr(x) if isinstance(x, _np.ndarray) else x
    if isinstance(x, (bool, int, _np.integer, _np.bool_)):
        return x
    if isinstance(x, float):
        return _np.float32(x)
    if isinstance(x, _np.floating):
        return _np.float32(x)
    if isinstance(x, _np.ndarray) and x.dtype == _np.float64:
        return _jarr(x.astype(_np.float32))
    if isinstance(x, _np.ndarray):
        return _jarr(x)
    if isinstance(x, (list, tuple)) and x and all(isinstance(v, _np.ndarray) for v in x):
        return type(x)(_canon(v) for v in x)
    return x


def _wrap(fn):
    def f(*args, **kwargs):
        out = fn(*[_canon(a) for a in args], **{k: _canon(v) for k, v in kwargs.items()})
        if isinstance(out, (list, tuple)):
            return type(out)(_canon(o) for o in out)
        return _canon(out)
    f.__name__ = getattr(fn, "__name__", "f")
    return f


_DEFAULT = _np.float64 if _X64 else _np.float32


def zeros(shape, dtype=None):
    return _jarr(_np.zeros(shape, dtype=_DEFAULT if dtype is None else dtype))


def ones(shape, dtype=None):
    return _jarr(_np.ones(shape, dtype=_DEFAULT if dtype is None else dtype))


def array(x, dtype=None):
    return _canon(_np.array(x, dtype=dtype))


asarray = array
zeros_like = _wrap(_np.zeros_like)
ones_like = _wrap(_np.ones_like)


def __getattr__(name):          # everything else: the numpy function of the same name, canonicalised
    obj = getattr(_np, name)
    return _wrap(obj) if callable(obj) and not isinstance(obj, type) else obj
