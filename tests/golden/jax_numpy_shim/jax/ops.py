"""The functional index updates of old jax.ops (removed upstream, used by the reference's RQS file)."""
import numpy as _np

index = _np.index_exp


def index_update(x, idx, y):
    out = _np.array(x, copy=True)
    out[idx] = y
    return out


def index_add(x, idx, y):
    out = _np.array(x, copy=True)
    out[idx] += _np.asarray(y, dtype=out.dtype)
    return out
