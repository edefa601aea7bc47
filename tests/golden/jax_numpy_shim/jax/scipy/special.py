import numpy as _np


def logsumexp(a, axis=None):
    m = _np.max(a, axis=axis, keepdims=True)
    return (_np.log(_np.sum(_np.exp(a - m), axis=axis, keepdims=True)) + m).squeeze(axis)
