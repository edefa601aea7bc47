"""Deterministic numpy PRNG behind the jax.random names the reference uses (keys are plain integers).  The streams differ
from JAX's; the drawn parameters are stored in the fixture, so only determinism matters."""
import numpy as _np


def PRNGKey(seed):
    return int(seed)


def split(key, num=2):
    ss = _np.random.SeedSequence(int(key)).spawn(num)
    return [int(s.generate_state(1)[0]) for s in ss]


def normal(key, shape, dtype=_np.float32):
    return _np.random.default_rng(int(key)).standard_normal(shape).astype(dtype)


def uniform(key, shape, dtype=_np.float32, minval=0.0, maxval=1.0):
    return _np.random.default_rng(int(key)).uniform(minval, maxval, shape).astype(dtype)
