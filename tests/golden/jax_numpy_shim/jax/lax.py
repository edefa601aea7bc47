"""jax.lax.cond / while_loop as plain Python control flow (float32 loop state, as under jit with x64 disabled)."""
import numpy as _np


def cond(pred, true_fn, false_fn, *operands):
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def _f32(x):
    return _np.float32(x) if isinstance(x, (float, _np.floating)) else x


def while_loop(cond_fun, body_fun, init_val):
    state = tuple(_f32(v) for v in init_val) if isinstance(init_val, tuple) else _f32(init_val)
    while bool(cond_fun(state)):
        state = body_fun(state)
        state = tuple(_f32(v) for v in state) if isinstance(state, tuple) else _f32(state)
    return state
