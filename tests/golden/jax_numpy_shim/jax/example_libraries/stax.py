"""stax.Dense / Tanh / Relu / serial with the (init_fun, apply_fun) protocol of jax.example_libraries.stax, on numpy.
Every parameter set an init_fun creates is also appended to CREATED, so that the golden-vector script can store the weights
that the reference's closures captured."""
import numpy as _np

from .. import random as _random

CREATED = []


def Dense(out_dim):
    def init_fun(rng, input_shape):
        k1, k2 = _random.split(rng)
        fan_in = input_shape[-1]
        W = (_random.normal(k1, (fan_in, out_dim)) * _np.sqrt(2.0 / (fan_in + out_dim))).astype(_np.float32)   # glorot normal
        b = (_random.normal(k2, (out_dim,)) * 1e-2).astype(_np.float32)                                          # normal(1e-2)
        return input_shape[:-1] + (out_dim,), (W, b)

    def apply_fun(params, inputs, **kwargs):
        W, b = params
        return _np.dot(inputs, W) + b

    return init_fun, apply_fun


def _elementwise(fn):
    return (lambda rng, input_shape: (input_shape, ())), (lambda params, inputs, **kwargs: fn(inputs))


Tanh = _elementwise(_np.tanh)
Relu = _elementwise(lambda x: _np.maximum(x, _np.zeros((), dtype=x.dtype)))


def serial(*layers):
    inits, applies = zip(*layers)

    def init_fun(rng, input_shape):
        params = []
        for init in inits:
            rng, layer_rng = _random.split(rng)
            input_shape, p = init(layer_rng, input_shape)
            params.append(p)
        CREATED.append(params)
        return input_shape, params

    def apply_fun(params, inputs, **kwargs):
        for fn, p in zip(applies, params):
            inputs = fn(p, inputs)
        return inputs

    return init_fun, apply_fun
