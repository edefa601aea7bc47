"""Serial and Reverse -- reference: flows/bijections/bijections.py:317-347,417-467 (same closure protocol).

`Serial` recognises the layer pattern model_factory builds -- [BoxTransformLayer]? + (IMADE, Reverse) x L with identical
IMADE settings -- and then evaluates the whole stack with ONE fused kernel launch (wf_live_forward); any other
composition is evaluated layer by layer through the layers' own functions.
"""
from __future__ import annotations

import torch

from .. import _live
from .._ffi import f32


def Reverse():
    def init_fun(rng, input_dim, **kwargs):
        def direct_fun(params, inputs, **kwargs):
            x = f32(inputs)
            return torch.flip(x, dims=[1]), torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)

        def inverse_fun(params, inputs, **kwargs):
            x = f32(inputs)
            return torch.flip(x, dims=[1]), torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)

        direct_fun.wf_layer = ("reverse", None)
        return (), direct_fun, inverse_fun

    return init_fun


def _split_rng(rng):
    """Deterministic child generators (stand-in for jax.random.split)."""
    if isinstance(rng, torch.Generator):
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=rng).item())
    else:
        seed = int(rng if rng is not None else 0)
    g1, g2 = torch.Generator(), torch.Generator()
    g1.manual_seed((seed * 6364136223846793005 + 1442695040888963407) % (2 ** 63))
    g2.manual_seed((seed * 2862933555777941757 + 3037000493) % (2 ** 63))
    return g1, g2


def fuse_pattern(layers, input_dim):
    """[(kind, cfg)] -> LiveSpec (without prior) if the stack is [box]? + (imade, reverse) x L with one IMADE config."""
    kinds = [k for k, _ in layers]
    box = None
    i = 0
    if kinds and kinds[0] == "box":
        box = layers[0][1]; i = 1
    rest = layers[i:]
    if len(rest) % 2 or not rest:
        return None
    cfg0 = None
    for j in range(0, len(rest), 2):
        if rest[j][0] != "imade" or rest[j + 1][0] != "reverse":
            return None
        cfg = rest[j][1]
        if cfg0 is None:
            cfg0 = cfg
        elif any(cfg[k] != cfg0[k] for k in ("k", "n_knots", "reg", "tol", "left", "right", "grad_to_zero", "T")):
            return None
    if cfg0["grad_to_zero"]:
        return None
    spec = _live.LiveSpec(D=input_dim, n_layers=len(rest) // 2, tab_I=cfg0["tables"], k_I=cfg0["k"], reg=cfg0["reg"],
                          tol=cfg0["tol"], bc_I_left=cfg0["left"], bc_I_right=cfg0["right"],
                          box=None if box is None else box["box_side"], coord="mean" if box is None else box["coord"])
    return spec if spec.fusible() else None


def Serial(*init_funs):
    def init_fun(rng, input_dim, **kwargs):
        all_params, direct_funs, inverse_funs = [], [], []
        for f in init_funs:
            rng, layer_rng = _split_rng(rng)
            param, direct_fun, inverse_fun = f(layer_rng, input_dim)
            all_params.append(param); direct_funs.append(direct_fun); inverse_funs.append(inverse_fun)
        layers = [getattr(d, "wf_layer", ("opaque", None)) for d in direct_funs]
        spec = fuse_pattern(layers, input_dim)

        def feed_forward(params, apply_funs, inputs):
            x = f32(inputs)
            log_det = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
            for apply_fun, param in zip(apply_funs, params):
                x, ld = apply_fun(param, x)
                log_det = log_det + ld
            return x, log_det

        def direct_fun(params, inputs, **kwargs):
            if spec is None:
                return feed_forward(params, direct_funs, inputs)
            x = f32(inputs)
            w = _live.packed_for(spec, params, None, x.device)
            out = _live.forward(spec, w, x, want=("u", "logdet"))
            return out["u"], out["logdet"]

        def inverse_fun(params, inputs, **kwargs):
            if spec is None:
                return feed_forward(list(reversed(list(params))), list(reversed(inverse_funs)), inputs)
            x = f32(inputs)
            w = _live.packed_for(spec, params, None, x.device)
            return _live.inverse(spec, w, x), 0

        direct_fun.wf_spec = spec
        direct_fun.wf_layerwise = lambda params, inputs: feed_forward(params, direct_funs, inputs)
        inverse_fun.wf_layerwise = lambda params, inputs: feed_forward(list(reversed(list(params))),
                                                                        list(reversed(inverse_funs)), inputs)
        return all_params, direct_fun, inverse_fun

    return init_fun
