"""ORACLE (test infrastructure): rational-quadratic spline path of the reference.

Line-by-line numpy restatement of flows/bijections/neural_splines.py (vendored, dead code upstream -- it
uses the removed jax.ops API), including its quirks: softmax/softplus applied twice by the
coupling layer (Q7), `+1e-6` on the last knot in searchsorted, identity tails.

PINNED BY THE REFERENCE'S OWN SOURCE: the reference holds no runnable test or golden vector for this path (its only test,
tests/test_bijections.py:136-138, imports a non-existent package) and JAX cannot be installed here, but the file needs array
primitives only: tests/golden/make_rqs_golden.py executes the unmodified neural_splines.py on a numpy stand-in for its jax
imports (float32 "x64 disabled" dtype rules) and stores inputs, layer weights, bin indices, knots and outputs in
tests/golden/ref_rqs_vectors.npz.  tests/test_rqs_reference_vectors.py: every bin index equal (forward and inverse), knots within
a few float32 ulp (> 50 % bit-identical), values float32-grade, coupling layer to 4e-6.  What the stand-in cannot reproduce is
XLA's own float32 exp and reduction order.  Further pins: bijectivity and the analytic derivative check in tests/test_oracle_rqs.py.
"""
from __future__ import annotations

import numpy as np

MIN_BIN_WIDTH = 1e-3      # neural_splines.py:6-8
MIN_BIN_HEIGHT = 1e-3
MIN_DERIVATIVE = 1e-3


def softmax(x):
    """jax.nn.softmax(axis=-1): exp(x - max) / sum, in the working precision of `x`.

    JAX leaves two things to XLA that decide the last bit of a float32 knot: the exp implementation and the order of the
    reduction.  This restatement pins both so that the float32 arithmetic is a SPECIFICATION the CUDA exact-bin path
    (WF_RQS_EXACT_BINS, csrc/rqs_device.cuh: knots_exact) reproduces operation by operation: exp is correctly rounded
    (evaluated in float64, rounded once to the working precision) and the sum runs sequentially left to right."""
    d = x - x.max(-1, keepdims=True)
    e = np.exp(d.astype(np.float64)).astype(x.dtype)
    s = np.cumsum(e, axis=-1, dtype=x.dtype)[..., -1:]
    return e / s


def softplus(x):
    """jax.nn.softplus = logaddexp(x, 0)."""
    return np.logaddexp(x, x.dtype.type(0))


def searchsorted(bin_locations, inputs, eps=1e-6):
    """neural_splines.py:11-13."""
    b = bin_locations.copy()
    b[..., -1] = b[..., -1] + b.dtype.type(eps)
    return np.sum(inputs[..., None] >= b, axis=-1) - 1


def _knots(un, lo, hi, min_bin):
    """neural_splines.py:98-107 (widths) / :111-120 (heights): -> (cum [.., K+1], sizes [.., K])."""
    dt = un.dtype.type
    K = un.shape[-1]
    w = softmax(un)
    w = dt(min_bin) + (dt(1) - dt(min_bin) * dt(K)) * w
    cw = np.cumsum(w, axis=-1, dtype=un.dtype)                      # sequential left-to-right
    cw = np.concatenate([np.zeros_like(cw[..., :1]), cw], axis=-1)
    cw = dt(hi - lo) * cw + dt(lo)
    cw[..., 0] = dt(lo)
    cw[..., -1] = dt(hi)
    return cw, cw[..., 1:] - cw[..., :-1]


def rqs(inputs, uw, uh, ud, inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0, return_bin=False):
    """RQS (neural_splines.py:74-184).  inputs [...], uw/uh [..., K], ud [..., K+1] (already padded)."""
    dt = inputs.dtype.type
    if inputs.size and (inputs.min() < left or inputs.max() > right):
        raise ValueError("Input outside domain")                     # :88-89
    K = uw.shape[-1]
    if MIN_BIN_WIDTH * K > 1.0 or MIN_BIN_HEIGHT * K > 1.0:
        raise ValueError("Minimal bin width/height too large for the number of bins")
    cw, w = _knots(uw, left, right, MIN_BIN_WIDTH)
    d = dt(MIN_DERIVATIVE) + softplus(ud)                            # :109
    ch, h = _knots(uh, bottom, top, MIN_BIN_HEIGHT)
    idx = searchsorted(ch if inverse else cw, inputs)[..., None]    # :122-125
    take = lambda a: np.take_along_axis(a, idx, -1)[..., 0]
    in_cw, in_w, in_ch, in_h = take(cw), take(w), take(ch), take(h)
    in_delta = take(h / w)
    in_d, in_d1 = take(d), take(d[..., 1:])
    if inverse:                                                      # :142-167
        a = (inputs - in_ch) * (in_d + in_d1 - 2 * in_delta) + in_h * (in_delta - in_d)
        b = in_h * in_d - (inputs - in_ch) * (in_d + in_d1 - 2 * in_delta)
        c = -in_delta * (inputs - in_ch)
        disc = np.square(b) - 4 * a * c
        assert (disc >= 0).all()
        root = (2 * c) / (-b - np.sqrt(disc))
        out = root * in_w + in_cw
        t1mt = root * (1 - root)
        den = in_delta + (in_d + in_d1 - 2 * in_delta) * t1mt
        num = np.square(in_delta) * (in_d1 * np.square(root) + 2 * in_delta * t1mt + in_d * np.square(1 - root))
        lad = -(np.log(num) - 2 * np.log(den))
    else:                                                            # :168-184
        theta = (inputs - in_cw) / in_w
        t1mt = theta * (1 - theta)
        numer = in_h * (in_delta * np.square(theta) + in_d * t1mt)
        den = in_delta + (in_d + in_d1 - 2 * in_delta) * t1mt
        out = in_ch + numer / den
        num = np.square(in_delta) * (in_d1 * np.square(theta) + 2 * in_delta * t1mt + in_d * np.square(1 - theta))
        lad = np.log(num) - 2 * np.log(den)
    if return_bin:
        return out, lad, idx[..., 0].astype(np.int32)
    return out, lad


def unconstrained_rqs(inputs, uw, uh, ud, inverse=False, tail_bound=1.0, return_bin=False):
    """unconstrained_RQS (neural_splines.py:16-71): identity + logabsdet 0 outside [-B, B]."""
    dt = inputs.dtype.type
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    out = np.where(inside, 0, inputs).astype(inputs.dtype)
    lad = np.zeros_like(inputs)
    bins = np.full(inputs.shape, -1, dtype=np.int32)
    pad = [(0, 0)] * (ud.ndim - 1) + [(1, 1)]
    udp = np.pad(ud, pad)
    const = np.log(np.exp(dt(1) - dt(MIN_DERIVATIVE)) - dt(1))
    udp[..., 0] = const
    udp[..., -1] = const
    if inside.any():
        r = rqs(inputs[inside], uw[inside, :], uh[inside, :], udp[inside, :], inverse=inverse,
                left=-tail_bound, right=tail_bound, bottom=-tail_bound, top=tail_bound, return_bin=True)
        out[inside], lad[inside], bins[inside] = r
    if return_bin:
        return out, lad, bins
    return out, lad


# --------------------------------------------------------------------------- coupling layer
def fcnn(params, x):
    """FCNN = Dense, Tanh, Dense, Tanh, Dense (neural_splines.py:187-188).  params = [(W,b)]*3."""
    (W1, b1), (W2, b2), (W3, b3) = params
    return np.tanh(np.tanh(x @ W1 + b1) @ W2 + b2) @ W3 + b3


def _raw_spline_params(params, cond, half, K, B):
    """neural_splines.py:258-262: reshape, array_split -> (K, K, K-1), softmax * 2B, softplus (Q7: applied again in RQS)."""
    dt = cond.dtype.type
    out = fcnn(params, cond).reshape(-1, half, 3 * K - 1)
    W, H, Dv = out[..., :K], out[..., K:2 * K], out[..., 2 * K:]
    return dt(2 * B) * softmax(W), dt(2 * B) * softmax(H), softplus(Dv)


def coupling_direct(f1, f2, x, K, B):
    """NeuralSplineCoupling.direct_fun (neural_splines.py:254-272)."""
    D = x.shape[1]
    idx = D // 2
    lower, upper = x[:, :idx], x[:, idx:]
    W, H, Dv = _raw_spline_params(f1, lower, D // 2, K, B)
    upper, ld = unconstrained_rqs(upper, W, H, Dv, inverse=False, tail_bound=B)
    log_det = ld.sum(1)
    W, H, Dv = _raw_spline_params(f2, upper, D // 2, K, B)
    lower, ld = unconstrained_rqs(lower, W, H, Dv, inverse=False, tail_bound=B)
    log_det = log_det + ld.sum(1)
    return np.concatenate([lower, upper], axis=1), log_det


def coupling_inverse(f1, f2, z, K, B):
    """NeuralSplineCoupling.inverse_fun (neural_splines.py:274-292)."""
    D = z.shape[1]
    idx = D // 2
    lower, upper = z[:, :idx], z[:, idx:]
    W, H, Dv = _raw_spline_params(f2, upper, D // 2, K, B)
    lower, ld = unconstrained_rqs(lower, W, H, Dv, inverse=True, tail_bound=B)
    log_det = ld.sum(1)
    W, H, Dv = _raw_spline_params(f1, lower, D // 2, K, B)
    upper, ld = unconstrained_rqs(upper, W, H, Dv, inverse=True, tail_bound=B)
    log_det = log_det + ld.sum(1)
    return np.concatenate([lower, upper], axis=1), log_det


def coupling_flow_direct(layers, x, K, B):
    """Serial(NeuralSplineCoupling * L).direct_fun.  layers = [(f1, f2), ...]."""
    ld = np.zeros(x.shape[0], dtype=x.dtype)
    for f1, f2 in layers:
        x, l = coupling_direct(f1, f2, x, K, B)
        ld = ld + l
    return x, ld


def coupling_flow_inverse(layers, z, K, B):
    ld = np.zeros(z.shape[0], dtype=z.dtype)
    for f1, f2 in reversed(layers):
        z, l = coupling_inverse(f1, f2, z, K, B)
        ld = ld + l
    return z, ld


def random_fcnn(rng, in_dim, hidden, out_dim, dtype=np.float32):
    """Weights N(0, 1/fan_in), biases 0."""
    g = lambda a, b: (rng.standard_normal((a, b)) / np.sqrt(a)).astype(dtype)
    z = lambda n: np.zeros(n, dtype=dtype)
    return [(g(in_dim, hidden), z(hidden)), (g(hidden, hidden), z(hidden)), (g(hidden, out_dim), z(out_dim))]
