"""Rational-quadratic splines and the coupling layer -- reference: flows/bijections/neural_splines.py.

`unconstrained_RQS` keeps the reference signature (neural_splines.py:16-26); the arithmetic runs in wf_rqs_apply.
"""
from __future__ import annotations

import torch

from .. import _ffi
from .._ffi import check, f32, lib, ptr, stream_ptr

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


def unconstrained_RQS(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives, inverse=False,
                      tail_bound=1.0, min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                      min_derivative=DEFAULT_MIN_DERIVATIVE, return_bin_idx=False):
    """inputs [...], widths/heights [..., K], derivatives [..., K-1] -> (outputs, logabsdet) (+ int32 bin index)."""
    if (min_bin_width, min_bin_height, min_derivative) != (1e-3, 1e-3, 1e-3):
        raise _ffi.WaveflowB200Error("only the reference's default minimum bin width/height/derivative (1e-3) are compiled in")
    x = f32(inputs)
    shape = x.shape
    K = unnormalized_widths.shape[-1]
    if unnormalized_heights.shape[-1] != K or unnormalized_derivatives.shape[-1] != K - 1:
        raise _ffi.WaveflowB200Error("widths/heights need K entries and derivatives K-1 (neural_splines.py:33-42)")
    xf = x.reshape(-1)
    uw = f32(unnormalized_widths).reshape(-1, K)
    uh = f32(unnormalized_heights).reshape(-1, K)
    ud = f32(unnormalized_derivatives).reshape(-1, K - 1)
    M = xf.shape[0]
    if uw.shape[0] != M or uh.shape[0] != M or ud.shape[0] != M:
        raise _ffi.WaveflowB200Error("parameter batch shape does not match inputs")
    out = torch.empty_like(xf)
    lad = torch.empty_like(xf)
    bins = torch.empty(M, dtype=torch.int32, device=xf.device) if return_bin_idx else None
    st = lib.wf_rqs_apply(ptr(xf), ptr(uw), ptr(uh), ptr(ud), M, K, float(tail_bound), int(bool(inverse)), ptr(out),
                          ptr(lad), ptr(bins), stream_ptr())
    check(st, "wf_rqs_apply")
    if return_bin_idx:
        return out.reshape(shape), lad.reshape(shape), bins.reshape(shape)
    return out.reshape(shape), lad.reshape(shape)
