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
                      min_derivative=DEFAULT_MIN_DERIVATIVE, return_bin_idx=False, exact_bins=False):
    """inputs [...], widths/heights [..., K], derivatives [..., K-1] -> (outputs, logabsdet) (+ int32 bin index).

    exact_bins=True (WF_RQS_EXACT_BINS): knot positions in exactly the reference's float32 operation sequence
    (neural_splines.py:98-107), bin indices bit-identical to it; the default is the fast ex2.approx softmax."""
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
    st = lib.wf_rqs_apply(ptr(xf), ptr(uw), ptr(uh), ptr(ud), M, K, float(tail_bound), int(bool(inverse)) | (2 if exact_bins else 0), ptr(out),
                          ptr(lad), ptr(bins), stream_ptr())
    check(st, "wf_rqs_apply")
    if return_bin_idx:
        return out.reshape(shape), lad.reshape(shape), bins.reshape(shape)
    return out.reshape(shape), lad.reshape(shape)


# ---------------------------------------------------------------------------------------------------- coupling layer
def _pack_fcnn(net, half: int, K: int, Hd: int, device) -> torch.Tensor:
    """stax.serial(Dense, Tanh, Dense, Tanh, Dense) params -> the chunked layout of wf_rqs_coupling_flow."""
    (W1, b1), _, (W2, b2), _, (W3, b3) = net
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
    W1, b1, W2, b2, W3, b3 = t(W1), t(b1), t(W2), t(b2), t(W3), t(b3)
    if W1.shape != (half, Hd) or W2.shape != (Hd, Hd) or W3.shape != (Hd, (3 * K - 1) * half):
        raise _ffi.WaveflowB200Error(f"unexpected FCNN shapes {tuple(W1.shape)} {tuple(W2.shape)} {tuple(W3.shape)}")
    KP = 8 if K <= 8 else 32
    head = torch.cat([W1.reshape(-1), b1, W2.reshape(-1), b2])
    pad = (-head.numel()) % 4
    parts = [head, torch.zeros(pad, dtype=torch.float32, device=device)]
    W3 = W3.reshape(Hd, half, 3 * K - 1)                                    # out.reshape(-1, dim//2, 3K-1)  (:258)
    b3 = b3.reshape(half, 3 * K - 1)
    for j in range(half):
        Wj = torch.zeros(Hd, 3 * KP, dtype=torch.float32, device=device)
        bj = torch.zeros(3 * KP, dtype=torch.float32, device=device)
        for blk, (lo, n) in enumerate([(0, K), (K, K), (2 * K, K - 1)]):     # array_split -> (K, K, K-1)  (:259)
            Wj[:, blk * KP:blk * KP + n] = W3[:, j, lo:lo + n]
            bj[blk * KP:blk * KP + n] = b3[j, lo:lo + n]
        parts += [Wj.reshape(-1), bj]
    return torch.cat(parts)


_COUPLING_PACKS = _ffi.PackCache(size=8)


def pack_coupling(layers, D: int, K: int, hidden_dim: int, device) -> torch.Tensor:
    """All 2L conditioners of Serial(NeuralSplineCoupling * L) in the layout of wf_rqs_coupling_flow, cached on the identity /
    version of the parameter leaves (the ~40 small torch ops per conditioner run once per parameter set, not per call)."""
    key = (D, K, hidden_dim, str(device), _ffi.params_key(layers))
    return _COUPLING_PACKS.get(key, layers, lambda: torch.cat(
        [_pack_fcnn(f, D // 2, K, hidden_dim, device) for pair in layers for f in pair]).contiguous())


def coupling_flow(layers, x, K: int, B: float, hidden_dim: int, inverse: bool = False):
    """Serial(NeuralSplineCoupling * L) in one launch.  layers = [(f1_params, f2_params), ...] -> (y, log_det)."""
    x = f32(x)
    N, D = x.shape
    w = pack_coupling(layers, D, K, hidden_dim, x.device)
    expect = lib.wf_rqs_coupling_net_floats(D, K, hidden_dim) * 2 * len(layers)
    if w.numel() != expect:
        raise _ffi.WaveflowB200Error("packed coupling weights have the wrong size")
    y = torch.empty_like(x)
    ld = torch.empty(N, dtype=torch.float32, device=x.device)
    st = lib.wf_rqs_coupling_flow(ptr(w), len(layers), D, K, hidden_dim, float(B), int(bool(inverse)), ptr(x), N, ptr(y),
                                  ptr(ld), stream_ptr())
    check(st, "wf_rqs_coupling_flow")
    return y, ld


def _tf32_planes(x: torch.Tensor):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(lib.wf_tf32_split(ptr(x), x.numel(), ptr(hi), ptr(lo), stream_ptr()), "wf_tf32_split")
    return hi, lo


def pack_fcnn_tc(net, device) -> torch.Tensor:
    """FCNN(32 -> 512 -> 512 -> 32*191) params -> the layout of wf_rqs_coupling_flow_tc (D = 64, K = 64, hidden 512)."""
    (W1, b1), _, (W2, b2), _, (W3, b3) = net
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device).contiguous()
    W1, b1, W2, b2, W3, b3 = t(W1), t(b1), t(W2), t(b2), t(W3), t(b3)
    K, half, Hd, NT = 64, 32, 512, 192
    if W1.shape != (half, Hd) or W2.shape != (Hd, Hd) or W3.shape != (Hd, (3 * K - 1) * half):
        raise _ffi.WaveflowB200Error("the tensor-core coupling path is built for D = 64, K = 64, hidden_dim = 512")
    W3 = W3.reshape(Hd, half, 3 * K - 1)
    W3p = torch.zeros(Hd, half, NT, dtype=torch.float32, device=device)
    W3p[:, :, :3 * K - 1] = W3
    b3p = torch.zeros(half, NT, dtype=torch.float32, device=device)
    b3p[:, :3 * K - 1] = b3.reshape(half, 3 * K - 1)
    parts = []
    for Wt in (W1.t().contiguous(), ):
        parts += list(_tf32_planes(Wt))
    parts.append(b1)
    parts += list(_tf32_planes(W2.t().contiguous()))
    parts.append(b2)
    parts += list(_tf32_planes(W3p.reshape(Hd, half * NT).t().contiguous()))
    parts.append(b3p.reshape(-1))
    out = torch.cat([p.reshape(-1) for p in parts]).contiguous()
    assert out.numel() == lib.wf_rqs_coupling_tc_net_floats()
    return out


def coupling_flow_tc(packed_weights: torch.Tensor, n_layers: int, x, B: float, inverse: bool = False, chunk_rows: int = 1 << 17):
    """Tensor-core Serial(NeuralSplineCoupling * L) for D = 64 / K = 64 / hidden 512 -> (y, log_det)."""
    x = f32(x)
    N, D = x.shape
    if D != 64:
        raise _ffi.WaveflowB200Error("D must be 64")
    rows = min(chunk_rows, (N + 127) // 128 * 128)
    ws = torch.empty(lib.wf_rqs_coupling_tc_workspace_floats(rows), dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    ld = torch.empty(N, dtype=torch.float32, device=x.device)
    st = lib.wf_rqs_coupling_flow_tc(ptr(packed_weights), n_layers, float(B), int(bool(inverse)), ptr(x), N, ptr(y), ptr(ld),
                                     ptr(ws), ws.numel(), stream_ptr())
    check(st, "wf_rqs_coupling_flow_tc")
    return y, ld


def FCNN(out_dim, hidden_dim):
    """stax.serial(Dense, Tanh, Dense, Tanh, Dense) (neural_splines.py:187-188): -> init_fun(rng, in_dim) -> params."""
    def init_fun(rng, in_dim):
        from ..splines.factories import _gen
        g = _gen(rng)

        def dense(i, o):       # stax.Dense: glorot-normal W, N(0, 1e-2) b
            return (torch.randn(i, o, generator=g) * (2.0 / (i + o)) ** 0.5, torch.randn(o, generator=g) * 1e-2)
        return [dense(in_dim, hidden_dim), (), dense(hidden_dim, hidden_dim), (), dense(hidden_dim, out_dim)]
    return init_fun


def NeuralSplineCoupling(K=5, B=3, hidden_dim=8, network=FCNN):
    """neural_splines.py:244-296.  Unlike the reference (which closes over its initial parameters and ignores the
    `params` argument, quirk Q7) the functions here use the parameters they are given."""
    def init_fun(rng, dim, **kwargs):
        from .bijections import _split_rng
        f1_rng, f2_rng = _split_rng(rng)
        f1_params = network((3 * K - 1) * dim // 2, hidden_dim)(f1_rng, dim // 2)
        f2_params = network((3 * K - 1) * dim // 2, hidden_dim)(f2_rng, dim // 2)

        def direct_fun(params, x, **kwargs):
            return coupling_flow([params], x, K, B, hidden_dim, inverse=False)

        def inverse_fun(params, z, **kwargs):
            return coupling_flow([params], z, K, B, hidden_dim, inverse=True)

        direct_fun.wf_layer = ("coupling", dict(K=K, B=float(B), hidden=hidden_dim))
        return (f1_params, f2_params), direct_fun, inverse_fun

    return init_fun
