"""Host-side driver of the fused live-path kernels (wf_live_forward / wf_local_energy).

Turns a model description + the reference's parameter pytree into the packed device buffers the kernels read:
MADE masks applied (model_factory.py:8-19,31-33), third layer re-ordered per dimension and zero-padded to 32
coefficients (model_factory.py:59-60).  Everything here is torch-on-device plumbing; the arithmetic of the path is in
csrc/live_device.cuh.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _ffi
from ._ffi import LiveModelStruct, check, lib, ptr, stream_ptr
from .splines.tables import SplineTables

HIDDEN = _ffi.WF_HIDDEN
MAXP = _ffi.WF_MAX_P


def made_masks(D: int, hidden: int = HIDDEN):
    """model_factory.py:8-19 (num_hidden=1): [D,H], [H,H], [H,D] float32 masks."""
    deg = [np.arange(D), np.arange(hidden) % (D - 1), np.arange(hidden) % (D - 1), np.arange(D) % D - 1]
    return [(d1[:, None] >= d0[None, :]).T.astype(np.float32) for d0, d1 in zip(deg[:-1], deg[1:])]


_MASK_CACHE: dict = {}


def _masks_on(D: int, device):
    key = (D, str(device))
    if key not in _MASK_CACHE:
        _MASK_CACHE[key] = [torch.from_numpy(m).to(device) for m in made_masks(D)]
    return _MASK_CACHE[key]


def net_arrays(net):
    """(stax.serial params, zero_params) -> W1, b1, W2, b2, W3, b3  (reference pytree, model_factory.py:86-88)."""
    nn = net[0]
    (W1, b1), _, (W2, b2), _, (W3, b3) = nn
    return W1, b1, W2, b2, W3, b3


def pack_net(net, D: int, P: int, device, fold: torch.Tensor | None = None) -> torch.Tensor:
    """One conditioner -> flat float32 [wf_live_net_floats(D)] in the kernel's layout.

    fold (B prior only): the [P, P] matrix  diag(boundary mask) @ ob_to_b.  Everything between the third layer and the
    un-normalised B-spline coefficients is linear (bsplines_jax.py:132-134, 173-198), so it is multiplied into W3 / b3 here,
    once per parameter set, instead of once per walker and dimension in the kernel; column 31 of every dimension carries
    sum_p o_p, whose sign the conditioner's own normalisation leaves behind (model_factory.py:69-70)."""
    W1, b1, W2, b2, W3, b3 = [torch.as_tensor(a, dtype=torch.float32, device=device) for a in net_arrays(net)]
    if W1.shape != (D, HIDDEN) or W2.shape != (HIDDEN, HIDDEN) or W3.shape != (HIDDEN, D * P):
        raise _ffi.WaveflowB200Error(f"unexpected conditioner shapes {tuple(W1.shape)}, {tuple(W2.shape)}, {tuple(W3.shape)} "
                                     f"for D={D}, P={P} (hidden width is fixed at 64, model_factory.py:72)")
    m1, m2, m3 = _masks_on(D, device)
    # hidden units sorted by MADE degree (stable): a relabelling of the units that lets the kernel skip whole masked
    # segments (csrc/live_device.cuh, deg_prefix); invisible in the results
    perm = torch.from_numpy(np.argsort(np.arange(HIDDEN) % (D - 1), kind="stable")).to(device)
    W1m, b1p = (W1 * m1)[:, perm], b1[perm]
    W2m, b2p = (W2 * m2)[perm][:, perm], b2[perm]
    W3m = (W3 * m3.repeat(1, P))[perm].reshape(HIDDEN, P, D).permute(0, 2, 1)    # [64, D, P]
    b3m = b3.reshape(P, D).t()                                                     # [D, P]
    W3p = torch.zeros(HIDDEN, D, MAXP, dtype=torch.float32, device=device)
    b3p = torch.zeros(D, MAXP, dtype=torch.float32, device=device)
    if fold is not None:
        F = fold.to(device=device, dtype=torch.float64)
        W3p[:, :, :P] = (W3m.double() @ F).float()
        b3p[:, :P] = (b3m.double() @ F).float()
        W3p[:, :, MAXP - 1] = W3m.double().sum(-1).float()
        b3p[:, MAXP - 1] = b3m.double().sum(-1).float()
    else:
        W3p[:, :, :P] = W3m
        b3p[:, :P] = b3m
    return torch.cat([W1m.reshape(-1), b1p.reshape(-1), W2m.reshape(-1), b2p.reshape(-1), W3p.reshape(-1), b3p.reshape(-1)])


def _bc_bits_I(left: dict, right: dict) -> int | None:
    """Constraint sets the fused kernel supports for I-splines: {} or {0:0} on the left, {} or {0:1} on the right."""
    bits = 0
    if left:
        if list(left.items()) != [(0, 0)] and list(left.items()) != [(0, 0.0)]:
            return None
        bits |= 1
    if right:
        if list(right.keys()) != [0] or float(list(right.values())[0]) != 1.0:
            return None
        bits |= 2
    return bits


def _bc_bits_P(left: dict, right: dict) -> int | None:
    bits = 0
    for i, d in enumerate((left, right)):
        if d:
            if list(d.keys()) != [0] or float(list(d.values())[0]) != 0.0:
                return None
            bits |= 1 << i
    return bits


@dataclass
class LiveSpec:
    """Everything static about a model assembled by model_factory (mirrors struct wf_live_model)."""
    D: int
    n_layers: int
    tab_I: SplineTables
    k_I: int
    reg: float = 0.0
    tol: float = 1e-6
    bc_I_left: dict = field(default_factory=dict)
    bc_I_right: dict = field(default_factory=dict)
    prior: str | None = None            # 'B', 'M' or None (uniform prior)
    tab_P: SplineTables | None = None
    k_P: int = 0
    bc_P_left: dict = field(default_factory=dict)
    bc_P_right: dict = field(default_factory=dict)
    box: float | None = None
    coord: str = "mean"

    def fusible(self) -> bool:
        return (2 <= self.D <= 4 and self.tab_I.P <= MAXP and (self.tab_P is None or self.tab_P.P <= MAXP)
                and _bc_bits_I(self.bc_I_left, self.bc_I_right) is not None
                and _bc_bits_P(self.bc_P_left, self.bc_P_right) is not None)

    def struct(self) -> LiveModelStruct:
        s = LiveModelStruct()
        s.D, s.n_layers, s.T = self.D, self.n_layers, self.tab_I.T
        s.P_I, s.k_I = self.tab_I.P, self.k_I
        s.prior_kind = {"B": _ffi.KIND_B, "M": _ffi.KIND_M, None: -1}[self.prior]
        s.P_P = self.tab_P.P if self.tab_P is not None else 0
        s.k_P = self.k_P
        s.has_box = int(self.box is not None)
        s.coord_mean = int(self.coord == "mean")
        s.bc_I = _bc_bits_I(self.bc_I_left, self.bc_I_right)
        s.bc_P = _bc_bits_P(self.bc_P_left, self.bc_P_right)
        s.box = float(self.box or 0.0)
        s.reg, s.tol = float(self.reg), float(self.tol)
        s.n_knots_P = len(self.tab_P.knots) if self.tab_P is not None else 0
        return s


def pack_params(spec: LiveSpec, transform_params, sp_params, device, fold_prior: bool = True) -> torch.Tensor:
    """Reference pytrees -> one flat device buffer: IMADE nets in order, then the prior net.

    fold_prior: pre-multiply the B prior's third layer by mask @ ob_to_b (see pack_net); the returned tensor is tagged
    `wf_folded` and wf_live_forward / wf_local_energy are told through bit 2 of bc_P.  The sampler / inverse kernels need the
    raw conditioner outputs: pack with fold_prior=False for them."""
    nets = [p for p in transform_params if len(p)]
    if len(nets) != spec.n_layers:
        raise _ffi.WaveflowB200Error(f"expected {spec.n_layers} IMADE parameter blocks, got {len(nets)}")
    parts = [pack_net(n, spec.D, spec.tab_I.P, device) for n in nets]
    folded = False
    if spec.prior is not None:
        fold = None
        if fold_prior and spec.prior == "B" and spec.tab_P.P <= MAXP - 1 and _bc_bits_P(spec.bc_P_left, spec.bc_P_right) is not None:
            P = spec.tab_P.P
            mask = np.ones(P)
            bits = _bc_bits_P(spec.bc_P_left, spec.bc_P_right)
            if bits & 1:
                mask[0] = 0.0
            if bits & 2:
                mask[P - 1] = 0.0
            fold = torch.from_numpy(mask[:, None] * np.asarray(spec.tab_P.ob_to_b64, dtype=np.float64))
            folded = True
        parts.append(pack_net(sp_params, spec.D, spec.tab_P.P, device, fold=fold))
    out = torch.cat(parts).contiguous()
    out.wf_folded = folded
    out.wf_tc = None
    if out.is_cuda and tc_capable(spec) and (spec.prior != "B" or folded):
        out.wf_tc = pack_tc(spec, out)
    return out


def tc_capable(spec: LiveSpec) -> bool:
    """Models the tensor-core kernels (csrc/live_tc.cuh) are instantiated for."""
    return 2 <= spec.D <= 4


def pack_tc(spec: LiveSpec, packed: torch.Tensor) -> torch.Tensor:
    """wf_live_pack_tc: the packed weights -> the tensor-core image (TF32 hi / lo planes of layers 2 and 3 in the swizzled
    K-major layout of the tcgen05 shared-memory descriptors); one launch."""
    n_nets = spec.n_layers + (1 if spec.prior is not None else 0)
    out = torch.empty(n_nets * int(lib.wf_live_net_floats_tc(spec.D)), dtype=torch.float32, device=packed.device)
    check(lib.wf_live_pack_tc(spec.D, n_nets, ptr(packed), ptr(out), stream_ptr()), "wf_live_pack_tc")
    out.wf_folded = getattr(packed, "wf_folded", False)
    return out


# 'auto' = the tensor-core kernels whenever the model has an image: measured on the B200 (tools/tc_time.py) they are faster
# than the CUDA-core kernels at every batch size tried, 256 ... 65 536 walkers (1.05x - 2.8x), so there is no size threshold
TC_MIN_ROWS = 0
TC_AUTO = True


def select_weights(spec: LiveSpec, weights: torch.Tensor, n: int, lap: bool, mode: str | None = None):
    """-> (weights buffer, layout) for a forward / local-energy call on `n` walkers.  mode (or the environment variable
    WAVEFLOW_B200_LIVE) = 'simt' | 'tc' forces a path ('tc' raises when the model has no tensor-core image); default 'auto':
    tensor cores once the batch fills the machine."""
    import os
    mode = mode or os.environ.get("WAVEFLOW_B200_LIVE", "auto")
    tc = getattr(weights, "wf_tc", None)
    if mode == "tc" and tc is None:
        raise _ffi.WaveflowB200Error("WAVEFLOW_B200_LIVE=tc but this model / parameter set has no tensor-core image")
    rows = n * (spec.D + 2 if lap else 1)
    if tc is not None and mode != "simt" and (mode == "tc" or (TC_AUTO and rows >= TC_MIN_ROWS)):
        return tc, _ffi.WEIGHTS_TC
    return weights, _ffi.WEIGHTS_SIMT


def packed_for(spec: LiveSpec, transform_params, sp_params, device, fold_prior: bool = True) -> torch.Tensor:
    """pack_params through a per-model cache keyed on the identity / version of the parameter leaves (_ffi.params_key): the
    reference's call signature h_fn(params, x) / psi(params, x) / log_pdf(params, x) carries the raw pytree on every call,
    the packed kernel layout is rebuilt only when the parameters changed."""
    cache = spec.__dict__.setdefault("_pack_cache", _ffi.PackCache())
    tree = (transform_params, sp_params)
    key = (str(device), bool(fold_prior), _ffi.params_key(tree))
    return cache.get(key, tree, lambda: pack_params(spec, transform_params, sp_params, device, fold_prior=fold_prior))


def _struct_for(spec: LiveSpec, weights, layout: int = _ffi.WEIGHTS_SIMT) -> LiveModelStruct:
    st = spec.struct()
    if getattr(weights, "wf_folded", False):
        st.bc_P |= 4
    st.weight_layout = layout
    return st


def _tables(spec: LiveSpec, device) -> _ffi.LiveTablesStruct:
    """struct wf_live_tables for this model on `device` (the tensors are cached in SplineTables.dev)."""
    t = _ffi.LiveTablesStruct()
    dI = spec.tab_I.dev(device)
    p = lambda x: None if x is None else x.data_ptr()
    t.dense_I, t.rec_I, t.lo_I, t.rec_I_t = p(dI["dense32"]), p(dI["rec"]), p(dI["lo"]), p(dI["rec_t"])
    if spec.prior == "B":
        dP = spec.tab_P.dev(device)
        t.dense_P, t.ob_to_b, t.b_to_ob = p(dP["ob_dense32"]), p(dP["ob_to_b"]), p(dP["b_to_ob"])
    elif spec.prior == "M":
        dP = spec.tab_P.dev(device)
        t.dense_P, t.rec_P, t.lo_P, t.rec_P_t = p(dP["dense32"]), p(dP["rec"]), p(dP["lo"]), p(dP["rec_t"])
    return t


def forward(spec: LiveSpec, weights: torch.Tensor, x: torch.Tensor, want=("u", "logdet"), mode: str | None = None):
    """wf_live_forward.  want: subset of {'u','logdet','logpdf','psi'} -> dict of tensors."""
    x = _ffi.f32(x)
    N = x.shape[0]
    dev = x.device
    tabs = _tables(spec, dev)
    out = {}
    if "u" in want:
        out["u"] = torch.empty(N, spec.D, dtype=torch.float32, device=dev)
    for k in ("logdet", "logpdf", "psi"):
        if k in want:
            out[k] = torch.empty(N, dtype=torch.float32, device=dev)
    weights, layout = select_weights(spec, weights, N, lap=False, mode=mode)
    st = lib.wf_live_forward(C.byref(_struct_for(spec, weights, layout)), C.byref(tabs), ptr(weights), ptr(x), N, ptr(out.get("u")),
                             ptr(out.get("logdet")), ptr(out.get("logpdf")), ptr(out.get("psi")), stream_ptr())
    check(st, "wf_live_forward")
    return out


def local_energy(spec: LiveSpec, weights: torch.Tensor, x: torch.Tensor, protons, want=("psi", "hpsi", "eloc"),
                 sums: torch.Tensor | None = None, exchange=None, mode: str | None = None):
    """wf_local_energy.  want: subset of {'psi','hpsi','eloc','grad','lap'}; sums: float64 [4] accumulator (zeroed by caller).
    exchange: a vqmc.PeerExchange -- the block sums are all-reduced over the ranks into exchange.out by the same call."""
    x = _ffi.f32(x)
    N = x.shape[0]
    dev = x.device
    tabs = _tables(spec, dev)
    prot = _ffi.host_f32(np.asarray(protons, dtype=np.float32).reshape(-1))
    out = {}
    for k in ("psi", "hpsi", "eloc", "lap"):
        if k in want:
            out[k] = torch.empty(N, dtype=torch.float32, device=dev)
    if "grad" in want:
        out["grad"] = torch.empty(N, spec.D, dtype=torch.float32, device=dev)
    if sums is not None and (sums.dtype != torch.float64 or sums.numel() != 4):
        raise _ffi.WaveflowB200Error("sums must be a float64 tensor with 4 elements")
    weights, layout = select_weights(spec, weights, N, lap=True, mode=mode)
    if exchange is None:
        st = lib.wf_local_energy(C.byref(_struct_for(spec, weights, layout)), C.byref(tabs), ptr(weights), _ffi.np_ptr(prot),
                                 int(prot.size), ptr(x), N, ptr(out.get("psi")), ptr(out.get("hpsi")), ptr(out.get("eloc")),
                                 ptr(out.get("grad")), ptr(out.get("lap")), ptr(sums), stream_ptr())
        check(st, "wf_local_energy")
    else:
        # local energy + estimator exchange over peer memory (vqmc.PeerExchange): fused into the kernel tail on the
        # tensor-core path, a second 32-thread launch otherwise
        if sums is None:
            raise _ffi.WaveflowB200Error("the estimator exchange needs the `sums` accumulator")
        step = exchange.next_step()
        st = lib.wf_local_energy_exchange(C.byref(_struct_for(spec, weights, layout)), C.byref(tabs), ptr(weights), _ffi.np_ptr(prot),
                                          int(prot.size), ptr(x), N, ptr(out.get("psi")), ptr(out.get("hpsi")), ptr(out.get("eloc")),
                                          ptr(out.get("grad")), ptr(out.get("lap")), ptr(sums), ptr(exchange.ptrs), exchange.rank,
                                          exchange.world, C.c_uint64(step), ptr(exchange.out), ptr(exchange.counter), stream_ptr())
        check(st, "wf_local_energy_exchange")
    return out


def inverse(spec: LiveSpec, weights: torch.Tensor, u: torch.Tensor, exact: bool = False) -> torch.Tensor:
    """wf_live_inverse: prior-space points [N, D] -> data space (Serial.inverse_fun)."""
    u = _ffi.f32(u)
    N = u.shape[0]
    x = torch.empty_like(u)
    tabs = _tables(spec, u.device)
    st = lib.wf_live_inverse(C.byref(_struct_for(spec, weights)), C.byref(tabs), ptr(weights), ptr(u), N, int(bool(exact)), ptr(x), stream_ptr())
    check(st, "wf_live_inverse")
    return x


def seed_of(rng) -> int:
    """The reference passes jax PRNG keys; here an int seed or a torch.Generator (one draw from it) seeds the Philox streams."""
    if isinstance(rng, torch.Generator):
        return int(torch.randint(0, 2 ** 62, (1,), generator=rng).item())
    return int(rng if rng is not None else 0)


def sample(spec: LiveSpec, weights: torch.Tensor, seed: int, n: int, device, exact: bool = False):
    """wf_live_sample -> (x [n, D] data-space samples, u [n, D] prior-space draws)."""
    x = torch.empty(n, spec.D, dtype=torch.float32, device=device)
    u = torch.empty(n, spec.D, dtype=torch.float32, device=device)
    tabs = _tables(spec, device)
    st = lib.wf_live_sample(C.byref(_struct_for(spec, weights)), C.byref(tabs), ptr(weights), C.c_uint64(int(seed) & (2 ** 64 - 1)), n,
                            int(bool(exact)), ptr(x), ptr(u), stream_ptr())
    check(st, "wf_live_sample")
    return x, u
