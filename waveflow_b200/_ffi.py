"""ctypes binding of the C ABI in include/waveflow_b200.h (libwaveflow_b200.so, sm_100a).

There is NO CPU fallback: importing this module without the shared library raises, and every op raises on
non-CUDA tensors.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libwaveflow_b200.so"

WF_MAX_P = 32
WEIGHTS_SIMT, WEIGHTS_TC = 0, 1
WF_HIDDEN = 64
WF_WIN = 8
KIND_I, KIND_M, KIND_B = 0, 1, 2
KIND = {"I": KIND_I, "M": KIND_M, "B": KIND_B}


class WaveflowB200Error(RuntimeError):
    pass


class LiveModelStruct(C.Structure):
    """struct wf_live_model (include/waveflow_b200.h)."""
    _fields_ = [("D", C.c_int32), ("n_layers", C.c_int32), ("T", C.c_int32), ("P_I", C.c_int32), ("k_I", C.c_int32),
                ("prior_kind", C.c_int32), ("P_P", C.c_int32), ("k_P", C.c_int32), ("has_box", C.c_int32),
                ("coord_mean", C.c_int32), ("bc_I", C.c_int32), ("bc_P", C.c_int32), ("box", C.c_float),
                ("reg", C.c_float), ("tol", C.c_float), ("n_knots_P", C.c_int32), ("weight_layout", C.c_int32)]


class LiveTablesStruct(C.Structure):
    """struct wf_live_tables (include/waveflow_b200.h): device pointers."""
    _fields_ = [(n, C.c_void_p) for n in ("dense_I", "rec_I", "lo_I", "dense_P", "rec_P", "lo_P", "ob_to_b", "b_to_ob",
                                          "rec_I_t", "rec_P_t")]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        if os.environ.get("WAVEFLOW_B200_NO_AUTOBUILD"):
            raise WaveflowB200Error(f"{LIB_PATH} is missing: run `python -m waveflow_b200.build` (needs nvcc)")
        from . import build as _build
        _build.build()           # serialised across processes by a file lock; the library appears atomically (os.replace)
    return C.CDLL(str(LIB_PATH))


lib = _load()

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_SIGS = {
    "wf_abi_version": (C.c_int, [C.POINTER(C.c_int)]),
    "wf_status_string": (C.c_char_p, [_i]),
    "wf_probe_fma": (_i, [_i, _i, _p, _p]),
    "wf_table_layout_host": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "wf_spline_apply_dense": (_i, [_p, _i, _i, _p, _p, _l, _i, _i, C.POINTER(_p), _p, _p]),
    "wf_spline_apply_local": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _l, _p, _p, _p, _p]),
    "wf_bspline_apply": (_i, [_p, _p, _i, _i, _p, _p, _l, _i, _p, _p]),
    "wf_remove_bias": (_i, [_i, _i, _i, _p, _l, _p, _p]),
    "wf_enforce_bc": (_i, [_i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _l, _p, _p]),
    "wf_spline_reverse": (_i, [_p, _i, _i, _p, _p, _l, _f, _p, _p, _p]),
    "wf_spline_sample": (_i, [_p, _i, _i, _i, _p, _p, _i, _p, _l, _i, C.c_uint64, _p, _p, _p]),
    "wf_rqs_apply": (_i, [_p, _p, _p, _p, _l, _i, _f, _i, _p, _p, _p, _p]),
    "wf_live_net_floats": (_l, [_i]),
    "wf_live_net_floats_tc": (_l, [_i]),
    "wf_live_pack_tc": (_i, [_i, _i, _p, _p, _p]),
    "wf_local_energy_exchange": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, _p, _i, _p, _l, _p, _p, _p, _p, _p,
                                      _p, _p, _i, _i, C.c_uint64, _p, _p, _p]),
    "wf_tf32_split": (_i, [_p, _l, _p, _p, _p]),
    "wf_rqs_coupling_tc_net_floats": (_l, []),
    "wf_rqs_coupling_tc_workspace_floats": (_l, [_l]),
    "wf_rqs_coupling_flow_tc": (_i, [_p, _i, _f, _i, _p, _l, _p, _p, _p, _l, _p]),
    "wf_tc_dense": (_i, [_p, _p, _l, _i, _p, _p, _i, _p, _i, _p, _p, _p]),
    "wf_rqs_coupling_net_floats": (_l, [_i, _i, _i]),
    "wf_rqs_coupling_flow": (_i, [_p, _i, _i, _i, _i, _f, _i, _p, _l, _p, _p, _p]),
    "wf_live_forward": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, _p, _l, _p, _p, _p, _p, _p]),
    "wf_live_inverse": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, _p, _l, _i, _p, _p]),
    "wf_live_sample": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, C.c_uint64, _l, _i, _p, _p, _p]),
    "wf_vqmc_param_floats": (_l, [C.POINTER(LiveModelStruct)]),
    "wf_vqmc_grad_workspace_floats": (_l, [C.POINTER(LiveModelStruct), _l]),
    "wf_vqmc_loss_grad": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, _p, _i, _p, _l, _f, _p, _f, _p, _p, _p,
                               _p, _p, _p, _l, _p]),
    "wf_adam_step": (_i, [_p, _p, _p, _p, _l, _l, _p, _f, _f, _f, _f, _p]),
    "wf_p2p_allreduce_buffer_bytes": (_l, [_i]),
    "wf_p2p_allreduce_sums": (_i, [_p, _i, _i, C.c_uint64, _p, _p, _p]),
    "wf_p2p_allreduce_vec_buffer_bytes": (_l, [_i, _l]),
    "wf_p2p_allreduce_vec": (_i, [_p, _i, _i, C.c_uint64, _p, _p, _l, _p, _p, _p, _p]),
    "wf_p2p_allreduce_emulated": (_i, [_p, _i, C.c_uint64, _p, _p, _i, _l, _p]),
    "wf_local_energy": (_i, [C.POINTER(LiveModelStruct), C.POINTER(LiveTablesStruct), _p, _p, _i, _p, _l, _p, _p, _p, _p, _p,
                             _p, _p]),
}
class _DeviceGuardedLib:
    """The C ABI launches on the CURRENT device and stream.  Every call through this proxy first makes the device that holds
    the tensors of the call current (set by `ptr` / `f32` while the arguments are marshalled, see `_touch`), so tensors on
    cuda:1 are never launched on cuda:0 when one process drives several GPUs."""

    def __init__(self, cdll):
        object.__setattr__(self, "_cdll", cdll)

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if not name.startswith("wf_") or name in _HOST_ONLY:
            return fn

        def call(*args):
            dev = _pending_device()
            if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
                return fn(*args)                      # the common case (one process per GPU): nothing to switch
            with torch.cuda.device(dev):
                return fn(*args)

        call.__name__ = name
        object.__setattr__(self, name, call)
        return call


_HOST_ONLY = {"wf_abi_version", "wf_status_string", "wf_table_layout_host", "wf_live_net_floats", "wf_live_net_floats_tc",
              "wf_vqmc_param_floats",
              "wf_vqmc_grad_workspace_floats", "wf_p2p_allreduce_buffer_bytes", "wf_p2p_allreduce_vec_buffer_bytes", "wf_rqs_coupling_net_floats",
              "wf_rqs_coupling_tc_net_floats", "wf_rqs_coupling_tc_workspace_floats"}
_CALL_DEVICE = {"dev": None}


def _touch(t: torch.Tensor):
    """Remember the device of the tensors marshalled for the next call; mixing devices in one call is an error."""
    d = t.device
    cur = _CALL_DEVICE["dev"]
    if cur is not None and cur != d:
        _CALL_DEVICE["dev"] = None
        raise WaveflowB200Error(f"tensors of one call live on different devices ({cur} and {d})")
    _CALL_DEVICE["dev"] = d


def _pending_device():
    d = _CALL_DEVICE["dev"]
    _CALL_DEVICE["dev"] = None
    return d


for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
_cdll = lib
lib = _DeviceGuardedLib(_cdll)

ABI_VERSION = 2
if lib.wf_abi_version(None) != ABI_VERSION:
    raise WaveflowB200Error("libwaveflow_b200.so ABI version mismatch: rebuild with `python -m waveflow_b200.build --force`")


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib.wf_status_string(status).decode()
        raise WaveflowB200Error(f"{what or 'waveflow_b200'} failed: {msg} (status {status})")


def ptr(t) -> C.c_void_p:
    """Device pointer of a CUDA float32/int32/float64 tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise WaveflowB200Error("waveflow_b200 ops take CUDA tensors only (there is no CPU path)")
    if not t.is_contiguous():
        raise WaveflowB200Error("tensor must be contiguous")
    _touch(t)
    return C.c_void_p(t.data_ptr())


def f32(t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise WaveflowB200Error("expected a torch.Tensor on a CUDA device")
    if not t.is_cuda:
        raise WaveflowB200Error("waveflow_b200 ops take CUDA tensors only (there is no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def stream_ptr() -> C.c_void_p:
    """Current stream of the device that holds the tensors marshalled so far for this call (see `_touch`)."""
    d = _CALL_DEVICE["dev"]
    return C.c_void_p(torch.cuda.current_stream(d).cuda_stream)


def host_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def np_ptr(a: np.ndarray | None) -> C.c_void_p:
    return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)


def table_layouts(tab: np.ndarray, kind: str):
    """tab [4][P][T] (any float) -> (dense_t [T][4][PP] f32, rec [T][4][8] f32 | None, lo [T] i32 | None)  (host)."""
    tab32 = host_f32(tab)
    _, P, T = tab32.shape
    PP = (P + 3) & ~3
    dense = np.zeros((T, 4, PP), dtype=np.float32)
    rec = np.zeros((T, 4, WF_WIN), dtype=np.float32)
    lo = np.zeros(T, dtype=np.int32)
    st = lib.wf_table_layout_host(np_ptr(tab32), KIND[kind], P, T, np_ptr(dense), np_ptr(rec), np_ptr(lo))
    if st == -2:      # WF_ERR_UNSUPPORTED: no compact local-support form (e.g. orthonormalised B tables)
        return dense, None, None
    check(st, "wf_table_layout_host")
    return dense, rec, lo


# ---------------------------------------------------------------------------------------------- parameter identity
def bump_version(t: torch.Tensor):
    """Tell torch that `t` was modified in place by a kernel launched through the C ABI (a raw-pointer write does not touch
    the tensor's version counter, which `params_key` uses to detect changed parameters)."""
    try:
        torch._C._autograd._unsafe_set_version_counter((t,), (t._version + 1,))
    except Exception:            # noqa: BLE001 -- older / newer torch without the hook: a real (tiny) in-place op
        t.add_(0)


def params_key(tree) -> tuple:
    """Identity of a parameter pytree: (storage address, version counter) per torch leaf -- in-place updates bump
    the version (optimisers through torch, wf_adam_step through `bump_version`), new tensors have new addresses that are
    kept alive by the cache entry holding them.  numpy leaves are keyed by object identity and buffer address: like the
    reference's immutable JAX arrays they are expected not to be mutated in place."""
    key = []

    def walk(t):
        if isinstance(t, (tuple, list)):
            for c in t:
                walk(c)
        elif isinstance(t, torch.Tensor):
            key.append((t.data_ptr(), t._version))
        elif isinstance(t, np.ndarray):
            key.append((id(t), t.__array_interface__["data"][0], t.size))
        else:
            key.append(("py", id(t)))
    walk(tree)
    return tuple(key)


class PackCache:
    """Small LRU of packed device buffers keyed by params_key: packing (masks, regrouping, the prior fold -- dozens of small
    launches) runs once per parameter set instead of once per psi / log_pdf / h_fn call."""

    def __init__(self, size: int = 4):
        self.size, self.items = size, []

    def get(self, key, tree, make):
        for i, (k, _keep, val) in enumerate(self.items):
            if k == key:
                if i:
                    self.items.insert(0, self.items.pop(i))
                return val
        val = make()
        self.items.insert(0, (key, tree, val))      # `tree` keeps the leaves (and thus their addresses) alive
        del self.items[self.size:]
        return val
