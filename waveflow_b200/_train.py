"""Host-side driver of the VQMC training step (wf_vqmc_loss_grad / wf_adam_step).

The parameter pytree of the reference ((transform_params, sp_params), model_factory.py:86-88, wavefunctions.py:110) is
kept as ONE flat float32 device buffer whose order is the pytree's own leaf order; `unravel` hands out views into it, so
the tree the user sees, the buffer the kernels read and the buffer Adam updates are the same memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi, _live
from ._ffi import check, lib, ptr, stream_ptr


def tree_leaves(tree):
    """Leaves in jax.tree_util order (tuples / lists flattened in place, () and [] contribute nothing)."""
    if isinstance(tree, (tuple, list)):
        out = []
        for t in tree:
            out.extend(tree_leaves(t))
        return out
    return [tree]


def tree_unflatten(like, leaves):
    it = iter(leaves)

    def build(t):
        if isinstance(t, (tuple, list)):
            return type(t)(build(c) for c in t)
        return next(it)

    return build(like)


def ravel(params, device) -> torch.Tensor:
    """Pytree -> flat float32 device buffer."""
    parts = [torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(device=device, dtype=torch.float32).reshape(-1)
             for a in tree_leaves(params)]
    return torch.cat(parts).contiguous()


def unravel(like, flat: torch.Tensor):
    """Flat buffer -> pytree of VIEWS with the shapes of `like`."""
    leaves, o = [], 0
    for a in tree_leaves(like):
        shape = tuple(a.shape)
        n = int(np.prod(shape)) if shape else 1
        leaves.append(flat[o:o + n].view(shape))
        o += n
    if o != flat.numel():
        raise _ffi.WaveflowB200Error(f"parameter tree has {o} elements, flat buffer {flat.numel()}")
    return tree_unflatten(like, leaves)


_WS: dict = {}


def workspace_floats(spec: _live.LiveSpec, n: int, max_chunk: int) -> int:
    chunk = max(1, min(n, max_chunk))
    need = int(lib.wf_vqmc_grad_workspace_floats(C.byref(spec.struct()), chunk))
    if need < 0:
        raise _ffi.WaveflowB200Error("wf_vqmc_loss_grad does not support this model (Waveflow, D in 2..4, D*P <= 128)")
    return need


def _workspace(spec: _live.LiveSpec, n: int, device, max_chunk: int) -> torch.Tensor:
    """Process-wide scratch of the EAGER calls (grown on demand, stream-ordered reuse).  Captured CUDA graphs must not use
    it -- a later, larger call would replace the tensor under the graph's recorded pointer: GraphedTrainStep owns its own."""
    need = workspace_floats(spec, n, max_chunk)
    key = str(device)
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=device)
        _WS[key] = ws
    return ws


def loss_grad(spec: _live.LiveSpec, flat: torch.Tensor, x: torch.Tensor, protons, running_average: float,
              n_total: int | None = None, grad: torch.Tensor | None = None, want=(), sums: torch.Tensor | None = None,
              with_grad: bool = True, max_chunk: int = 65536, running_average_dev: torch.Tensor | None = None,
              ws: torch.Tensor | None = None):
    """wf_vqmc_loss_grad -> (grad flat [n_params] (accumulated into `grad` when given), dict of the `want`ed outputs).
    ws: caller-owned workspace (>= workspace_floats(spec, N, max_chunk) floats); default: the shared eager scratch."""
    x = _ffi.f32(x)
    N, dev = x.shape[0], x.device
    nparam = int(lib.wf_vqmc_param_floats(C.byref(spec.struct())))
    if nparam < 0:
        raise _ffi.WaveflowB200Error("wf_vqmc_loss_grad does not support this model (Waveflow, D in 2..4, D*P <= 128)")
    if flat.numel() != nparam or flat.dtype != torch.float32:
        raise _ffi.WaveflowB200Error(f"flat parameter buffer must hold {nparam} float32 values, got {flat.numel()}")
    if with_grad and grad is None:
        grad = torch.zeros(nparam, dtype=torch.float32, device=dev)
    out = {k: torch.empty(N, dtype=torch.float32, device=dev) for k in ("psi", "hpsi", "eloc") if k in want}
    if N == 0:                                   # empty shard: nothing to launch (a rank may own no walkers)
        return grad, out
    prot = _ffi.host_f32(np.asarray(protons, dtype=np.float32).reshape(-1))
    need = workspace_floats(spec, N, max_chunk)
    if ws is None:
        ws = _workspace(spec, N, dev, max_chunk)
    elif ws.numel() < need or ws.device != dev or ws.dtype != torch.float32:
        raise _ffi.WaveflowB200Error("workspace too small / wrong device or dtype for this batch")
    tabs = _live._tables(spec, dev)
    st = lib.wf_vqmc_loss_grad(C.byref(spec.struct()), C.byref(tabs), ptr(flat), _ffi.np_ptr(prot), int(prot.size), ptr(x), N,
                               float(running_average), ptr(running_average_dev), 1.0 / float(n_total or N),
                               ptr(grad if with_grad else None),
                               ptr(out.get("psi")), ptr(out.get("hpsi")), ptr(out.get("eloc")), ptr(sums), ptr(ws), need,   # `need`, not ws.numel():
                               # the C side sizes its walker chunks from the workspace it is told about, so max_chunk holds
                               # even when the shared scratch has grown beyond it
                               stream_ptr())
    check(st, "wf_vqmc_loss_grad")
    return grad, out


class AdamState:
    """optimizers.adam state: (x, m, v) as three flat buffers + the pytree template."""

    def __init__(self, like, flat: torch.Tensor):
        self.like, self.flat = like, flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.tree = unravel(like, flat)


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8, device="cuda"):
    """jax.example_libraries.optimizers.adam -> (opt_init, opt_update, get_params); the update is wf_adam_step, in place."""

    def opt_init(params):
        return AdamState(params, ravel(params, torch.device(device)))

    def opt_update(i, grads, state: AdamState, step_dev: torch.Tensor | None = None):
        g = grads if isinstance(grads, torch.Tensor) else ravel(grads, state.flat.device)
        lr = step_size(0 if step_dev is not None else i) if callable(step_size) else step_size
        st = lib.wf_adam_step(ptr(state.flat), ptr(state.m), ptr(state.v), ptr(g), state.flat.numel(),
                              0 if step_dev is not None else int(i), ptr(step_dev), float(lr), float(b1), float(b2), float(eps),
                              stream_ptr())
        check(st, "wf_adam_step")
        _ffi.bump_version(state.flat)        # the views handed out by get_params share this counter (pack caches key on it)
        return state

    opt_update.graphable = not callable(step_size)      # a schedule is evaluated on the host, per step

    def get_params(state: AdamState):
        return state.tree

    return opt_init, opt_update, get_params


class GraphedTrainStep:
    """One training step (zero the accumulators -> wf_vqmc_loss_grad -> wf_adam_step) captured in a CUDA graph.

    At the reference's batch sizes (128 / 256 walkers, vqmc.py:19-20) the step is ~90 short kernels and launch-bound; the
    graph replays them without host involvement.  Everything that changes between steps lives in device memory: the walkers
    (static buffer), the running average (float32[1]) and the Adam step index (int64[1])."""

    def __init__(self, spec, opt_state: AdamState, opt_update, protons, batch_shape, device, n_total=None, exchange=None):
        import weakref
        self.spec, self.state = spec, weakref.ref(opt_state)     # no strong reference: the cache entry must not keep it alive
        self.x = torch.zeros(batch_shape, dtype=torch.float32, device=device)
        self.ra = torch.zeros(1, dtype=torch.float32, device=device)
        self.step = torch.zeros(1, dtype=torch.int64, device=device)
        self.grad = torch.zeros_like(opt_state.flat)
        self.sums = torch.zeros(4, dtype=torch.float64, device=device)
        self.loss = torch.zeros((), dtype=torch.float32, device=device)
        n = batch_shape[0]
        # the graph records raw pointers: it owns every buffer it touches, including the activation workspace
        self.ws = torch.empty(workspace_floats(spec, n, n), dtype=torch.float32, device=device)

        def body():
            self.grad.zero_()
            self.sums.zero_()
            # sharded step (exchange given): this rank's shard with the GLOBAL 1 / n_total, then the flat gradient and the loss
            # sums are all-reduced by one peer-memory kernel inside the same graph
            loss_grad(spec, opt_state.flat, self.x, protons, 0.0, n_total=n_total or n, grad=self.grad, sums=self.sums,
                      running_average_dev=self.ra, max_chunk=n, ws=self.ws)
            total = exchange.all_reduce(self.grad, self.sums) if exchange is not None else self.sums
            opt_update(0, self.grad, opt_state, step_dev=self.step)
            self.loss.copy_((total[0] / float(n_total or n)).to(torch.float32))

        # warm-up on a side stream (first-call attribute setup, workspace allocation), with the optimiser state restored after
        keep = (opt_state.flat.clone(), opt_state.m.clone(), opt_state.v.clone())
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        for dst, src in zip((opt_state.flat, opt_state.m, opt_state.v), keep):
            dst.copy_(src)

    def __call__(self, epoch: int, batch: torch.Tensor, running_average: float) -> torch.Tensor:
        self.x.copy_(batch)
        self.ra.fill_(float(running_average))
        self.step.fill_(int(epoch))
        self.graph.replay()
        state = self.state()
        if state is not None:
            _ffi.bump_version(state.flat)
        return self.loss.clone()
