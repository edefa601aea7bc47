"""VQMC entry points -- reference: vqmc.py:19-221.

create_train_state (model construction + Adam state, vqmc.py:123-139), loss_fn_efficient (vqmc.py:193-200),
train_step_efficient = value_and_grad(loss_fn_efficient) + Adam update (vqmc.py:202-221) and a walker-sharded energy
estimator.  The gradient is wf_vqmc_loss_grad: a reverse pass through the forward-mode Laplacian (csrc/train.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _live, _train
from .model_factory import get_waveflow_model
from .utils import physics


def create_train_state(box_length, learning_rate, n_particle, rng=0, xu_coord_type='mean', spline_degree=6, num_knots=23,
                       n_flow_layers=3, cached_bases_root='./cached_splines_bases', device='cuda'):
    """-> (psi, log_pdf, sample, opt_state, opt_update, get_params)  (vqmc.py:123-139)."""
    init_fun = get_waveflow_model(n_particle, base_spline_degree=spline_degree, i_spline_degree=spline_degree,
                                  n_prior_internal_knots=num_knots, n_i_internal_knots=num_knots, i_spline_reg=0.05,
                                  i_spline_reverse_fun_tol=0.000001, n_flow_layers=n_flow_layers, box_size=box_length,
                                  xu_coord_type=xu_coord_type, cached_bases_root=cached_bases_root)
    params, psi, log_pdf, sample = init_fun(rng, n_particle)
    opt_init, opt_update, get_params = _train.adam(step_size=learning_rate, device=device)
    opt_state = opt_init(params)
    return psi, log_pdf, sample, opt_state, opt_update, get_params


def loss_fn_efficient(params, psi, h_fn, batch, running_average=None):
    """Forward value of vqmc.py:193-200: mean over the batch of E_loc = H psi / (psi + 1e-8)."""
    out = h_fn(params, batch, return_all=True)
    return out["eloc"].mean()


def _flat_of(params, opt_state, device):
    """The flat buffer behind `params`: opt_state's own buffer when params is get_params(opt_state), else a copy."""
    if opt_state is not None and params is opt_state.tree:
        return opt_state.flat
    return _train.ravel(params, device)


def value_and_grad_efficient(params, psi, h_fn, batch, running_average, group=None, opt_state=None, flat_grad=False,
                             n_total=None, exchange=None):
    """value_and_grad(loss_fn_efficient, argnums=0)(params, psi, h_fn, batch, running_average)  (vqmc.py:220).

    -> (loss, gradients in the structure of `params`).  With torch.distributed initialised, `batch` is this rank's shard of
    the walkers: the loss sums and the flat gradient are all-reduced, so every rank gets the global result.
    n_total: number of walkers over all ranks when the caller knows it (saves the count all-reduce and its host
    synchronisation every step; shards may be ragged, so it is not inferred).
    """
    spec = h_fn.wf_spec
    if spec is None:
        raise _live._ffi.WaveflowB200Error("train_step_efficient needs the fused Waveflow configuration of get_waveflow_model")
    x = _live._ffi.f32(batch)
    dev = x.device
    flat = _flat_of(params, opt_state, dev)
    n_total = int(n_total) if n_total is not None else total_walkers(x.shape[0], dev, group)
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    grad, _ = _train.loss_grad(spec, flat, x, h_fn.protons, float(running_average), n_total=n_total, sums=sums)
    loss = reduce_loss_and_grad(grad, sums, n_total, group, exchange=exchange)
    return loss, (grad if flat_grad else _train.unravel(params, grad))


def _world(group=None):
    dist = torch.distributed
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def total_walkers(n_local: int, device, group=None) -> int:
    """Number of walkers over all ranks (the 1/N of the loss mean and of the gradient; shards may be ragged)."""
    if _world(group) == 1:
        return int(n_local)
    cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
    torch.distributed.all_reduce(cnt, group=group)
    return int(cnt.item())


def reduce_loss_and_grad(grad: torch.Tensor, sums: torch.Tensor, n_total: int, group=None, exchange=None) -> torch.Tensor:
    """The exchange step of the sharded training step: every rank evaluated its shard with the global 1/N, so the
    flat gradients simply add; the loss is the mean over ALL walkers, as jnp.mean does: a non-finite E_loc is not dropped, it makes
    the loss (and the gradient) non-finite on every path -- kernels, eager loss_fn_efficient and the graph replay agree;
    sums[2] counts all walkers.  In place on grad / sums; backend-agnostic (NCCL, gloo in the CPU test)."""
    if _world(group) > 1:
        if exchange is not None:
            sums = exchange.all_reduce(grad, sums)        # one kernel over NVLink peer memory (GradExchange)
        else:
            torch.distributed.all_reduce(grad, group=group)
            torch.distributed.all_reduce(sums, group=group)
    return (sums[0] / float(n_total)).to(torch.float32)


GRAPH_MAX_BATCH = 16384     # at or below this many walkers per rank the step is launch-bound and is replayed from a CUDA graph
_GRAPHS: dict = {}


def train_step_efficient(epoch, psi, h_fn, opt_update, opt_state, params, batch, running_average, group=None, use_graph=None,
                         n_total=None, exchange=None):
    """vqmc.py:214-221: -> (opt_update(epoch, gradients, opt_state), loss_val).

    Steps on small batches (the reference trains with 128 / 256 walkers) are launch bound: they are captured once per (optimiser
    state, batch shape) in a CUDA graph and replayed (use_graph=False disables, True forces).  Under torch.distributed `batch`
    is this rank's shard; with `exchange` (a GradExchange: flat gradient + loss sums all-reduced by one kernel over NVLink peer
    memory) and a known `n_total` the whole sharded step -- gradient, exchange, Adam -- is ONE graph replay per rank; without
    it the exchange is an NCCL / gloo all-reduce between eagerly launched kernels."""
    dist = torch.distributed
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    x = _live._ffi.f32(batch)
    graphable = (getattr(opt_update, "graphable", False) and params is opt_state.tree and h_fn.wf_spec is not None
                 and (world == 1 or (exchange is not None and n_total is not None)))
    if use_graph is None:
        use_graph = graphable and x.shape[0] <= GRAPH_MAX_BATCH
    if use_graph and graphable:
        key = (id(opt_state), tuple(x.shape), str(x.device), n_total if world > 1 else None)
        g = _GRAPHS.get(key)
        if g is None or g.state() is not opt_state:
            import weakref
            g = _GRAPHS[key] = _train.GraphedTrainStep(h_fn.wf_spec, opt_state, opt_update, h_fn.protons, tuple(x.shape), x.device,
                                                       n_total=n_total if world > 1 else None,
                                                       exchange=exchange if world > 1 else None)
            weakref.finalize(opt_state, _GRAPHS.pop, key, None)      # the graph (and its buffers) die with the optimiser state
        return opt_state, g(epoch, x, float(running_average))
    loss_val, gradients = value_and_grad_efficient(params, psi, h_fn, batch, running_average, group=group, opt_state=opt_state,
                                                   flat_grad=True, n_total=n_total, exchange=exchange)
    return opt_update(epoch, gradients, opt_state), loss_val


class GradExchange:
    """All-reduce of the flat parameter gradient (+ the four loss sums) over NVLink peer memory: wf_p2p_allreduce_vec on
    symmetric buffers from torch.distributed._symmetric_memory -- one kernel of ~24 CTAs instead of two NCCL all-reduces, and,
    with the step counter in device memory, capturable in the CUDA graph of the training step.  Construction is collective."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._ffi import lib
        dist = torch.distributed
        group = group if group is not None else dist.group.WORLD
        self.n = int(n)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib.wf_p2p_allreduce_vec_buffer_bytes(self.world, self.n))
        if nbytes < 0:
            raise RuntimeError("unsupported world size for the peer-memory gradient exchange")
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.step = torch.ones(1, dtype=torch.int64, device=device)          # advanced by the kernel itself
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.sums_out = torch.zeros(4, dtype=torch.float64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def all_reduce(self, grad: torch.Tensor, sums: torch.Tensor) -> torch.Tensor:
        """grad (flat float32 [n]) is summed over the ranks in place; -> the summed loss sums (float64 [4])."""
        from ._ffi import check, lib, ptr, stream_ptr
        import ctypes as C
        if grad.numel() != self.n or grad.dtype != torch.float32:
            raise ValueError("gradient buffer does not match the exchange")
        check(lib.wf_p2p_allreduce_vec(ptr(self.ptrs), self.rank, self.world, C.c_uint64(0), ptr(self.step), ptr(grad), self.n, ptr(sums),
                                       ptr(self.sums_out), ptr(self.counter), stream_ptr()), "wf_p2p_allreduce_vec")
        return self.sums_out

    def failed_step(self) -> int:
        return int(self.buf.view(torch.int64)[-8].item())


class PeerExchange:
    """Estimator exchange over NVLink peer memory (wf_p2p_allreduce_sums): symmetric buffers from
    torch.distributed._symmetric_memory, one 32-thread kernel per step instead of an NCCL all-reduce (pure latency at 32 bytes).
    Construction is collective; raises if symmetric memory is unavailable (callers fall back to NCCL)."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._ffi import lib
        dist = torch.distributed
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib.wf_p2p_allreduce_buffer_bytes(self.world))
        if nbytes < 0:
            raise RuntimeError("unsupported world size for the peer-memory exchange")
        self.buf = symm_mem.empty(nbytes // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.out = torch.zeros(4, dtype=torch.float64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)      # CTA retirement counter of the fused tail
        self.step = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # every buffer is zeroed and mapped before the first flag is written

    def next_step(self) -> int:
        self.step += 1
        return self.step

    def all_reduce(self, sums: torch.Tensor) -> torch.Tensor:
        from ._ffi import check, lib, ptr, stream_ptr
        import ctypes as C
        self.step += 1
        check(lib.wf_p2p_allreduce_sums(ptr(self.ptrs), self.rank, self.world, C.c_uint64(self.step), ptr(sums), ptr(self.out),
                                        stream_ptr()), "wf_p2p_allreduce_sums")
        return self.out

    def failed_step(self) -> int:
        """Sticky error word of this rank's buffer: 0, or the step at which a peer did not show up within ~2 s (from then
        on every exchange returns NaN on ALL ranks, see include/waveflow_b200.h).  Synchronises the device."""
        return int(self.buf.view(torch.int64)[2 * self.world * 8].item())

    def check(self):
        step = self.failed_step()
        if step:
            raise RuntimeError(f"peer-memory estimator exchange timed out at step {step} on rank {self.rank}: a peer rank is "
                               "missing or more than ~2 s behind")


class EnergyEstimator:
    """Walker-sharded energy estimator: each rank evaluates its row block with wf_local_energy (block sums accumulated
    in-kernel in float64) and one 32-byte exchange merges {sum E, sum E^2, n, sum psi^2}: the peer-memory
    one-shot kernel when `peer_exchange` is set (see PeerExchange), an NCCL / gloo all-reduce otherwise."""

    def __init__(self, h_fn, params, device, group=None, peer_exchange=None):
        self.h_fn, self.spec = h_fn, h_fn.wf_spec
        self.device = device
        self.group = group
        self.peer = peer_exchange
        self.packed = _live.pack_params(self.spec, params[0], params[1], device)

    @staticmethod
    def shard(n_total: int, rank: int, world: int):
        """Contiguous row block of rank `rank`: rows [lo, hi)."""
        base, rem = divmod(n_total, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def local_sums(self, walkers: torch.Tensor, sums: torch.Tensor | None = None) -> torch.Tensor:
        if sums is None:
            sums = torch.zeros(4, dtype=torch.float64, device=self.device)
        _live.local_energy(self.spec, self.packed, walkers, self.h_fn.protons, want=(), sums=sums)
        return sums

    def step_sums(self, walkers: torch.Tensor, sums: torch.Tensor) -> torch.Tensor:
        """One estimator step: local block sums of this rank's walkers + the exchange, -> the global sums on every rank.
        With a peer exchange this is ONE call of wf_local_energy_exchange (on the tensor-core path one kernel launch: the
        last CTA to retire does the peer stores and the flag wait); otherwise local_sums followed by exchange."""
        if _world(self.group) > 1 and self.peer is not None:
            _live.local_energy(self.spec, self.packed, walkers, self.h_fn.protons, want=(), sums=sums, exchange=self.peer)
            return self.peer.out
        return self.exchange(self.local_sums(walkers, sums))

    def exchange(self, sums: torch.Tensor) -> torch.Tensor:
        """Device-side exchange of the block sums (no host synchronisation): -> the global sums on every rank."""
        if _world(self.group) == 1:
            return sums
        if self.peer is not None:
            return self.peer.all_reduce(sums)
        torch.distributed.all_reduce(sums, group=self.group)
        return sums

    @staticmethod
    def reduce(sums: torch.Tensor, group=None):
        """All-reduce the per-rank {sum E, sum E^2, n, sum psi^2} (float64 [4]) and turn them into the estimator.
        Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(sums, group=group)
        return EnergyEstimator.finish(sums)

    def finish_checked(self, sums: torch.Tensor):
        """finish() that turns a failed peer exchange (NaN sums + sticky error word) into an exception on every rank."""
        out = self.finish(sums)
        if self.peer is not None and not np.isfinite(float(out["n"])):      # the walker count is NaN only after a time-out
            self.peer.check()
        return out

    @staticmethod
    def finish(sums: torch.Tensor):
        s = sums.detach().cpu().numpy()
        mean = s[0] / s[2]
        n = int(s[2]) if np.isfinite(s[2]) else float("nan")
        return dict(energy=float(mean), variance=float(s[1] / s[2] - mean * mean), n=n, psi2=float(s[3]))

    def estimate(self, walkers: torch.Tensor):
        """-> dict(energy, variance, n) over all ranks."""
        return self.finish_checked(self.step_sums(walkers, torch.zeros(4, dtype=torch.float64, device=self.device)))


class ModelTrainer:
    """Mirror of vqmc.ModelTrainer's configuration surface (vqmc.py:19-51); `estimate_energy` runs sample -> local energy."""

    def __init__(self, system_name='He', learning_rate=1e-4, box_length=10, num_epochs=200000, batch_size=128, log_every=2000):
        self.system_name = system_name
        self.n_space_dimension = 1
        self.system, self.n_particle = physics.system_catalogue[self.n_space_dimension][self.system_name]
        self.box_length = box_length
        self.xu_coord_type = 'mean'
        self.spline_degree = 6
        self.num_knots = 23
        self.n_flow_layer = 3
        self.learning_rate = learning_rate
        self.num_epochs = num_epochs
        self.batch_size = batch_size
        self.log_every = log_every
        self.save_dir = f'./results/{self.system_name}_{self.n_space_dimension}d_L{self.box_length}box'

    def build(self, rng=2, cached_bases_root='./cached_splines_bases'):
        psi, log_pdf, sample, opt_state, opt_update, get_params = create_train_state(
            self.box_length, self.learning_rate, n_particle=self.n_particle, rng=rng, xu_coord_type=self.xu_coord_type,
            spline_degree=self.spline_degree, num_knots=self.num_knots, n_flow_layers=self.n_flow_layer,
            cached_bases_root=cached_bases_root)
        h_fn = physics.construct_hamiltonian_function(psi, protons=self.system, n_space_dimensions=self.n_space_dimension, eps=0.0)
        self.opt_state, self.opt_update, self.get_params = opt_state, opt_update, get_params
        return psi, log_pdf, sample, get_params(opt_state), h_fn

    def start_training(self, restart=False, num_epochs=None, rng=2, cached_bases_root='./cached_splines_bases', save=True,
                       callback=None):
        """The optimisation loop of vqmc.py:53-117: checkpoint at epoch 1 and every log_every epochs
        (helpers.create_checkpoint_wavefunc), sample a batch from |psi|^2, one train_step_efficient, running average of the last
        100 losses refreshed every 100 epochs.  restart=True resumes from save_dir/checkpoints.  -> (params, loss history)."""
        from pathlib import Path
        from .utils import helpers
        psi, log_pdf, sample, params, h_fn = self.build(rng=rng, cached_bases_root=cached_bases_root)
        opt_state, opt_update, get_params = self.opt_state, self.opt_update, self.get_params
        system_dict = {"system_name": self.system_name, "box_length": self.box_length, "n_particle": self.n_particle,
                       "n_space_dimension": self.n_space_dimension, "window": 100, "n_plotting": 200}
        start_epoch, loss, energies = 0, [0.0], []
        if restart and Path(f"{self.save_dir}/checkpoints").exists():
            saved, start_epoch, loss, energies = helpers.load_checkpoint(self.save_dir)
            opt_state.flat.copy_(_train.ravel(saved, opt_state.flat.device))        # Adam moments restart from zero
        if save:
            import json
            helpers.make_result_dirs(self.save_dir)
            with open(f"{self.save_dir}/system_info.json", "w") as f:
                json.dump(system_dict, f, indent=4)
        running_average = float(np.mean(loss[-100:])) if start_epoch >= 100 else 0.0
        for epoch in range(start_epoch + 1, start_epoch + (num_epochs or self.num_epochs) + 1):
            if save and (epoch % self.log_every == 0 or epoch == 1):
                helpers.create_checkpoint_wavefunc(rng * 7919 + epoch, self.save_dir, psi, sample, params, epoch, loss, energies,
                                                   system_dict)
            batch = sample(rng * 1000003 + epoch, params, self.batch_size)
            opt_state, new_loss = train_step_efficient(epoch, psi, h_fn, opt_update, opt_state, params, batch, running_average)
            if epoch % 100 == 0:
                running_average = float(np.mean(loss[-100:]))
            params = get_params(opt_state)
            loss.append(float(new_loss))
            energies.append([float(new_loss)])
            if callback is not None:
                callback(epoch, loss[-1], params)
            if epoch % self.log_every == 0:
                print(f"epoch {epoch} | Loss: {loss[-1]:.3f}")
        return params, loss
