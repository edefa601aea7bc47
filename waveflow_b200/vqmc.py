"""VQMC entry points -- reference: vqmc.py:19-221.

What is built here is the reference's *energy evaluation* path: create_train_state (model construction, vqmc.py:123-139),
loss_fn_efficient's forward value mean(H psi / (psi + 1e-8)) (vqmc.py:193-200) and a sharded energy estimator.  The
parameter gradient / Adam step of train_step_efficient (vqmc.py:202-221) is the next row of the scope table
and raises NotImplementedError.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _live
from .model_factory import get_waveflow_model
from .utils import physics


def create_train_state(box_length, learning_rate, n_particle, rng=0, xu_coord_type='mean', spline_degree=6, num_knots=23,
                       n_flow_layers=3, cached_bases_root='./cached_splines_bases'):
    """-> (psi, log_pdf, sample, params).  The reference also returns the Adam state (vqmc.py:136-139); see module doc."""
    init_fun = get_waveflow_model(n_particle, base_spline_degree=spline_degree, i_spline_degree=spline_degree,
                                  n_prior_internal_knots=num_knots, n_i_internal_knots=num_knots, i_spline_reg=0.05,
                                  i_spline_reverse_fun_tol=0.000001, n_flow_layers=n_flow_layers, box_size=box_length,
                                  xu_coord_type=xu_coord_type, cached_bases_root=cached_bases_root)
    params, psi, log_pdf, sample = init_fun(rng, n_particle)
    return psi, log_pdf, sample, params


def loss_fn_efficient(params, psi, h_fn, batch, running_average=None):
    """Forward value of vqmc.py:193-200: mean over the batch of E_loc = H psi / (psi + 1e-8)."""
    out = h_fn(params, batch, return_all=True)
    return out["eloc"].mean()


def train_step_efficient(*args, **kwargs):
    raise NotImplementedError("parameter gradients of the local energy (vqmc.py:202-221) are not built yet (SURVEY 8f rank 1)")


class EnergyEstimator:
    """Walker-sharded energy estimator: each rank evaluates its row block with wf_local_energy (block sums accumulated
    in-kernel in float64) and one 32-byte all-reduce merges {sum E, sum E^2, n, sum psi^2}."""

    def __init__(self, h_fn, params, device, group=None):
        self.h_fn, self.spec = h_fn, h_fn.wf_spec
        self.device = device
        self.group = group
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

    @staticmethod
    def reduce(sums: torch.Tensor, group=None):
        """All-reduce the per-rank {sum E, sum E^2, n, sum psi^2} (float64 [4]) and turn them into the estimator.
        Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(sums, group=group)
        s = sums.detach().cpu().numpy()
        mean = s[0] / s[2]
        return dict(energy=float(mean), variance=float(s[1] / s[2] - mean * mean), n=int(s[2]), psi2=float(s[3]))

    def estimate(self, walkers: torch.Tensor):
        """-> dict(energy, variance, n) over all ranks."""
        return self.reduce(self.local_sums(walkers), self.group)


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
        psi, log_pdf, sample, params = create_train_state(self.box_length, self.learning_rate, n_particle=self.n_particle,
                                                          rng=rng, xu_coord_type=self.xu_coord_type,
                                                          spline_degree=self.spline_degree, num_knots=self.num_knots,
                                                          n_flow_layers=self.n_flow_layer, cached_bases_root=cached_bases_root)
        h_fn = physics.construct_hamiltonian_function(psi, protons=self.system, n_space_dimensions=self.n_space_dimension, eps=0.0)
        return psi, log_pdf, sample, params, h_fn

    def start_training(self, restart=False):
        raise NotImplementedError("the optimisation loop needs parameter gradients (SURVEY 8f rank 1); "
                                  "use build() + EnergyEstimator for the energy evaluation path")
