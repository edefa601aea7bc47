"""Hamiltonian of the 1-D soft-Coulomb systems -- reference: utils/physics.py:6-93.

`construct_hamiltonian_function(psi, protons)` returns h_fn(params, x) -> [N, 1] = -1/2 lap psi + V psi.  The reference
obtains the Laplacian from jax.hessian; here a fused forward-mode Laplacian kernel (wf_local_energy) evaluates psi, its
gradient and Laplacian in one pass, so `fn` must be the psi of a wavefunctions.Waveflow model."""
from __future__ import annotations

import numpy as np
import torch

from .. import _live
from .._ffi import WaveflowB200Error, f32

system_catalogue = {            # utils/physics.py:6-26 (+ the 4-electron box of BASELINE config 4)
    1: {
        'Laplacian_interactive_particles': (np.zeros((0, 1)), 2),
        'H': (np.array([[0.0]]), 1),
        'He+': (np.array([[0.0], [0.0]]), 1),
        'H2+': (np.array([[-0.9], [0.9]]), 1),
        'H2+_wide': (np.array([[-3.0], [3.0]]), 1),
        'He': (np.array([[0.0], [0.0]]), 2),
        'He_off_center': (np.array([[2.5], [2.5]]), 2),
        'H2': (np.array([[-0.9], [0.9]]), 2),
        'H2_wide': (np.array([[-3.0], [3.0]]), 2),
        'Be_1d': (np.array([[0.0]] * 4), 4),
    },
}


def get_potential(protons, max_val=None):
    """Soft-Coulomb potential (physics.py:60-76), torch ops; x [N, D] -> [N]."""
    prot = np.asarray(protons, dtype=np.float32).reshape(-1)

    def potential(x):
        x = f32(x)
        p = torch.as_tensor(prot, device=x.device)
        pe = -(1 / torch.sqrt(1 + (p[None, :, None] - x[:, None, :]) ** 2)).sum((-1, -2))
        diff = x[:, :, None] - x[:, None, :]
        il = torch.tril_indices(x.shape[1], x.shape[1], offset=-1, device=x.device)
        ee = (1 / torch.sqrt(1 + diff ** 2)[:, il[0], il[1]]).sum(-1)
        return pe + ee

    return potential


def construct_hamiltonian_function(fn, protons=np.array([[0, 0]]), n_space_dimensions=2, eps=0.0, max_potential_val=None):
    spec = getattr(fn, "wf_spec", None)
    if spec is None:
        raise WaveflowB200Error("construct_hamiltonian_function needs the psi of a Waveflow model in its fused "
                                "configuration (there is no generic autodiff Laplacian in this build)")
    if eps != 0.0:
        raise NotImplementedError("finite-difference Laplacian (physics.py:28-46) is not part of the hot path")
    prot = np.asarray(protons, dtype=np.float32).reshape(-1)

    def _construct(weight_dict, x, return_all=False, sums=None, packed=None, exchange=None, want=("psi", "hpsi", "eloc")):
        """h_fn(params, x) -> H psi [N, 1] (physics.py:84-93).  Extras of this build: return_all -> dict(psi, hpsi, eloc);
        sums: float64 [4] accumulator of {sum E, sum E^2, n, sum psi^2}; exchange: a vqmc.PeerExchange (the sums are
        all-reduced over the ranks by the same call); packed: pre-packed weights (default: the per-model cache)."""
        xx = f32(x)
        w = packed if packed is not None else _live.packed_for(spec, weight_dict[0], weight_dict[1], xx.device)
        out = _live.local_energy(spec, w, xx, prot, want=want, sums=sums, exchange=exchange)
        if return_all:
            return out
        return out["hpsi"][:, None]

    _construct.wf_spec = spec
    _construct.protons = prot
    return _construct
