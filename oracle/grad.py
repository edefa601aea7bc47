"""ORACLE (test infrastructure): parameter gradient of the VQMC loss.

Restates value_and_grad(loss_fn_efficient) (vqmc.py:193-221):

    loss  = mean_w  H psi_w / (psi_w + 1e-8)                                     (vqmc.py:193-200)
    dloss = mean_w [ 2 dpsi_w (E_w - avg) / psi_w + (dH_w psi_w - H_w dpsi_w) / psi_w^2 ]   (custom_jvp, vqmc.py:202-212)

i.e. the gradient of the surrogate  mean_w [ a_w psi_w + b_w (H psi)_w ]  with the per-walker constants
a_w = 2 (E_w - avg) / psi_w - H_w / psi_w^2 and b_w = 1 / psi_w held fixed, H psi = -1/2 lap psi + V psi
(physics.py:79-93).  The table lookup is differentiated the way jax differentiates its custom_jvp: every derivative
of lookup(nd) w.r.t. its argument is lookup(nd+1), and table index 4 clamps to 3 (jnp out-of-bounds indexing, SURVEY
quirk Q5) -- the parameter gradient is the only place that clamp is reachable.

``loss_and_grad`` takes three reverse passes through oracle/laplacian.py::autograd_psi_lap in float64.
PINNED BY THE REFERENCE'S OWN SOURCE: no gradients are stored anywhere upstream, but
tests/golden/make_energy_golden.py evaluates the unmodified vqmc.loss_fn_efficient -- including the estimator it registers with
custom_jvp -- on parameters seeded with dual numbers (numpy stand-in for jax, float64): the loss value and <grad loss, v> for two
random parameter directions v agree with ``loss_and_grad`` to 1e-12 -- also for a three-layer D = 4 model, where the clamp of
table index 4 to 3 (quirk Q5) is reached (tests/test_energy_reference_vectors.py).  Further pins:
the chain psi-KAT -> Laplacian oracles -> this function, and a finite-difference check of the surrogate on the parameters
whose gradient does not pass through a table argument (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

from . import laplacian, live


def coefficients(psi, hpsi, running_average):
    """(E_loc, a, b) of the surrogate, from numpy psi / H psi."""
    eloc = hpsi / (psi + 1e-8)
    a = 2.0 * (eloc - running_average) / psi - hpsi / psi ** 2
    b = 1.0 / psi
    return eloc, a, b


def loss_and_grad(m: live.LiveModel, params, x: np.ndarray, protons: np.ndarray, running_average: float, dtype=np.float64):
    """-> (loss, grads) with grads in the reference's parameter tree structure (zero_params get zeros).
    dtype=np.float32 runs the same three reverse passes in float32: the yardstick for what float32 arithmetic can deliver."""
    import torch

    psi, _, lap, leaves = laplacian.autograd_psi_lap(m, params, x, param_grad=True, dtype=dtype)
    V = torch.as_tensor(live.potential(np.asarray(x, dtype=np.float64), np.asarray(protons, dtype=np.float64)).astype(dtype))
    hpsi = -0.5 * lap + V * psi
    eloc, a, b = coefficients(psi.detach().numpy(), hpsi.detach().numpy(), running_average)
    surrogate = (torch.as_tensor(a.astype(dtype)) * psi + torch.as_tensor(b.astype(dtype)) * hpsi).mean()
    order = list(leaves.values())
    gs = torch.autograd.grad(surrogate, order, allow_unused=True)
    by_id = {k: (np.zeros(tuple(t.shape)) if g is None else g.numpy()) for (k, t), g in zip(leaves.items(), gs)}

    def net_grad(net):
        nn, zero = net
        return ([tuple(by_id[id(a)] for a in lay) for lay in nn], np.zeros_like(np.asarray(zero, dtype=np.float64)))

    tg = [net_grad(p) if len(p) else () for p in params[0]]
    return float(eloc.mean()), (tg, net_grad(params[1]))


def surrogate_value(m: live.LiveModel, params, x, protons, a, b):
    """mean_w [a_w psi_w + b_w H psi_w] through the *independent* numpy bundle oracle (finite-difference checks)."""
    r = laplacian.local_energy_bundle(m.cast(np.float64), params, x, protons)
    return float(np.mean(a * r["psi"] + b * r["hpsi"]))
