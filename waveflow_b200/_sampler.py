"""Samplers: Waveflow.sample / MFlow.sample (wf_live_sample) and the spline factories' per-row rejection samplers."""
from __future__ import annotations

import torch

from . import _live
from ._ffi import WaveflowB200Error


def _seed_of(rng) -> int:
    if isinstance(rng, torch.Generator):
        return int(torch.randint(0, 2 ** 62, (1,), generator=rng).item())
    return int(rng if rng is not None else 0)


def sample(spec, weights, rng, n: int, device, exact: bool = False):
    """-> (x, u): data-space samples and the prior-space draws they come from."""
    return _live.sample(spec, weights, _seed_of(rng), n, device, exact=exact)


def rejection_sample_spline(tabs, kind, rng_array, params, num_samples, n_knots=None):
    raise WaveflowB200Error("stand-alone sample_fun_vec is served through Waveflow.sample / MFlow.sample "
                            "(wf_live_sample fuses the conditioner, the rejection sampler and the inverse flow)")
