"""Model zoo of the 2-D density benchmarks -- reference: benchmark_tests.py:50-77 (`get_model`).

Only the spline flows of the hot path are built here ('IFlow': uniform prior, 'MFlow': conditional M-spline prior); the
MADE-affine baseline ('Flow') is outside the path.  BASELINE configs[0] is
`get_model('MFlow', 0.02, spline_degree=5, num_knots=23, num_layers=3)` evaluated at batch 256 (bench.py, `configs.c1`).
"""
from __future__ import annotations

from . import flows
from ._ffi import WaveflowB200Error
from .model_factory import get_masked_transform


def get_model(model_type, spline_reg, spline_degree=3, num_knots=15, num_layers=5, reverse_tol=1e-6, prior_spline_degree=3,
              prior_num_knots=15, cached_bases_root=None):
    root = (lambda k: None) if cached_bases_root is None else (lambda k: f"{cached_bases_root}/{k}/")
    layers = (flows.IMADE(get_masked_transform(), spline_degree=spline_degree, n_internal_knots=num_knots,
                          spline_regularization=spline_reg, reverse_fun_tol=reverse_tol, cached_bases_path_root=root("I")),
              flows.Reverse()) * num_layers
    if model_type == 'IFlow':
        return flows.Flow(flows.Serial(*layers), flows.Uniform(), prior_support=(0.0, 1.0))
    if model_type == 'MFlow':
        return flows.MFlow(flows.Serial(*layers), get_masked_transform(), spline_degree=prior_spline_degree,
                           n_internal_knots=prior_num_knots, cached_bases_path_root=root("M"))
    raise WaveflowB200Error(f"model type {model_type!r} is not part of the spline hot path (supported: 'IFlow', 'MFlow')")
