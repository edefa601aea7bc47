"""IMADE and BoxTransformLayer -- reference: flows/bijections/made.py:44-204 (same closure protocol).

Standalone, an IMADE layer runs at the reference's OPERATOR boundary: conditioner -> +reg -> wf_remove_bias ->
wf_enforce_bc -> wf_spline_apply (value, derivative, log-derivative in one pass).  Inside flows.Serial / MFlow /
Waveflow the whole stack is fused into one kernel (see bijections.Serial).
"""
from __future__ import annotations

import torch

from .. import _live
from .._ffi import WaveflowB200Error, f32
from ..splines.factories import ISpline_fun


def IMADE(transform, spline_degree=4, n_internal_knots=12, spline_regularization=0.0, reverse_fun_tol=0.0001,
          constraints_dict_left={0: 0.0}, constraints_dict_right={0: 1.0}, set_nn_output_grad_to_zero=False,
          n_spline_base_mesh_points=2000, cached_bases_path_root='./cached_splines_bases/I/'):
    def init_fun(rng, input_dim, **kwargs):
        (params_i, apply_fun_vec_i, apply_fun_vec_grad_i, reverse_fun_vec_i, knots_i, enforce_boundary_conditions,
         remove_bias) = ISpline_fun()(rng, spline_degree, n_internal_knots, use_cached_bases=True, cardinal_splines=True,
                                      zero_border=False, reverse_fun_tol=reverse_fun_tol,
                                      n_mesh_points=n_spline_base_mesh_points,
                                      cached_bases_path_root=cached_bases_path_root,
                                      constraints_dict_left=constraints_dict_left,
                                      constraints_dict_right=constraints_dict_right)
        P = params_i.shape[0]
        params, apply_fun = transform(rng, input_dim, P, set_nn_output_grad_to_zero=set_nn_output_grad_to_zero)

        def _coeffs(params, inputs):
            bp = apply_fun(params, inputs) + spline_regularization                # made.py:67-68
            bp = remove_bias(bp.reshape(-1, P))
            return enforce_boundary_conditions(bp)                                # [N*D, P]

        def direct_fun(params, inputs, **kwargs):
            x = f32(inputs)
            c = _coeffs(params, x)
            val, _grad, logd = apply_fun_vec_i.fused(c, x.reshape(-1))
            return val.reshape(-1, input_dim), logd.reshape(-1, input_dim).sum(-1)   # made.py:75-79

        def inverse_fun(params, inputs, **kwargs):
            y = f32(inputs)
            c = _coeffs(params, y).reshape(-1, input_dim, P)                      # conditioned on the INPUTS (quirk Q1)
            cols = [reverse_fun_vec_i(c[:, d, :].contiguous(), y[:, d].contiguous()) for d in range(input_dim)]
            return torch.stack(cols, dim=1), 0

        direct_fun.wf_layer = ("imade", dict(k=spline_degree, n_knots=n_internal_knots, reg=float(spline_regularization),
                                             tol=float(reverse_fun_tol), left=dict(constraints_dict_left),
                                             right=dict(constraints_dict_right), grad_to_zero=bool(set_nn_output_grad_to_zero),
                                             T=n_spline_base_mesh_points, tables=apply_fun_vec_i.tables))
        return params, direct_fun, inverse_fun

    return init_fun


def BoxTransformLayer(box_side=1, xu_coord_type='mean'):
    """Physical box [-L, L]^D (sorted coordinates) <-> unit cube (made.py:108-204)."""

    def init_fun(rng, input_dim, **kwargs):
        coord = "mean" if xu_coord_type == "mean" else "first"

        def direct_fun(params, inputs, **kwargs):
            from ..splines.tables import SplineTables
            x = f32(inputs)
            # box only: the fused kernel with zero flow layers (tables are not touched)
            spec = _live.LiveSpec(D=input_dim, n_layers=0, tab_I=SplineTables.get("I", 3, 6, 50), k_I=3, box=float(box_side),
                                  coord=coord)
            out = _live.forward(spec, None, x, want=("u", "logdet"))
            return out["u"], out["logdet"]

        def inverse_fun(params, inputs, **kwargs):
            u = f32(inputs)
            L = float(box_side)
            if coord == "mean":
                # made.py:186-197 (correct for D = 2 only, quirk Q2 -- reproduced as is)
                out = torch.zeros_like(u)
                out[:, 1:] = torch.cumsum(u[:, :-1], dim=-1)
                mean = out.mean(-1)
                w = out[:, -1]
                pm = u[:, -1] * (1 - w) - (0.5 - mean)
                return (out - mean[:, None] + pm[:, None]) * 2 * L, 0
            x = u.clone()
            x[:, 0] = (x[:, 0] - 0.5) * 2 * L
            for i in range(1, input_dim):
                x[:, i] = x[:, i] * (L - x[:, i - 1]) + x[:, i - 1]
            return x, 0

        direct_fun.wf_layer = ("box", dict(box_side=float(box_side), coord=coord))
        return (), direct_fun, inverse_fun

    return init_fun
