"""Model factory -- reference: model_factory.py:7-146.  Same function names, arguments and parameter pytrees
((stax.serial params, zero_params) per conditioner), so the reference's published checkpoints load unchanged."""
from __future__ import annotations

import torch

from . import flows, wavefunctions
from ._ffi import f32
from ._live import HIDDEN, _masks_on
from .splines.factories import _gen


def get_masked_transform(return_simple_masked_transform=False, allow_negative_params=False):
    """MADE-masked conditioner D -> 64 -> 64 -> D*P (model_factory.py:7-93)."""
    if return_simple_masked_transform:
        raise NotImplementedError("the MADE-affine baseline (simple_masked_transform) is outside the spline hot path")

    def masked_transform(rng, input_dim, output_shape=2, set_nn_output_grad_to_zero=False):
        g = _gen(rng)

        def dense(fan_in, fan_out):                        # MaskedDense.init_fun (model_factory.py:22-29)
            bound = 1.0 / (fan_in ** 0.5)
            W = (torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * bound
            b = (torch.rand(fan_out, generator=g) * 2 - 1) * bound
            return (W, b)

        zero_params = torch.rand(input_dim, output_shape, generator=g) - 0.5
        nn = [dense(input_dim, HIDDEN), (), dense(HIDDEN, HIDDEN), (), dense(HIDDEN, input_dim * output_shape)]
        params = (nn, zero_params)

        def calculate_bijection_params(params, x):
            """[N, D] -> [N, D, P] normalised coefficients (model_factory.py:56-70); standalone utility (torch ops) --
            inside Serial / MFlow / Waveflow the conditioner runs fused in the live-path kernel."""
            x = f32(x)
            params_nn, zero = params
            (W1, b1), _, (W2, b2), _, (W3, b3) = params_nn
            dev = x.device
            t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
            m1, m2, m3 = _masks_on(input_dim, dev)
            h = torch.tanh(x @ (t(W1) * m1) + t(b1))
            h = torch.tanh(h @ (t(W2) * m2) + t(b2))
            o = h @ (t(W3) * m3.repeat(1, output_shape)) + t(b3)
            p = o.reshape(-1, output_shape, input_dim).transpose(1, 2)             # p[n, d, q] = o[n, q*D + d]
            zero = t(zero)
            if not allow_negative_params:
                p = torch.sigmoid(p)
                zero = zero.abs()
            if set_nn_output_grad_to_zero:
                gate = torch.roll(torch.cumprod(x ** 3, dim=-1), 1, dims=-1)
                gate[:, 0] = 1
                p = gate[..., None] * p + zero
            return p / p.sum(-1, keepdim=True)

        calculate_bijection_params.allow_negative = allow_negative_params
        return params, calculate_bijection_params

    return masked_transform


def get_model(base_spline_degree=5, i_spline_degree=5, n_prior_internal_knots=15, n_i_internal_knots=15, i_spline_reg=0,
              i_spline_reverse_fun_tol=0.000001, n_flow_layers=1, prior_constraint_dict_left={},
              prior_constraint_dict_right={}, i_constraint_dict_left={}, i_constraint_dict_right={},
              set_nn_output_grad_to_zero=False, cached_bases_root='./cached_splines_bases'):
    """MFlow of (IMADE, Reverse) x L (model_factory.py:96-116)."""
    root = (lambda k: None) if cached_bases_root is None else (lambda k: f"{cached_bases_root}/{k}/")
    return flows.MFlow(
        flows.Serial(*(flows.IMADE(get_masked_transform(), spline_degree=i_spline_degree, n_internal_knots=n_i_internal_knots,
                                   spline_regularization=i_spline_reg, reverse_fun_tol=i_spline_reverse_fun_tol,
                                   constraints_dict_left=i_constraint_dict_left, constraints_dict_right=i_constraint_dict_right,
                                   set_nn_output_grad_to_zero=set_nn_output_grad_to_zero,
                                   cached_bases_path_root=root("I")), flows.Reverse()) * n_flow_layers),
        get_masked_transform(), spline_degree=base_spline_degree, n_internal_knots=n_prior_internal_knots,
        constraints_dict_left=prior_constraint_dict_left, constraints_dict_right=prior_constraint_dict_right,
        set_nn_output_grad_to_zero=set_nn_output_grad_to_zero, cached_bases_path_root=root("M"))


def get_waveflow_model(n_dimension, base_spline_degree=5, i_spline_degree=5, n_prior_internal_knots=16,
                       n_i_internal_knots=16, i_spline_reg=0, i_spline_reverse_fun_tol=0.000001, n_flow_layers=1,
                       box_size=1, xu_coord_type='mean', cached_bases_root='./cached_splines_bases'):
    """Waveflow of BoxTransformLayer + (IMADE, Reverse) x L with a conditional B-spline prior (model_factory.py:121-146)."""
    if xu_coord_type == 'mean':
        cons_left = list(range(0, n_dimension - 1))
    else:
        cons_left = list(range(1, n_dimension))
    root = (lambda k: None) if cached_bases_root is None else (lambda k: f"{cached_bases_root}/{k}/")
    return wavefunctions.Waveflow(
        flows.Serial(flows.BoxTransformLayer(box_size, xu_coord_type=xu_coord_type),
                     *(flows.IMADE(get_masked_transform(), spline_degree=i_spline_degree, n_internal_knots=n_i_internal_knots,
                                   spline_regularization=i_spline_reg, reverse_fun_tol=i_spline_reverse_fun_tol,
                                   constraints_dict_left={0: 0}, constraints_dict_right={0: 1},
                                   set_nn_output_grad_to_zero=False, cached_bases_path_root=root("I")),
                       flows.Reverse()) * n_flow_layers),
        get_masked_transform(allow_negative_params=True), spline_degree=base_spline_degree,
        n_internal_knots=n_prior_internal_knots, constraints_dict_left={0: 0}, constraints_dict_right={0: 0},
        constrained_dimension_indices_left=cons_left, constrained_dimension_indices_right=[],
        set_nn_output_grad_to_zero=False, cached_bases_path_root=root("B"))
