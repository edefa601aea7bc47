"""Normal / Uniform priors, Flow and MFlow -- reference: flows/distributions.py:8-41,67-194 (same closure protocol)."""
from __future__ import annotations

import math

import torch

from .. import _live
from .._ffi import WaveflowB200Error, f32
from ..splines.factories import MSpline_fun, _gen
from .bijections import _split_rng


def Normal(offset=0.0):
    def init_fun(rng, input_dim):
        def log_pdf(params, inputs):
            z = f32(inputs) + offset
            return (-0.5 * z * z - 0.5 * math.log(2 * math.pi)).sum(1)

        def sample(rng, params, num_samples=1, device="cuda"):
            return torch.randn(num_samples, input_dim, generator=_gen(rng)).to(device)

        return (), log_pdf, sample

    return init_fun


def Uniform():
    def init_fun(rng, input_dim):
        def log_pdf(params, inputs):
            u = f32(inputs)
            inside = ((u >= 0) & (u <= 1)).all(dim=1)
            return torch.where(inside, torch.zeros_like(u[:, 0]), torch.full_like(u[:, 0], -float("inf")))

        def sample(rng, params, num_samples=1, device="cuda"):
            return torch.rand(num_samples, input_dim, generator=_gen(rng)).to(device)

        log_pdf.wf_prior = "uniform"
        return (), log_pdf, sample

    return init_fun


def Flow(transformation, prior=Normal(), prior_support=None):
    """distributions.py:67-112."""

    def init_fun(rng, input_dim):
        transformation_rng, prior_rng = _split_rng(rng)
        params, direct_fun, inverse_fun = transformation(transformation_rng, input_dim)
        prior_params, prior_log_pdf, prior_sample = prior(prior_rng, input_dim)

        def log_pdf(params, inputs, return_sample=False):
            u, log_det = direct_fun(params, inputs)
            if prior_support is not None:
                u = torch.clamp(u, *prior_support)
            log_probs = prior_log_pdf(prior_params, u)
            if return_sample:
                return log_probs + log_det, u
            return log_probs + log_det

        def sample(rng, params, num_samples=1, return_original_samples=False, device="cuda"):
            prior_samples = prior_sample(rng, prior_params, num_samples, device=device)
            out = inverse_fun(params, prior_samples)[0]
            if return_original_samples:
                return out, prior_samples
            return out

        return params, log_pdf, sample

    return init_fun


def MFlow(transformation, sp_transformation, spline_degree, n_internal_knots, constraints_dict_left={0: 0},
          constraints_dict_right={0: 0}, set_nn_output_grad_to_zero=False, n_spline_base_mesh_points=2000,
          cached_bases_path_root='./cached_splines_bases/M/'):
    """distributions.py:116-194: flow + conditional M-spline prior."""

    def init_fun(rng, input_dim):
        rng, transformation_rng = _split_rng(rng)
        rng, sp_transformation_rng = _split_rng(rng)
        transform_params, direct_fun, partial_inverse_fun = transformation(transformation_rng, input_dim)
        (prior_params_init, mspline_apply_fun_vec, _g, mspline_sample_fun_vec, knots, enforce_boundary_conditions,
         remove_bias) = MSpline_fun()(rng, spline_degree, n_internal_knots, zero_border=False, cardinal_splines=True,
                                      use_cached_bases=True, n_mesh_points=n_spline_base_mesh_points,
                                      cached_bases_path_root=cached_bases_path_root,
                                      constraints_dict_left=constraints_dict_left,
                                      constraints_dict_right=constraints_dict_right)
        P = prior_params_init.shape[0]
        sp_params_init, sp_transform_apply_fun = sp_transformation(transformation_rng, input_dim, P,
                                                                   set_nn_output_grad_to_zero=set_nn_output_grad_to_zero)
        flow_spec = getattr(direct_fun, "wf_spec", None)
        spec = None
        if flow_spec is not None and not set_nn_output_grad_to_zero and not getattr(sp_transform_apply_fun, "allow_negative", False):
            import copy
            spec = copy.copy(flow_spec)
            spec.prior, spec.tab_P, spec.k_P = "M", mspline_apply_fun_vec.tables, spline_degree
            spec.bc_P_left, spec.bc_P_right = dict(constraints_dict_left), dict(constraints_dict_right)
            if not spec.fusible():
                spec = None

        def _prior_coeffs(sp_transform_params, u):
            pp = sp_transform_apply_fun(sp_transform_params, u)
            pp = remove_bias(pp.reshape(-1, P))
            return enforce_boundary_conditions(pp)

        def log_pdf(params, inputs, return_sample=False):
            x = f32(inputs)
            if x.dim() == 1:
                x = x[None]
            tp, sp = params
            if spec is not None:
                w = _live.packed_for(spec, tp, sp, x.device)
                out = _live.forward(spec, w, x, want=("u", "logpdf") if return_sample else ("logpdf",))
                if return_sample:
                    return out["logpdf"], torch.clamp(out["u"], 0.0, 1.0)
                return out["logpdf"]
            u, log_det = direct_fun(tp, x)                                        # operator-boundary path
            c = _prior_coeffs(sp, u)
            u = torch.clamp(u, 0.0, 1.0)
            probs = mspline_apply_fun_vec(c, u.reshape(-1)).reshape(u.shape[0], -1)
            log_probs = torch.log(probs + 1e-7).sum(-1)
            if return_sample:
                return log_probs + log_det, u
            return log_probs + log_det

        def sample(rng, params, num_samples=1, return_original_samples=False, device="cuda", exact_inverse=False):
            if spec is None:
                raise WaveflowB200Error("MFlow.sample needs the fused configuration built by model_factory.get_model")
            tp, sp = params
            w = _live.packed_for(spec, tp, sp, torch.device(device))
            x, u = _live.sample(spec, w, _live.seed_of(rng), num_samples, torch.device(device), exact=exact_inverse)
            if return_original_samples:
                return x, u
            return x

        log_pdf.wf_spec = spec
        return (transform_params, sp_params_init), log_pdf, sample

    return init_fun
