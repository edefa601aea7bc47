"""Waveflow: square-normalised wavefunction psi = prod_d phi_d(u_d) * exp(log|det J| / 2)
-- reference: wavefunctions.py:9-112 (same closure protocol: init_fun -> (params, psi, log_pdf, sample))."""
from __future__ import annotations

import copy

import torch

from . import _live
from ._ffi import WaveflowB200Error, f32
from .flows.bijections import _split_rng
from .splines.factories import BSpline_fun


def Waveflow(transformation, sp_transformation, spline_degree, n_internal_knots, constraints_dict_left={0: 0, 2: 0},
             constraints_dict_right={0: 0}, constrained_dimension_indices_left=(), constrained_dimension_indices_right=(),
             set_nn_output_grad_to_zero=True, n_spline_base_mesh_points=2000,
             cached_bases_path_root='./cached_splines_bases/B/'):
    def init_fun(rng, input_dim):
        rng, transformation_rng = _split_rng(rng)
        rng, sp_transformation_rng = _split_rng(rng)
        transform_params, direct_fun, partial_inverse_fun = transformation(transformation_rng, input_dim)
        (prior_params_init, bspline_apply_fun_vec, _g, _sample_vec, knots,
         enforce_boundary_conditions) = BSpline_fun()(rng, spline_degree, n_internal_knots, cardinal_splines=True,
                                                      use_cached_bases=True, n_mesh_points=n_spline_base_mesh_points,
                                                      cached_bases_path_root=cached_bases_path_root,
                                                      constraints_dict_left=constraints_dict_left,
                                                      constraints_dict_right=constraints_dict_right)
        P = prior_params_init.shape[0]
        sp_params_init, sp_transform_apply_fun = sp_transformation(transformation_rng, input_dim, P,
                                                                   set_nn_output_grad_to_zero=set_nn_output_grad_to_zero)
        cons = [int(i) for i in constrained_dimension_indices_left]
        flow_spec = getattr(direct_fun, "wf_spec", None)
        spec = None
        if flow_spec is not None and not set_nn_output_grad_to_zero and getattr(sp_transform_apply_fun, "allow_negative", False):
            expect = list(range(0, input_dim - 1)) if flow_spec.coord == "mean" else list(range(1, input_dim))
            if cons == expect and flow_spec.box is not None:        # model_factory.py:124-129
                spec = copy.copy(flow_spec)
                spec.prior, spec.tab_P, spec.k_P = "B", bspline_apply_fun_vec.tables, spline_degree
                spec.bc_P_left, spec.bc_P_right = dict(constraints_dict_left), dict(constraints_dict_right)
                if not spec.fusible():
                    spec = None

        def _factors(params, x):
            """operator-boundary path (wavefunctions.py:57-65): -> (phi [N, D], log_det [N])."""
            tp, sp = params
            u, log_det = direct_fun(tp, x)
            pp = enforce_boundary_conditions(sp_transform_apply_fun(sp, u).reshape(-1, P))
            u = torch.clamp(u, 0.0, 1.0)
            return bspline_apply_fun_vec(pp, u.reshape(-1)).reshape(u.shape[0], -1), log_det, u

        def log_pdf(params, inputs, return_sample=False):
            x = f32(inputs)
            if x.dim() == 1:
                x = x[None]
            if spec is not None:
                w = _live.packed_for(spec, params[0], params[1], x.device)
                out = _live.forward(spec, w, x, want=("u", "logpdf") if return_sample else ("logpdf",))
                if return_sample:
                    return out["logpdf"], torch.clamp(out["u"], 0.0, 1.0)
                return out["logpdf"]
            phi, log_det, u = _factors(params, x)
            probs = phi ** 2
            probs[:, cons] = probs[:, cons] / 2
            lp = torch.log(probs + 1e-7).sum(-1) + log_det
            return (lp, u) if return_sample else lp

        def psi(params, inputs, log_tol=1e-7):
            x = f32(inputs)
            if x.dim() == 1:
                x = x[None]
            if spec is not None:
                w = _live.packed_for(spec, params[0], params[1], x.device)
                return _live.forward(spec, w, x, want=("psi",))["psi"]
            phi, log_det, _ = _factors(params, x)
            phi[:, cons] = phi[:, cons] / (2.0 ** 0.5)
            return phi.prod(-1) * torch.exp(0.5 * log_det)

        def sample(rng, params, num_samples=1, device="cuda", exact_inverse=False):
            if spec is None:
                raise WaveflowB200Error("Waveflow.sample needs the fused configuration built by get_waveflow_model")
            w = _live.packed_for(spec, params[0], params[1], torch.device(device), fold_prior=False)
            return _live.sample(spec, w, _live.seed_of(rng), num_samples, torch.device(device), exact=exact_inverse)[0]

        psi.wf_spec = spec
        log_pdf.wf_spec = spec
        return (transform_params, sp_params_init), psi, log_pdf, sample

    return init_fun
