"""ISpline_fun / MSpline_fun / BSpline_fun -- same closure protocol and return tuples as the reference
(splines/isplines_jax.py:84-207, msplines_jax.py:67-196, bsplines_jax.py:52-203), backed by the C-ABI kernels.

Differences forced by the host framework: arrays are torch CUDA tensors, `rng` is a torch.Generator or an int seed
(JAX's threefry stream cannot be reproduced), only cardinal splines with cached bases are supported (the reference's
other branches exit()).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _ffi
from .._ffi import check, lib, ptr, stream_ptr, f32
from .tables import SplineTables


def _gen(rng):
    if isinstance(rng, torch.Generator):
        return rng
    g = torch.Generator()
    g.manual_seed(int(rng) if rng is not None else 0)
    return g


def _uniform(rng, shape, lo, hi):
    return torch.rand(shape, generator=_gen(rng), dtype=torch.float32) * (hi - lo) + lo


def _bc_arrays(tabs: SplineTables, cdict: dict, right: bool):
    """Constraint dictionary -> (n, nd[], val[], boundary-values[n*4]) as wf_enforce_bc wants them."""
    nds, vals, bv = [], [], []
    P = tabs.P
    for nd, val in cdict.items():
        nd = int(nd)
        if nd > 3:
            raise _ffi.WaveflowB200Error("boundary constraints are limited to derivative orders 0..3 (4 tables)")
        nds.append(nd); vals.append(float(val))
        row = [tabs.boundary_value(nd, (P - 1 - j) if right else j, right) for j in range(nd + 1)] + [0.0] * (3 - nd)
        bv.extend(row)
    return (len(nds), np.asarray(nds, dtype=np.int32), np.asarray(vals, dtype=np.float32),
            np.asarray(bv if bv else [0.0], dtype=np.float32))


def _enforce_bc(tabs: SplineTables, kind: str, left: dict, right: dict, w: torch.Tensor) -> torch.Tensor:
    w = f32(w)
    M, P = w.shape
    out = torch.empty_like(w)
    nl, ndl, vl, bl = _bc_arrays(tabs, left, False)
    nr, ndr, vr, br = _bc_arrays(tabs, right, True)
    st = lib.wf_enforce_bc(_ffi.KIND[kind], P, nl, _ffi.np_ptr(ndl), _ffi.np_ptr(vl), _ffi.np_ptr(bl), nr,
                           _ffi.np_ptr(ndr), _ffi.np_ptr(vr), _ffi.np_ptr(br), ptr(w), M, ptr(out), stream_ptr())
    check(st, "wf_enforce_bc")
    return out


def _remove_bias(kind: str, k: int, p: torch.Tensor) -> torch.Tensor:
    p = f32(p)
    out = torch.empty_like(p)
    check(lib.wf_remove_bias(_ffi.KIND[kind], k, p.shape[1], ptr(p), p.shape[0], ptr(out), stream_ptr()), "wf_remove_bias")
    return out


def spline_apply(tabs: SplineTables, c: torch.Tensor, x: torch.Tensor, nd0: int = 0, n_out: int = 1, logd: bool = False,
                 dense_key: str = "dense", force_dense: bool = False):
    """sum_q c[m,q] basis_q^{(nd0+k)}(x[m]) for k < n_out (+ log(out[1] + 1e-7)).  Returns a list of tensors."""
    c, x = f32(c), f32(x).reshape(-1)
    M, P = c.shape
    if x.shape[0] != M or P != tabs.P:
        raise _ffi.WaveflowB200Error(f"shape mismatch: params {tuple(c.shape)}, x {tuple(x.shape)}, bases {tabs.P}")
    d = tabs.dev(c.device)
    outs = [torch.empty(M, dtype=torch.float32, device=c.device) for _ in range(n_out)]
    lg = torch.empty(M, dtype=torch.float32, device=c.device) if logd else None
    local_ok = (not force_dense and dense_key == "dense" and d["rec"] is not None and nd0 == 0 and n_out <= 2
                and c.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0)
    if local_ok:
        st = lib.wf_spline_apply_local(ptr(d["rec"]), ptr(d["lo"]), ptr(d["dense"]), _ffi.KIND[tabs.kind], tabs.T, P,
                                       ptr(c), ptr(x), M, ptr(outs[0]), ptr(outs[1]) if n_out > 1 else None, ptr(lg),
                                       stream_ptr())
        check(st, "wf_spline_apply_local")
    else:
        arr = (C.c_void_p * n_out)(*[o.data_ptr() for o in outs])
        st = lib.wf_spline_apply_dense(ptr(d[dense_key]), tabs.T, P, ptr(c), ptr(x), M, nd0, n_out, arr, ptr(lg), stream_ptr())
        check(st, "wf_spline_apply_dense")
    return outs + ([lg] if logd else [])


# ============================================================================================ I-splines
def ISpline_fun():
    def init_fun(rng, k, n_internal_knots, cardinal_splines=True, zero_border=True, reverse_fun_tol=None,
                 use_cached_bases=True, cached_bases_path_root='./cached_splines_bases/I/', n_mesh_points=1000,
                 constraints_dict_left={0: 0.0}, constraints_dict_right={0: 1.0}):
        if not (use_cached_bases and cardinal_splines):
            raise NotImplementedError("only cardinal splines with cached bases are supported (isplines_jax.py:106-110)")
        if reverse_fun_tol is None:
            reverse_fun_tol = 1 / n_mesh_points
        tabs = SplineTables.get("I", k, n_internal_knots, n_mesh_points, cached_bases_path_root)
        n_bases = tabs.P
        n_par = n_bases - 2 if zero_border else n_bases
        initial_params = _uniform(rng, (n_par,), 0.0, 1.0).abs()
        initial_params = initial_params / initial_params.sum()

        def _pad(params):
            # zero_border: coefficient i multiplies basis i+1 (isplines_jax.py:72-75)
            if not zero_border:
                return f32(params)
            p = f32(params)
            z = torch.zeros(p.shape[0], 1, dtype=p.dtype, device=p.device)
            return torch.cat([z, p, z], dim=1).contiguous()

        def apply_fun_vec(params, x):
            return spline_apply(tabs, _pad(params), x, 0, 1)[0]

        def apply_fun_vec_grad(params, x):
            return spline_apply(tabs, _pad(params), x, 1, 1)[0]

        def apply_fun_vec_fused(params, x):
            """(value, derivative, log(derivative + 1e-7)) in one pass (extension; made.py:75-79 needs all three)."""
            return tuple(spline_apply(tabs, _pad(params), x, 0, 2, logd=True))

        def reverse_fun_vec(params, y):
            c, yv = _pad(params), f32(y).reshape(-1)
            out = torch.empty_like(yv)
            st = lib.wf_spline_reverse(ptr(tabs.dev(c.device)["dense"]), tabs.T, tabs.P, ptr(c), ptr(yv), yv.shape[0],
                                       float(reverse_fun_tol), ptr(out), None, stream_ptr())
            check(st, "wf_spline_reverse")
            return out

        def enforce_boundary_conditions(weights):
            return _enforce_bc(tabs, "I", constraints_dict_left, constraints_dict_right, weights)

        def remove_bias(params):
            return _remove_bias("I", k, params)

        apply_fun_vec.fused = apply_fun_vec_fused
        apply_fun_vec.tables = tabs
        knots = torch.from_numpy(np.asarray(tabs.knots, dtype=np.float32))
        return initial_params, apply_fun_vec, apply_fun_vec_grad, reverse_fun_vec, knots, enforce_boundary_conditions, remove_bias

    return init_fun


# ============================================================================================ M-splines
def MSpline_fun():
    def init_fun(rng, k, n_internal_knots, cardinal_splines=True, zero_border=False, use_cached_bases=True,
                 cached_bases_path_root='./cached_splines_bases/M/', n_mesh_points=1000,
                 constraints_dict_left={0: 0}, constraints_dict_right={0: 0}):
        if not (use_cached_bases and cardinal_splines):
            raise NotImplementedError("only cardinal splines with cached bases are supported (msplines_jax.py:84-88)")
        tabs = SplineTables.get("M", k, n_internal_knots, n_mesh_points, cached_bases_path_root)
        n_knots = len(tabs.knots)
        n_par = n_knots - k - 2 if zero_border else n_knots - k
        initial_params = _uniform(rng, (n_par,), 0.0, 1.0)
        initial_params = initial_params / initial_params.sum()

        def _pad(params):
            if not zero_border:
                return f32(params)
            p = f32(params)
            z = torch.zeros(p.shape[0], 1, dtype=p.dtype, device=p.device)
            return torch.cat([z, p, z], dim=1).contiguous()

        def apply_fun_vec(params, x):
            return spline_apply(tabs, _pad(params), x, 0, 1)[0]

        def apply_fun_vec_grad(params, x):
            return spline_apply(tabs, _pad(params), x, 1, 1)[0]

        def sample_fun_vec(rng_array, params, num_samples):
            return spline_sample(tabs, "M", rng_array, _pad(params), num_samples, n_knots=n_knots)

        def enforce_boundary_conditions(weights):
            return _enforce_bc(tabs, "M", constraints_dict_left, constraints_dict_right, weights)

        def remove_bias(params):
            return _remove_bias("M", k, params)

        apply_fun_vec.tables = tabs
        knots = torch.from_numpy(np.asarray(tabs.knots, dtype=np.float32))
        return initial_params, apply_fun_vec, apply_fun_vec_grad, sample_fun_vec, knots, enforce_boundary_conditions, remove_bias

    return init_fun


def spline_sample(tabs, kind: str, rng_array, params: torch.Tensor, num_samples: int, n_knots: int = 0) -> torch.Tensor:
    """sample_fun_vec (msplines_jax.py:129-154, bsplines_jax.py:144-171): params [M, P] -> rejection samples [M, num_samples].
    rng_array: an int seed, a torch.Generator, or an int64 tensor [M] of per-row keys (the reference passes one PRNG key per
    row); wf_spline_sample, one thread per (row, sample)."""
    from .._live import seed_of
    params = f32(params)
    Mrows, P = params.shape
    if P != tabs.P:
        raise _ffi.WaveflowB200Error(f"expected {tabs.P} spline coefficients per row, got {P}")
    d = tabs.dev(params.device)
    keys, seed = None, 0
    if isinstance(rng_array, torch.Tensor) and rng_array.numel() > 1:
        keys = rng_array.reshape(Mrows, -1)[:, -1].to(device=params.device, dtype=torch.int64).contiguous()
    else:
        seed = seed_of(rng_array if not isinstance(rng_array, torch.Tensor) else int(rng_array.reshape(-1)[0]))
    out = torch.empty(Mrows, int(num_samples), dtype=torch.float32, device=params.device)
    if kind == "B":
        st = lib.wf_spline_sample(ptr(d["ob_dense"]), _ffi.KIND["B"], tabs.T, P, ptr(d["ob_to_b"]), ptr(d["b_to_ob"]), 0, ptr(params), Mrows,
                                  int(num_samples), C.c_uint64(seed & (2 ** 64 - 1)), ptr(keys), ptr(out), stream_ptr())
    else:
        st = lib.wf_spline_sample(ptr(d["dense"]), _ffi.KIND["M"], tabs.T, P, None, None, int(n_knots), ptr(params), Mrows, int(num_samples),
                                  C.c_uint64(seed & (2 ** 64 - 1)), ptr(keys), ptr(out), stream_ptr())
    check(st, "wf_spline_sample")
    return out


# ============================================================================================ B-splines
def BSpline_fun():
    def init_fun(rng, k, n_internal_knots, cardinal_splines=True, use_cached_bases=True,
                 cached_bases_path_root='./cached_splines_bases/B/', n_mesh_points=1000,
                 constraints_dict_left={0: 0}, constraints_dict_right={0: 0}):
        if not (use_cached_bases and cardinal_splines):
            raise NotImplementedError("B spline basis only supports precached cardinal bases (bsplines_jax.py:68-72,117-119)")
        tabs = SplineTables.get("B", k, n_internal_knots, n_mesh_points, cached_bases_path_root)
        n_par = tabs.P
        initial_params = _uniform(rng, (n_par,), -1.0, 1.0)
        initial_params = initial_params / torch.sqrt((initial_params ** 2).sum())

        def _apply(params, x, nd):
            w, xv = f32(params), f32(x).reshape(-1)
            d = tabs.dev(w.device)
            out = torch.empty_like(xv)
            st = lib.wf_bspline_apply(ptr(d["ob_dense"]), ptr(d["ob_to_b"]), tabs.T, tabs.P, ptr(w), ptr(xv), xv.shape[0],
                                      nd, ptr(out), stream_ptr())
            check(st, "wf_bspline_apply")
            return out

        def apply_fun_vec(params, x):
            return _apply(params, x, 0)

        def apply_fun_vec_grad(params, x):
            return _apply(params, x, 1)

        def sample_fun_vec(rng_array, params, num_samples):
            return spline_sample(tabs, "B", rng_array, f32(params), num_samples)

        def enforce_boundary_conditions(weights):
            return _enforce_bc(tabs, "B", constraints_dict_left, constraints_dict_right, weights)

        apply_fun_vec.tables = tabs
        knots = torch.from_numpy(np.asarray(tabs.knots, dtype=np.float32))
        return initial_params, apply_fun_vec, apply_fun_vec_grad, sample_fun_vec, knots, enforce_boundary_conditions

    return init_fun
