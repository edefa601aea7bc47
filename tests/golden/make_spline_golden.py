"""Golden vectors for the table-spline operators, produced by the REFERENCE'S OWN SOURCE FILES (build container only):

    python tests/golden/make_spline_golden.py        # -> tests/golden/ref_spline_vectors.npz

waveflow/splines/isplines_jax.py (and utils/helpers.py's binary_search) are imported unmodified from /root/reference with
tests/golden/jax_numpy_shim first on sys.path (numpy stand-in for jax: jit = identity, vmap = loop, grad = first-order forward
mode through the file's own custom_jvp rule, lax.while_loop = Python loop with float32 state).  ISpline_fun is initialised the
way IMADE does (flows/bijections/made.py:51-60: zero_border=False, cached bases, constraints {0: 0} | {0: 1}) on the basis
tables the reference ships (tests/splines/cached_bases/I: degree 5, 16 internal knots, 2000 mesh points), then remove_bias,
enforce_boundary_conditions, apply_fun_vec, apply_fun_vec_grad and reverse_fun_vec are called on seeded inputs.  The same
for bsplines_jax.BSpline_fun (the orthonormalised B prior of wavefunctions.py:19-26): enforce_boundary_conditions, apply, grad.
"""
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def main():
    sys.path.insert(0, str(HERE / "jax_numpy_shim"))
    sys.path.insert(1, str(REF))
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    from waveflow.splines import isplines_jax

    k, n_internal, T, tol = 5, 16, 2000, 1e-4
    root = str(REF / "waveflow/tests/splines/cached_bases/I") + "/"
    (params_i, apply_vec, apply_vec_grad, reverse_vec, knots, enforce_bc, remove_bias) = isplines_jax.ISpline_fun()(
        7, k, n_internal, use_cached_bases=True, cardinal_splines=True, zero_border=False, reverse_fun_tol=tol,
        cached_bases_path_root=root, n_mesh_points=T, constraints_dict_left={0: 0.0}, constraints_dict_right={0: 1.0})
    P = int(params_i.shape[0])
    rng = np.random.default_rng(99)
    n = 600
    raw = rng.uniform(0.02, 1.0, (n, P)).astype(np.float32)            # positive conditioner outputs (sigmoid-like)
    p_rb = np.asarray(remove_bias(raw))
    p_bc = np.asarray(enforce_bc(p_rb))
    x = rng.uniform(0.0, 1.0, n).astype(np.float32)
    x[:6] = np.float32([0.0, 1.0, 0.5, 1.0 / (T - 1), 1.0 - 1e-7, 1e-7])  # mesh nodes, interval ends
    y = np.asarray(apply_vec(p_bc, x))
    dy = np.asarray(apply_vec_grad(p_bc, x))
    xr = np.asarray(reverse_vec(p_bc, y))
    out = dict(k=np.int32(k), n_internal=np.int32(n_internal), T=np.int32(T), tol=np.float32(tol), P=np.int32(P),
               raw=raw, remove_bias=p_rb, enforce_bc=p_bc, x=x, apply=y, apply_grad=dy, reverse=xr)
    # ---- B-spline prior (wavefunctions.py:19-26: BSpline_fun with the Waveflow default constraints {0: 0, 2: 0} | {0: 0})
    from waveflow.splines import bsplines_jax
    rootB = str(REF / "waveflow/tests/splines/cached_bases/B") + "/"
    (b_init, b_apply_vec, b_apply_vec_grad, _b_sample, _b_knots, b_enforce_bc) = bsplines_jax.BSpline_fun()(
        5, k, n_internal, cardinal_splines=True, use_cached_bases=True, cached_bases_path_root=rootB, n_mesh_points=T,
        constraints_dict_left={0: 0, 2: 0}, constraints_dict_right={0: 0})
    PB = int(b_init.shape[0])
    wB = rng.uniform(-1.0, 1.0, (n, PB)).astype(np.float32)
    wB_bc = np.asarray(b_enforce_bc(wB))
    xB = rng.uniform(0.0, 1.0, n).astype(np.float32)
    xB[:4] = np.float32([0.0, 1.0, 0.5, 3.0 / (T - 1)])
    out.update(B_P=np.int32(PB), B_raw=wB, B_enforce_bc=wB_bc, B_x=xB, B_apply=np.asarray(b_apply_vec(wB_bc, xB)),
               B_apply_grad=np.asarray(b_apply_vec_grad(wB_bc, xB)))
    for name, v in out.items():
        v = np.asarray(v)
        assert v.dtype != np.float64, (name, v.dtype)
    np.savez_compressed(HERE / "ref_spline_vectors.npz", **out)
    print("wrote", HERE / "ref_spline_vectors.npz", "P =", P, {k_: np.asarray(v).shape for k_, v in out.items() if np.asarray(v).ndim})


if __name__ == "__main__":
    main()
