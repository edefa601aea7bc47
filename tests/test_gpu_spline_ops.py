"""T2: operator-boundary parity (GPU through the C ABI vs the oracle) for the table splines."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import live
from tests.util import relerr

pytestmark = pytest.mark.gpu


def _coeffs(rng, M, P, k, tab64):
    p = rng.uniform(0, 1, (M, P))
    p = live.remove_bias_I(p, k)
    return live.enforce_bc(tab64, p, {0: 0}, {0: 1}, "I")


def _edge_x(rng, n):
    T = 2000
    nodes = rng.integers(0, T, 64) / (T - 1)
    special = np.array([0.0, 1.0, np.nextafter(np.float32(1), np.float32(2)), 0.5, 1.0 / 1999, 1998.0 / 1999,
                        np.nextafter(np.float32(0.5), np.float32(1)), 1e-8, 1 - 1e-7], dtype=np.float64)
    knots = np.linspace(0, 1, 23)
    return np.concatenate([rng.uniform(0, 1, n), nodes, special, knots]).astype(np.float32)


@pytest.mark.parametrize("k,n", [(6, 23), (5, 16), (5, 23)])
def test_ispline_apply_local_and_dense(cuda, k, n):
    from waveflow_b200.splines.factories import ISpline_fun
    from waveflow_b200.splines.factories import spline_apply
    rng = np.random.default_rng(0)
    tab64 = fx.tables_I(k, n)
    P = tab64.shape[1]
    x = _edge_x(rng, 20000)
    c = _coeffs(rng, len(x), P, k, tab64).astype(np.float32)
    init = ISpline_fun()(0, k, n, zero_border=False, n_mesh_points=2000, cached_bases_path_root=None,
                         reverse_fun_tol=1e-6)
    _, apply_v, apply_g, _, _, _, _ = init
    ct, xt = torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda)
    val, grad, logd = [t.cpu().numpy() for t in apply_v.fused(ct, xt)]
    v2 = apply_v(ct, xt).cpu().numpy()
    g2 = apply_g(ct, xt).cpu().numpy()
    # dense path, same inputs
    vd, gd, ld = [t.cpu().numpy() for t in spline_apply(apply_v.tables, ct, xt, 0, 2, logd=True, force_dense=True)]
    # local-support reformulation only drops exact zeros / adds exact ones in the same order: bit-identical
    assert np.array_equal(val, vd) and np.array_equal(grad, gd) and np.array_equal(logd, ld)
    assert np.array_equal(val, v2) and np.array_equal(gd, g2)
    # oracle (float64) on the same float32 inputs and float32 tables
    tab = tab64.astype(np.float32).astype(np.float64)
    ov = live.spline_apply(tab, c.astype(np.float64), x.astype(np.float64), 0)
    og = live.spline_apply(tab, c.astype(np.float64), x.astype(np.float64), 1)
    # the table index must be the reference's float32 one: evaluate it in float32 for the oracle too
    ov32 = live.spline_apply(tab64.astype(np.float32), c, x, 0)
    og32 = live.spline_apply(tab64.astype(np.float32), c, x, 1)
    assert relerr(val, ov32, 1.0) < 1e-6 and relerr(grad, og32, np.abs(og32).max()) < 1e-6
    ok = np.abs(ov - ov32) < 1e-5           # float64 index == float32 index except exactly at mesh nodes
    assert relerr(val[ok], ov[ok], 1.0) < 1e-5
    assert relerr(logd, np.log(og32.astype(np.float64) + 1e-7), 1.0) < 1e-5


def test_spline_apply_outside_unit_interval_matches_reference_gather_semantics(cuda):
    """x < 0 wraps the left index to T-1, x > 1 clamps (JAX gather)."""
    from waveflow_b200.splines.factories import spline_apply
    from waveflow_b200.splines.tables import SplineTables
    rng = np.random.default_rng(1)
    tabs = SplineTables.get("I", 6, 23)
    x = np.array([-1e-4, -3e-4, -0.2, 1.0002, 1.3, 7.0, -5.0], dtype=np.float32)
    c = rng.uniform(0, 1, (len(x), tabs.P)).astype(np.float32)
    for force in (False, True):
        v, g = [t.cpu().numpy() for t in spline_apply(tabs, torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda),
                                                      0, 2, force_dense=force)]
        assert relerr(v, live.spline_apply(tabs.tab32, c, x, 0), 1.0) < 1e-5
        assert relerr(g, live.spline_apply(tabs.tab32, c, x, 1), 1.0) < 1e-5


def test_higher_derivative_tables_and_clamp(cuda):
    from waveflow_b200.splines.factories import spline_apply
    from waveflow_b200.splines.tables import SplineTables
    rng = np.random.default_rng(2)
    tabs = SplineTables.get("I", 6, 23)
    x = rng.uniform(0, 1, 4096).astype(np.float32)
    c = rng.uniform(0, 1, (4096, tabs.P)).astype(np.float32)
    outs = spline_apply(tabs, torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda), 1, 4)
    for kk, o in enumerate(outs):
        ref = live.spline_apply(tabs.tab32, c, x, 1 + kk)            # nd 4 clamps to 3 (quirk Q5)
        assert relerr(o.cpu().numpy(), ref, np.abs(ref).max()) < 2e-6
    assert torch.equal(outs[2], outs[3])


@pytest.mark.parametrize("kind", ["I", "M"])
def test_remove_bias(cuda, kind):
    from waveflow_b200.splines.factories import ISpline_fun, MSpline_fun
    rng = np.random.default_rng(3)
    k, n = (6, 23) if kind == "I" else (3, 15)
    init = (ISpline_fun() if kind == "I" else MSpline_fun())(0, k, n, n_mesh_points=2000, cached_bases_path_root=None)
    rb = init[-1]
    P = n + k if kind == "I" else n + k - 2
    p = rng.uniform(0, 1, (5000, P)).astype(np.float32)
    ref = (live.remove_bias_I if kind == "I" else live.remove_bias_M)(p, k)
    got = rb(torch.from_numpy(p).to(cuda)).cpu().numpy()
    assert relerr(got, ref, 1.0) < 5e-7
    assert np.allclose(got.sum(-1), 1, atol=1e-5)


def test_enforce_boundary_conditions_all_kinds(cuda):
    from waveflow_b200.splines.factories import BSpline_fun, ISpline_fun, MSpline_fun
    rng = np.random.default_rng(4)
    # I: model default {0:0}|{0:1} and the test_boundary_constraints.py set {0:0, 2:0, 3:0}
    for left, right in [({0: 0.0}, {0: 1.0}), ({0: 0, 2: 0, 3: 0}, {0: 1.0}), ({}, {}), ({0: 0.0}, {0: 1.0, 1: 0.0, 2: 0.0})]:
        init = ISpline_fun()(0, 6, 23, zero_border=False, n_mesh_points=2000, cached_bases_path_root=None,
                             constraints_dict_left=left, constraints_dict_right=right)
        w = rng.uniform(0.1, 1, (3000, 29)).astype(np.float32)
        ref = live.enforce_bc(fx.tables_I(6, 23).astype(np.float32), w, left, right, "I")
        got = init[5](torch.from_numpy(w).to(cuda)).cpu().numpy()
        assert relerr(got, ref, np.abs(ref).max()) < 2e-6, (left, right)
    init = MSpline_fun()(0, 3, 15, n_mesh_points=2000, cached_bases_path_root=None, constraints_dict_left={0: 0, 1: 0},
                         constraints_dict_right={0: 0})
    w = rng.uniform(0.1, 1, (3000, 16)).astype(np.float32)
    ref = live.enforce_bc(fx.tables_M(3, 15).astype(np.float32), w, {0: 0, 1: 0}, {0: 0}, "M")
    assert relerr(init[5](torch.from_numpy(w).to(cuda)).cpu().numpy(), ref, np.abs(ref).max()) < 2e-6
    init = BSpline_fun()(0, 6, 23, n_mesh_points=2000, cached_bases_path_root=None, constraints_dict_left={0: 0, 2: 0},
                         constraints_dict_right={0: 0})
    w = rng.uniform(-1, 1, (3000, 28)).astype(np.float32)
    ref = live.enforce_bc(fx.tables_B(6, 23)["b"].astype(np.float32), w, {0: 0, 2: 0}, {0: 0}, "B")
    got = init[5](torch.from_numpy(w).to(cuda)).cpu().numpy()
    assert relerr(got, ref, np.abs(ref).max()) < 2e-6
    assert np.allclose((got ** 2).sum(-1), 1, atol=1e-5)
    with pytest.raises(Exception):       # isplines_jax.py:177-179: right {0: v != 1} is rejected
        ISpline_fun()(0, 6, 23, zero_border=False, n_mesh_points=2000, cached_bases_path_root=None,
                      constraints_dict_right={0: 0.5})[5](torch.from_numpy(w[:, :1].repeat(29, 1)).to(cuda))


def test_mspline_and_bspline_apply(cuda):
    from waveflow_b200.splines.factories import BSpline_fun, MSpline_fun
    rng = np.random.default_rng(5)
    x = _edge_x(rng, 8000)
    # M
    init = MSpline_fun()(0, 3, 15, n_mesh_points=2000, cached_bases_path_root=None)
    c = rng.uniform(0, 1, (len(x), 16)).astype(np.float32)
    tabM = fx.tables_M(3, 15).astype(np.float32)
    v = init[1](torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda)).cpu().numpy()
    g = init[2](torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda)).cpu().numpy()
    rv, rg = live.spline_apply(tabM, c, x, 0), live.spline_apply(tabM, c, x, 1)
    assert relerr(v, rv, np.abs(rv).max()) < 1e-6 and relerr(g, rg, np.abs(rg).max()) < 1e-6
    # B: c = w @ ob_to_b, normalised, orthonormalised tables
    Bt = fx.tables_B(6, 23)
    init = BSpline_fun()(0, 6, 23, n_mesh_points=2000, cached_bases_path_root=None)
    w = rng.uniform(-1, 1, (len(x), 28)).astype(np.float32)
    xc = np.clip(x, 0, 1)
    ref = live.spline_apply(Bt["ob"].astype(np.float32).astype(np.float64),
                            live.bspline_coeffs(w.astype(np.float64), Bt["ob_to_b"].astype(np.float32).astype(np.float64)),
                            xc.astype(np.float64), 0)
    ref32 = live.spline_apply(Bt["ob"].astype(np.float32), live.bspline_coeffs(w, Bt["ob_to_b"].astype(np.float32)), xc, 0)
    got = init[1](torch.from_numpy(w).to(cuda), torch.from_numpy(xc).to(cuda)).cpu().numpy()
    ok = np.abs(ref - ref32) < 1e-4
    assert relerr(got[ok], ref[ok], np.abs(ref).max()) < 1e-5
    gref = live.spline_apply(Bt["ob"].astype(np.float32), live.bspline_coeffs(w, Bt["ob_to_b"].astype(np.float32)), xc, 1)
    gg = init[2](torch.from_numpy(w).to(cuda), torch.from_numpy(xc).to(cuda)).cpu().numpy()
    assert relerr(gg, gref, np.abs(gref).max()) < 1e-5


def test_reverse_fun_vec_is_reference_bisection(cuda):
    from waveflow_b200 import _ffi
    from waveflow_b200.splines.factories import ISpline_fun
    rng = np.random.default_rng(6)
    k, n = 6, 23
    tab32 = fx.tables_I(k, n).astype(np.float32)
    init = ISpline_fun()(0, k, n, zero_border=False, n_mesh_points=2000, cached_bases_path_root=None, reverse_fun_tol=1e-6)
    M = 20000
    c = _coeffs(rng, M, 29, k, fx.tables_I(k, n)).astype(np.float32)
    xs = rng.uniform(0, 1, M).astype(np.float32)
    xs[:4] = [0.0, 1.0, 0.5, 1e-7]
    ct = torch.from_numpy(c).to(cuda)
    y = init[1](ct, torch.from_numpy(xs).to(cuda))
    xr = init[3](ct, y).cpu().numpy()
    ref = live.binary_search_inverse(tab32, c, y.cpu().numpy(), 1e-6)
    # lower bracket of a bisection with tol 1e-6: identical decisions except where f(mid) - y rounds across 0
    assert np.mean(xr == ref) > 0.98
    assert np.abs(xr - ref).max() <= 2e-6
    assert np.abs(xr - xs)[4:].max() < 3e-6          # inverse property (the lower bracket is within tol of x)
    # iteration count: 2^-20 < 1e-6/2... every lane runs the same fixed number of halvings
    out = torch.empty(M, dtype=torch.float32, device=cuda); it = torch.empty(M, dtype=torch.int32, device=cuda)
    tabs = init[1].tables
    _ffi.check(_ffi.lib.wf_spline_reverse(_ffi.ptr(tabs.dev(cuda)["dense"]), 2000, 29, _ffi.ptr(ct), _ffi.ptr(y), M, 1e-6,
                                          _ffi.ptr(out), _ffi.ptr(it), _ffi.stream_ptr()))
    assert int(it.min()) >= 19 and int(it.max()) <= 21


def test_empty_and_ragged_batches(cuda):
    from waveflow_b200.splines.factories import spline_apply
    from waveflow_b200.splines.tables import SplineTables
    tabs = SplineTables.get("I", 6, 23)
    rng = np.random.default_rng(7)
    for M in [0, 1, 3, 255, 256, 257, 256 * 148 + 5]:
        c = rng.uniform(0, 1, (M, tabs.P)).astype(np.float32)
        x = rng.uniform(0, 1, M).astype(np.float32)
        v, g = spline_apply(tabs, torch.from_numpy(c).to(cuda), torch.from_numpy(x).to(cuda), 0, 2)
        assert v.shape == (M,)
        if M:
            ref = live.spline_apply(tabs.tab32, c, x, 0)
            assert relerr(v.cpu().numpy(), ref, 1.0) < 1e-6


def test_spline_operators_against_vectors_from_the_reference_source(cuda):
    """The C-ABI spline operators (through the ISpline_fun / BSpline_fun mirrors) against tests/golden/ref_spline_vectors.npz:
    outputs of the reference's own isplines_jax.py / bsplines_jax.py / helpers.binary_search executed on a numpy stand-in for
    jax (tests/golden/make_spline_golden.py) on the reference's shipped tables (degree 5, 16 internal knots)."""
    from pathlib import Path
    from waveflow_b200.splines.factories import BSpline_fun, ISpline_fun
    G = np.load(Path(__file__).resolve().parent / "golden" / "ref_spline_vectors.npz")
    k, n, T, tol = int(G["k"]), int(G["n_internal"]), int(G["T"]), float(G["tol"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    _, apply_v, apply_g, reverse_v, _, enforce_bc, remove_bias = ISpline_fun()(
        0, k, n, zero_border=False, n_mesh_points=T, cached_bases_path_root=None, reverse_fun_tol=tol,
        constraints_dict_left={0: 0.0}, constraints_dict_right={0: 1.0})
    assert relerr(remove_bias(t(G["raw"])).cpu().numpy(), G["remove_bias"], 1.0) < 5e-7
    assert relerr(enforce_bc(t(G["remove_bias"])).cpu().numpy(), G["enforce_bc"], np.abs(G["enforce_bc"]).max()) < 2e-6
    c = t(G["enforce_bc"])
    assert relerr(apply_v(c, t(G["x"])).cpu().numpy(), G["apply"], 1.0) < 1e-6
    assert relerr(apply_g(c, t(G["x"])).cpu().numpy(), G["apply_grad"], np.abs(G["apply_grad"]).max()) < 1e-6
    xr = reverse_v(c, t(G["apply"])).cpu().numpy()
    assert np.mean(xr == G["reverse"]) > 0.98 and np.abs(xr - G["reverse"]).max() <= 2 * tol
    _, b_apply, b_grad, _, _, b_bc = BSpline_fun()(0, k, n, n_mesh_points=T, cached_bases_path_root=None,
                                                  constraints_dict_left={0: 0, 2: 0}, constraints_dict_right={0: 0})
    assert relerr(b_bc(t(G["B_raw"])).cpu().numpy(), G["B_enforce_bc"], np.abs(G["B_enforce_bc"]).max()) < 2e-6
    cb = t(G["B_enforce_bc"])
    assert relerr(b_apply(cb, t(G["B_x"])).cpu().numpy(), G["B_apply"], np.abs(G["B_apply"]).max()) < 3e-6
    assert relerr(b_grad(cb, t(G["B_x"])).cpu().numpy(), G["B_apply_grad"], np.abs(G["B_apply_grad"]).max()) < 3e-6
