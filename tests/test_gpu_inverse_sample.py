"""Inverse flow (reference-quirk and exact modes) and the rejection sampler, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import live
from tests.util import spec_from_live

pytestmark = pytest.mark.gpu


def _setup(m, seed, device, scale=3.0):
    from waveflow_b200 import _live
    params = fx.random_params(np.random.default_rng(seed), m, scale=scale)
    spec = spec_from_live(m)
    return params, spec, _live.pack_params(spec, params[0], params[1], device, fold_prior=False)


@pytest.mark.parametrize("kind", ["mflow", "waveflow"])
def test_inverse_reference_mode_matches_oracle(cuda, kind):
    """exact=0 reproduces IMADE.inverse_fun as written (quirk Q1: conditioned on the layer input) + helpers.binary_search."""
    from waveflow_b200 import _live
    m = fx.mflow_model(n_layers=3) if kind == "mflow" else fx.waveflow_model(2)
    params, spec, w = _setup(m, 0, cuda)
    u = np.random.default_rng(1).uniform(0.02, 0.98, (4000, 2)).astype(np.float32)
    got = _live.inverse(spec, w, torch.from_numpy(u).to(cuda), exact=False).cpu().numpy()
    ref = live.flow_inverse(m.cast(np.float32), fx.cast_params(params, np.float32)[0], u)
    scale = 1.0 if m.box is None else 2 * m.box
    d = np.abs(got - ref) / scale
    # bisection to tol = 1e-6 per layer: identical up to decisions taken at rounding level
    assert np.median(d) < 2e-6 and d.max() < 1e-4


@pytest.mark.parametrize("D,coord", [(2, "mean"), (3, "mean"), (4, "mean"), (3, "first"), (2, None)])
def test_exact_inverse_round_trip(cuda, D, coord):
    from waveflow_b200 import _live
    m = fx.mflow_model(n_layers=3) if coord is None else fx.waveflow_model(D, coord=coord)
    params, spec, w = _setup(m, D, cuda)
    rng = np.random.default_rng(2)
    if coord is None:
        x = rng.uniform(0.03, 0.97, (3000, 2)).astype(np.float32)
    else:
        x = np.sort(rng.uniform(-9.5, 9.5, (3000, D)), -1).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    u = _live.forward(spec, w, xt, want=("u",))["u"]
    xr = _live.inverse(spec, w, u, exact=True).cpu().numpy()
    scale = 1.0 if m.box is None else 2 * m.box
    err = np.abs(xr - x) / scale
    assert np.median(err) < 3e-6 and err.max() < 1e-3        # test_bijections.py:12-21 uses atol 1e-3
    # the reference-mode inverse is NOT an inverse for D > 1 (published reconstruction distances grow to 2.4e-2)
    xq = _live.inverse(spec, w, u, exact=False).cpu().numpy()
    if coord in (None,) or D == 2:
        assert np.abs(xq - x).max() / scale > 10 * err.max()


def _chi2_grid(u, density_fn, bins=8, sub=24):
    N = len(u)
    edges = np.linspace(0, 1, bins + 1)
    H, _, _ = np.histogram2d(u[:, 0], u[:, 1], bins=[edges, edges])
    g = (np.arange(bins * sub) + 0.5) / (bins * sub)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    p = density_fn(np.stack([gx.ravel(), gy.ravel()], -1)).reshape(bins * sub, bins * sub)
    E = p.reshape(bins, sub, bins, sub).mean(axis=(1, 3)) / bins ** 2 * N
    ok = E > 20
    return float((((H - E) ** 2) / np.maximum(E, 1e-9))[ok].sum()), int(ok.sum())


@pytest.mark.parametrize("kind", ["mflow", "waveflow"])
def test_sampler_draws_follow_the_prior_density(cuda, kind):
    from waveflow_b200 import _live
    m = fx.mflow_model(n_layers=2) if kind == "mflow" else fx.waveflow_model(2, n_layers=2)
    params, spec, w = _setup(m, 5, cuda, scale=2.0)
    N = 200000
    x, u = _live.sample(spec, w, 1234, N, cuda, exact=True)
    u = u.cpu().numpy().astype(np.float64)
    assert u.min() >= 0 and u.max() <= 1 and np.all(np.isfinite(x.cpu().numpy()))

    def dens(pts):
        f = live.prior_factors(m, params[1], pts)
        return np.prod(f ** 2 if kind == "waveflow" else f, axis=-1)
    chi2, dof = _chi2_grid(u, dens)
    assert chi2 < dof + 6 * np.sqrt(2 * dof) + 0.02 * N / 100, (chi2, dof)   # statistical + quadrature slack
    # determinism and stream independence
    x2, u2 = _live.sample(spec, w, 1234, N, cuda, exact=True)
    assert torch.equal(x, x2)
    x3, _ = _live.sample(spec, w, 1235, 1000, cuda, exact=True)
    assert not torch.equal(x3, x[:1000])
    # the data-space samples are the inverse flow of the draws
    xi = _live.inverse(spec, w, torch.from_numpy(u.astype(np.float32)).to(cuda), exact=True)
    assert torch.equal(xi, x)
    # exact-mode samples have density exp(log_pdf): E[log p(x)] is finite and the round trip recovers the draws
    ub = _live.forward(spec, w, x, want=("u",))["u"].cpu().numpy()
    assert np.median(np.abs(ub - u)) < 1e-5


@pytest.mark.parametrize("kind", ["M", "B"])
def test_standalone_sample_fun_vec(cuda, kind):
    """sample_fun_vec of MSpline_fun / BSpline_fun (msplines_jax.py:129-154, bsplines_jax.py:144-171) as an operator: the
    histogram of the draws of each row follows the row's own density (chi-square against the oracle's spline evaluation)."""
    from waveflow_b200.splines.factories import BSpline_fun, MSpline_fun
    rng = np.random.default_rng(3)
    if kind == "M":
        init, apply_vec, _g, sample_vec, knots, bc, rb = MSpline_fun()(0, 3, 15, cardinal_splines=True, zero_border=False,
                                                                      use_cached_bases=True, n_mesh_points=2000,
                                                                      cached_bases_path_root=None)
        P = init.shape[0]
        params = rng.uniform(0.05, 1.0, (3, P)).astype(np.float32)
        params /= params.sum(-1, keepdims=True)
        # the reference's bound ymax = max(c) * n_knots (msplines_jax.py:145-148) only dominates the density for coefficients
        # that went through remove_bias + the {0:0} boundary conditions, which is how MFlow.sample always calls it
        # (distributions.py:172-176); raw coefficients put up to 5x more mass at the ends than the bound allows
        params = bc(rb(torch.from_numpy(params).to(cuda))).cpu().numpy()
        dens = lambda p_, x_: apply_vec(p_, x_)
    else:
        init, apply_vec, _g, sample_vec, knots, bc = BSpline_fun()(0, 6, 23, cached_bases_path_root=None, n_mesh_points=2000)
        P = init.shape[0]
        params = rng.standard_normal((3, P)).astype(np.float32)
        dens = lambda p_, x_: apply_vec(p_, x_) ** 2
    n = 40000
    tp = torch.from_numpy(params).to(cuda)
    draws = sample_vec(7, tp, n)
    assert tuple(draws.shape) == (3, n)
    d = draws.cpu().numpy()
    assert d.min() >= 0.0 and d.max() <= 1.0
    again = sample_vec(7, tp, n)
    assert torch.equal(draws, again)                                  # counter-based streams: reproducible
    assert not torch.equal(draws, sample_vec(8, tp, n))
    edges = np.linspace(0, 1, 41)
    mids = torch.from_numpy(np.linspace(0, 1, 4001, dtype=np.float32)).to(cuda)
    for r in range(3):
        f = dens(tp[r:r + 1].expand(mids.shape[0], -1).contiguous(), mids).cpu().numpy().astype(np.float64)
        cdf = np.concatenate([[0], np.cumsum(0.5 * (f[1:] + f[:-1]))]); cdf /= cdf[-1]
        expect = np.diff(np.interp(edges, np.linspace(0, 1, 4001), cdf)) * n
        obs = np.histogram(d[r], bins=edges)[0]
        ok = expect > 20
        chi2 = float((((obs - expect) ** 2) / np.maximum(expect, 1e-9))[ok].sum())
        assert chi2 < 2.5 * ok.sum(), (kind, r, chi2, ok.sum())
    # per-row keys (the reference passes one PRNG key per row)
    keys = torch.arange(3, device=cuda, dtype=torch.int64)
    dk = sample_vec(keys, tp, 16)
    assert tuple(dk.shape) == (3, 16)
