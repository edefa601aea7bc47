"""T3/T4: fused live-path parity -- psi / log_pdf / u / log|det J| and the forward-Laplacian local energy."""
import numpy as np
import pytest
import torch

from oracle import fast_cpu
from oracle import fixtures as fx
from oracle import laplacian as olap
from oracle import live
from tests.util import assert_fp32_grade, record_flat, relerr, spec_from_live

pytestmark = pytest.mark.gpu


def _pack(spec, params, device):
    from waveflow_b200 import _live
    return _live.pack_params(spec, params[0], params[1], device)


def _grid():
    y, x = np.meshgrid(np.linspace(-10, 10, 100), np.linspace(-10, 10, 100))
    c = np.stack([x, y], -1).reshape(-1, 2)
    return c, (c[:, 0] > c[:, 1]).astype(int)


def test_psi_published_known_answer(cuda):
    """He checkpoint -> psi on the published 100x100 grid (utils/helpers.py:52-59), compared with the values the
    reference itself wrote (float32 JAX) and with the float64 oracle."""
    from waveflow_b200 import _live
    params, gold = fx.load_he_checkpoint()
    m64 = fx.waveflow_model(2)
    spec = spec_from_live(m64)
    w = _pack(spec, params, cuda)
    c, inv = _grid()
    sc = np.sort(c, -1).astype(np.float32)
    out = _live.forward(spec, w, torch.from_numpy(sc).to(cuda), want=("u", "logdet", "logpdf", "psi"))
    psi = out["psi"].cpu().numpy() * (-1.0) ** inv
    assert np.abs(psi - gold["psi_grid"]).max() < 3e-5                 # |psi| up to 1.53
    m32, p32 = fx.waveflow_model(2, dtype=np.float32), fx.cast_params(params, np.float32)
    ref = live.psi(m64, params, sc.astype(np.float64)) * (-1.0) ** inv
    assert_fp32_grade(psi, ref, live.psi(m32, p32, sc) * (-1.0) ** inv, 1e-5, name="psi")
    u64, ld64 = live.flow_direct(m64, params[0], sc.astype(np.float64))
    u32, ld32 = live.flow_direct(m32, p32[0], sc)
    assert_fp32_grade(out["u"].cpu().numpy(), u64, u32, 1e-5, 1.0, "u")
    assert_fp32_grade(out["logdet"].cpu().numpy(), ld64, ld32, 1e-5, 1.0, "logdet")
    lp64 = live.log_pdf(m64, params, sc.astype(np.float64))
    # log(phi^2 + 1e-7) amplifies float32 rounding without bound next to the nodes of psi: the yardstick there is the
    # reference's own float32 arithmetic (numpy restatement in float32) against the float64 oracle
    assert_fp32_grade(out["logpdf"].cpu().numpy(), lp64, live.log_pdf(m32, p32, sc), 1e-5, 1.0, "logpdf")


@pytest.mark.parametrize("D,coord", [(2, "mean"), (3, "mean"), (4, "mean"), (3, "first"), (4, "first")])
def test_waveflow_forward_random_params(cuda, D, coord):
    from waveflow_b200 import _live
    m = fx.waveflow_model(D, coord=coord)
    params = fx.random_params(np.random.default_rng(D), m, scale=3.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    x = np.sort(np.random.default_rng(1).uniform(-10, 10, (4001, D)), -1).astype(np.float32)
    out = _live.forward(spec, w, torch.from_numpy(x).to(cuda), want=("u", "logdet", "logpdf", "psi"))
    x64 = x.astype(np.float64)
    m32, p32 = m.cast(np.float32), fx.cast_params(params, np.float32)
    assert_fp32_grade(out["psi"].cpu().numpy(), live.psi(m, params, x64), live.psi(m32, p32, x), 1e-5, name="psi")
    assert_fp32_grade(out["logpdf"].cpu().numpy(), live.log_pdf(m, params, x64), live.log_pdf(m32, p32, x), 1e-5, 1.0, "logpdf")
    u64, ld64 = live.flow_direct(m, params[0], x64)
    u32, ld32 = live.flow_direct(m32, p32[0], x)
    assert_fp32_grade(out["u"].cpu().numpy(), u64, u32, 1e-5, 1.0, "u")
    assert_fp32_grade(out["logdet"].cpu().numpy(), ld64, ld32, 1e-5, 1.0, "logdet")


@pytest.mark.parametrize("bc", [({0: 0.0}, {0: 1.0}), ({}, {})])
def test_mflow_log_pdf(cuda, bc):
    """benchmark_tests.get_model('MFlow') shape: 3 x (IMADE, Reverse), I degree 5 / 23 knots, M prior degree 3 / 15 knots."""
    from waveflow_b200 import _live
    m = fx.mflow_model(bc_i_left=bc[0], bc_i_right=bc[1])
    params = fx.random_params(np.random.default_rng(0), m, scale=3.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    x = np.random.default_rng(1).uniform(0.025, 0.975, (5000, 2)).astype(np.float32)
    out = _live.forward(spec, w, torch.from_numpy(x).to(cuda), want=("u", "logdet", "logpdf"))
    lp64, u64 = live.log_pdf(m, params, x.astype(np.float64), return_sample=True)
    lp32 = live.log_pdf(m.cast(np.float32), fx.cast_params(params, np.float32), x)
    assert_fp32_grade(out["logpdf"].cpu().numpy(), lp64, lp32, 1e-5, 1.0, "logpdf")
    assert np.abs(np.clip(out["u"].cpu().numpy(), 0, 1) - u64).max() < 2e-5


def _check_energy(out, ref, psi32, tol_e=1e-4, ref32=None):
    psi, hpsi, eloc = [out[k].cpu().numpy() for k in ("psi", "hpsi", "eloc")]
    # north_star: local energies within 1e-4 relative -- recorded per walker (written to gpurun_out/parity_fractions.json)
    record_flat("E_loc (pointwise relative)", eloc, ref["eloc"], tol_e, None if ref32 is None else ref32["eloc"])
    record_flat("H psi (pointwise relative)", hpsi, ref["hpsi"], tol_e, None if ref32 is None else ref32["hpsi"])
    assert_fp32_grade(psi, ref["psi"], psi32, 1e-5, name="psi")
    assert relerr(out["grad"].cpu().numpy(), ref["grad"]) < tol_e
    assert relerr(out["lap"].cpu().numpy(), ref["lap"]) < tol_e
    assert relerr(hpsi, ref["hpsi"]) < tol_e
    # E_loc = H psi / (psi + 1e-8): first-order error budget  dE = dH / psi - E dpsi / psi  with the two float32-grade
    # bounds checked above (|dH| <= tol_e max|H psi|, |dpsi| <= 1e-5 max|psi|) -- i.e. 1e-4 relative wherever psi and
    # H psi are of the size of their batch maxima, and the unavoidable amplification next to the nodes of psi elsewhere
    apsi = np.abs(ref["psi"]) + 1e-30
    budget = tol_e * np.abs(ref["hpsi"]).max() / apsi + np.abs(ref["eloc"]) * 1e-5 * np.abs(ref["psi"]).max() / apsi
    assert np.all(np.abs(eloc - ref["eloc"]) <= 2 * budget + 1e-6)
    big = np.abs(ref["psi"]) > 1e-2 * np.abs(ref["psi"]).max()
    assert abs(eloc[big].astype(np.float64).mean() - ref["eloc"][big].mean()) <= tol_e * np.abs(ref["eloc"][big]).mean()
    return psi, eloc


def test_local_energy_he_checkpoint(cuda):
    from waveflow_b200 import _live
    params, gold = fx.load_he_checkpoint()
    m = fx.waveflow_model(2)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    prot = np.array([[0.0], [0.0]])
    rng = np.random.default_rng(5)
    x = np.concatenate([np.sort(gold["samples"], -1), np.sort(rng.uniform(-10, 10, (262, 2)), -1)]).astype(np.float32)
    sums = torch.zeros(4, dtype=torch.float64, device=cuda)
    out = _live.local_energy(spec, w, torch.from_numpy(x).to(cuda), prot, want=("psi", "hpsi", "eloc", "grad", "lap"), sums=sums)
    ref = olap.local_energy_bundle(m, params, x.astype(np.float64), prot)
    psi32 = live.psi(fx.waveflow_model(2, dtype=np.float32), fx.cast_params(params, np.float32), x)
    ref32 = fast_cpu.FastLocalEnergy(fx.waveflow_model(2, dtype=np.float32), params, prot, dtype=torch.float32)(x)
    psi, eloc = _check_energy(out, ref, psi32, ref32=ref32)
    s = sums.cpu().numpy()
    assert s[2] == len(x)
    assert abs(s[0] - eloc.astype(np.float64).sum()) <= 1e-6 * np.abs(eloc).sum()
    assert abs(s[1] - (eloc.astype(np.float64) ** 2).sum()) <= 1e-6 * (eloc.astype(np.float64) ** 2).sum()
    assert abs(s[3] - (psi.astype(np.float64) ** 2).sum()) <= 1e-6 * (psi.astype(np.float64) ** 2).sum()
    # the forward kernel and the Laplacian kernel agree on psi
    fw = _live.forward(spec, w, torch.from_numpy(x).to(cuda), want=("psi",))["psi"].cpu().numpy()
    assert np.abs(fw - psi).max() <= 2e-5 * np.abs(psi).max()


@pytest.mark.parametrize("D,coord,N", [(2, "mean", 1000), (3, "mean", 333), (4, "mean", 777), (4, "first", 100), (3, "first", 64)])
def test_local_energy_random_params(cuda, D, coord, N):
    from waveflow_b200 import _live
    m = fx.waveflow_model(D, coord=coord)
    params = fx.random_params(np.random.default_rng(10 + D), m, scale=2.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    prot = np.zeros((D, 1))
    x = np.sort(np.random.default_rng(2).uniform(-10, 10, (N, D)), -1).astype(np.float32)
    out = _live.local_energy(spec, w, torch.from_numpy(x).to(cuda), prot, want=("psi", "hpsi", "eloc", "grad", "lap"))
    ref = olap.local_energy_bundle(m, params, x.astype(np.float64), prot)
    ref32 = fast_cpu.FastLocalEnergy(m.cast(np.float32), params, prot, dtype=torch.float32)(x)
    _check_energy(out, ref, live.psi(m.cast(np.float32), fx.cast_params(params, np.float32), x), ref32=ref32)


def test_local_energy_shard_equality(cuda):
    """T5 (emulated ranks): evaluating disjoint row blocks separately gives bit-identical per-walker results and the
    same estimator sums as one call over all walkers."""
    from waveflow_b200 import _live
    m = fx.waveflow_model(4)
    params = fx.random_params(np.random.default_rng(1), m)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    prot = np.zeros((4, 1))
    x = torch.from_numpy(np.sort(np.random.default_rng(3).uniform(-10, 10, (4096, 4)), -1).astype(np.float32)).to(cuda)
    full_s = torch.zeros(4, dtype=torch.float64, device=cuda)
    full = _live.local_energy(spec, w, x, prot, want=("eloc", "psi"), sums=full_s)
    part_s = torch.zeros(4, dtype=torch.float64, device=cuda)
    parts = [_live.local_energy(spec, w, x[r * 512:(r + 1) * 512].contiguous(), prot, want=("eloc", "psi"), sums=part_s) for r in range(8)]
    assert torch.equal(torch.cat([p["eloc"] for p in parts]), full["eloc"])
    assert torch.equal(torch.cat([p["psi"] for p in parts]), full["psi"])
    assert torch.allclose(part_s, full_s, rtol=1e-12)
