"""Tensor-core live path (csrc/live_tc.cuh: conditioner layers 2 and 3 on tcgen05, 3xTF32, activations in tensor memory)
against the float64 oracle AND against the CUDA-core kernel, through the same C-ABI entry points (weight_layout = WF_WEIGHTS_TC)."""
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
    w = _live.pack_params(spec, params[0], params[1], device)
    assert w.wf_tc is not None
    return w


def test_tc_weight_image_layout(cuda):
    """wf_live_pack_tc: hi + lo reproduce the masked weights to 2^-21 relative, rows / k-blocks / 16-byte chunk swizzle as documented
    in include/waveflow_b200.h, first layer and biases copied."""
    from waveflow_b200 import _ffi, _live
    m = fx.waveflow_model(4)
    params = fx.random_params(np.random.default_rng(0), m)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    D, H, NP = 4, 64, 32
    ns, nt = int(_ffi.lib.wf_live_net_floats(D)), int(_ffi.lib.wf_live_net_floats_tc(D))
    ws, wt = w.cpu().numpy().reshape(4, ns), w.wf_tc.cpu().numpy().reshape(4, nt)
    for net in range(4):
        s, t = ws[net], wt[net]
        W1, b1 = s[:D * H], s[D * H:D * H + H]
        W2 = s[D * H + H:D * H + H + H * H].reshape(H, H)
        o = D * H + H + H * H
        b2, W3, b3 = s[o:o + H], s[o + H:o + H + H * D * NP].reshape(H, D * NP), s[o + H + H * D * NP:]

        def unswizzle(plane, rows):
            p = plane.reshape(2, rows, 8, 4)
            out = np.empty((rows, 64), dtype=np.float32)
            for kb in range(2):
                for r in range(rows):
                    for pos in range(8):
                        c = pos ^ (r & 7)
                        out[r, kb * 32 + c * 4: kb * 32 + c * 4 + 4] = p[kb, r, pos]
            return out                                                # [n][k]
        w2 = unswizzle(t[:H * H], H) + unswizzle(t[H * H:2 * H * H], H)
        o3 = 2 * H * H
        w3 = unswizzle(t[o3:o3 + D * NP * H], D * NP) + unswizzle(t[o3 + D * NP * H:o3 + 2 * D * NP * H], D * NP)
        assert np.abs(w2 - W2.T).max() <= 2.0 ** -20 * np.abs(W2).max()
        assert np.abs(w3 - W3.T).max() <= 2.0 ** -20 * np.abs(W3).max()
        hi = unswizzle(t[:H * H], H)
        assert np.all((hi.view(np.uint32) & 0x1FFF) == 0)              # TF32-exact plane
        small = t[o3 + 2 * D * NP * H:]
        assert np.array_equal(small, np.concatenate([W1, b1, b2, b3]))


@pytest.mark.parametrize("D,coord,N", [(2, "mean", 1000), (3, "mean", 333), (4, "mean", 777), (4, "first", 2100), (2, "first", 64)])
def test_tc_local_energy_vs_oracle_and_simt(cuda, D, coord, N):
    from waveflow_b200 import _live
    m = fx.waveflow_model(D, coord=coord)
    params = fx.random_params(np.random.default_rng(10 + D), m, scale=2.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    prot = np.zeros((D, 1))
    x = np.sort(np.random.default_rng(2).uniform(-10, 10, (N, D)), -1).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    want = ("psi", "hpsi", "eloc", "grad", "lap")
    s_tc = torch.zeros(4, dtype=torch.float64, device=cuda)
    tc = _live.local_energy(spec, w, xt, prot, want=want, sums=s_tc, mode="tc")
    s_si = torch.zeros(4, dtype=torch.float64, device=cuda)
    si = _live.local_energy(spec, w, xt, prot, want=want, sums=s_si, mode="simt")
    torch.cuda.synchronize()
    ref = olap.local_energy_bundle(m, params, x.astype(np.float64), prot)
    psi32 = live.psi(m.cast(np.float32), fx.cast_params(params, np.float32), x)
    ref32 = fast_cpu.FastLocalEnergy(m.cast(np.float32), params, prot, dtype=torch.float32)(x)
    g = {k: v.cpu().numpy() for k, v in tc.items()}
    assert_fp32_grade(g["psi"], ref["psi"], psi32, 1e-5, name=f"tc psi D={D}")
    for k in ("grad", "lap", "hpsi"):
        assert relerr(g[k], ref[k]) < 1e-4, (k, relerr(g[k], ref[k]), relerr(si[k].cpu().numpy(), ref[k]))
    record_flat(f"tc E_loc D={D} (pointwise relative)", g["eloc"], ref["eloc"], 1e-4, ref32["eloc"])
    # the two kernels evaluate the same formulas; they differ by float32 rounding order and the 3xTF32 products only
    for k in ("psi", "hpsi", "lap"):
        a, b = g[k].astype(np.float64), si[k].cpu().numpy().astype(np.float64)
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max() + 2 * np.abs(b - ref[k]).max(), k
    st, ss = s_tc.cpu().numpy(), s_si.cpu().numpy()
    assert st[2] == N == ss[2]
    assert abs(st[3] - ss[3]) <= 1e-5 * ss[3]


@pytest.mark.parametrize("D", [2, 4])
def test_tc_forward_logpdf_and_u(cuda, D):
    """Forward-only variant (rows = walkers): Waveflow psi / log_pdf / flow output u, and the MFlow log_pdf (M prior)."""
    from waveflow_b200 import _live
    m = fx.waveflow_model(D)
    params = fx.random_params(np.random.default_rng(20 + D), m, scale=2.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    x = np.sort(np.random.default_rng(3).uniform(-10, 10, (5000, D)), -1).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    tc = _live.forward(spec, w, xt, want=("u", "logdet", "logpdf", "psi"), mode="tc")
    si = _live.forward(spec, w, xt, want=("u", "logdet", "logpdf", "psi"), mode="simt")
    psi64 = live.psi(m, params, x.astype(np.float64))
    psi32 = live.psi(m.cast(np.float32), fx.cast_params(params, np.float32), x)
    assert_fp32_grade(tc["psi"].cpu().numpy(), psi64, psi32, 1e-5, name=f"tc forward psi D={D}")
    u64, ld64 = live.flow_direct(m, params[0], x.astype(np.float64))
    u32, ld32 = live.flow_direct(m.cast(np.float32), fx.cast_params(params, np.float32)[0], x)
    assert np.abs(tc["u"].cpu().numpy() - u64).max() < 2e-5
    assert_fp32_grade(tc["logdet"].cpu().numpy(), ld64, ld32, 1e-5, 1.0, f"tc logdet D={D}")
    lp64 = live.log_pdf(m, params, x.astype(np.float64))
    lp32 = live.log_pdf(m.cast(np.float32), fx.cast_params(params, np.float32), x)
    assert_fp32_grade(tc["logpdf"].cpu().numpy(), lp64, lp32, 1e-5, 1.0, f"tc logpdf D={D}")
    # the CUDA-core kernel evaluates the same formulas: the two agree wherever the result is well conditioned
    for k in ("u", "logdet"):
        a, b = tc[k].cpu().numpy().astype(np.float64), si[k].cpu().numpy().astype(np.float64)
        assert np.abs(a - b).max() <= 1e-4 * (np.abs(b).max() + 1.0), (k, np.abs(a - b).max())


def test_tc_mflow_logpdf(cuda):
    from waveflow_b200 import _live
    m = fx.mflow_model()
    params = fx.random_params(np.random.default_rng(4), m, scale=2.0)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    x = np.random.default_rng(5).uniform(0.025, 0.975, (4096, 2)).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    tc = _live.forward(spec, w, xt, want=("logpdf", "u"), mode="tc")
    r64 = live.log_pdf(m, params, x.astype(np.float64))
    r32 = live.log_pdf(m.cast(np.float32), fx.cast_params(params, np.float32), x)
    assert_fp32_grade(tc["logpdf"].cpu().numpy(), r64, r32, 1e-5, 1.0, "tc MFlow.log_pdf")


def test_tc_he_checkpoint_psi_kat(cuda):
    """The published He checkpoint -> psi on the published samples, tensor-core path."""
    from waveflow_b200 import _live
    params, gold = fx.load_he_checkpoint()
    m = fx.waveflow_model(2)
    spec = spec_from_live(m)
    w = _pack(spec, params, cuda)
    x = np.sort(np.tile(gold["samples"], (20, 1)), -1).astype(np.float32)
    tc = _live.forward(spec, w, torch.from_numpy(x).to(cuda), want=("psi",), mode="tc")["psi"].cpu().numpy()
    r64 = live.psi(m, params, x.astype(np.float64))
    assert np.abs(tc - r64).max() <= 2e-5 * np.abs(r64).max()


@pytest.mark.parametrize("mode", ["tc", "simt"])
def test_mflow_logpdf_against_vectors_from_the_reference_source(cuda, mode):
    """BASELINE configs[0]: the fused forward kernels (both weight layouts) against MFlow.log_pdf of the reference's own source
    files executed on a numpy stand-in for jax (tests/golden/make_mflow_golden.py): float64 vector as the truth, the reference's
    float32 vector as the yardstick."""
    from waveflow_b200 import _live
    from tests.test_mflow_reference_vectors import G, mflow_model, mflow_params
    m = mflow_model(np.float64)
    spec = spec_from_live(m)
    w = _pack(spec, mflow_params(np.float32), cuda)
    x = G["f32_x"]
    assert np.array_equal(x.astype(np.float64).astype(np.float32), x) and np.abs(G["f64_x"] - x).max() < 1e-7
    out = _live.forward(spec, w, torch.from_numpy(x).to(cuda), want=("logpdf", "u"), mode=mode)
    # the float64 vector was evaluated at the float64 inputs (x rounded to float32 moves log_pdf by < 1e-6): truth at x itself
    r64 = live.log_pdf(m, mflow_params(np.float64), x.astype(np.float64))
    assert_fp32_grade(out["logpdf"].cpu().numpy(), r64, G["f32_logpdf"], 1e-5, 1.0, f"{mode} MFlow.log_pdf vs reference source")
    assert np.abs(out["u"].cpu().numpy() - G["f32_u"]).max() <= 2e-6


@pytest.mark.parametrize("tag", ["d2_mean", "d3_first", "d4_mean_l3"])
def test_local_energy_against_vectors_from_the_reference_source(cuda, tag):
    """wf_local_energy (tensor-core and CUDA-core kernels) against psi and H psi produced by the reference's own model_factory /
    wavefunctions / flows / splines / physics.construct_hamiltonian_function (jax.hessian with the registered custom_jvp rules),
    executed in float64 on the numpy stand-in for jax (tests/golden/make_energy_golden.py)."""
    from waveflow_b200 import _live
    from tests.test_energy_reference_vectors import G, _model
    m, params, D = _model(tag)
    spec = spec_from_live(m)
    w = _pack(spec, fx.cast_params(params, np.float32), cuda)
    x = G[tag + "_x"].astype(np.float32)
    ref = olap.local_energy_bundle(m, params, x.astype(np.float64), np.zeros((D, 1)))      # == the reference vectors at float64 x
    assert np.abs(ref["hpsi"] - G[tag + "_hpsi"]).max() <= 1e-5 * np.abs(G[tag + "_hpsi"]).max()   # float32 rounding of x only
    # yardstick: the reference's arithmetic in float32 (the vectorised CPU port of the same formulas)
    r32 = fast_cpu.FastLocalEnergy(m.cast(np.float32), params, np.zeros((D, 1)), dtype=torch.float32)(x)
    for mode in ("tc", "simt"):
        out = _live.local_energy(spec, w, torch.from_numpy(x).to(cuda), np.zeros((D, 1)), want=("psi", "hpsi"), mode=mode)
        for key, tol in (("psi", 2e-5), ("hpsi", 2e-4)):
            got, gold = out[key].cpu().numpy().astype(np.float64), G[tag + "_" + key]
            scale = np.abs(gold).max()
            e32 = np.abs(np.asarray(r32[key], dtype=np.float64) - gold).max()
            err = np.abs(got - gold).max()
            print(f"{tag} {mode} {key}: max error / max|ref| = {err / scale:.2e} (float32 CPU arithmetic {e32 / scale:.2e})")
            assert err <= max(tol * scale, 4 * e32 + 1e-6 * scale), (mode, key, err / scale, e32 / scale)
