"""NeuralSplineCoupling stack (neural_splines.py:244-296) fused in one kernel vs the numpy restatement."""
import numpy as np
import pytest
import torch

from oracle import rqs as orqs
from tests.util import assert_fp32_grade

pytestmark = pytest.mark.gpu


def _layers(rng, D, K, hidden, L):
    out = (3 * K - 1) * D // 2
    return [(orqs.random_fcnn(rng, D // 2, hidden, out), orqs.random_fcnn(rng, D // 2, hidden, out)) for _ in range(L)]


def _to_stax(f, device):
    """[(W,b)]*3 -> stax.serial params [(W,b), (), (W,b), (), (W,b)] on the device."""
    t = lambda a: torch.from_numpy(a).to(device)
    (W1, b1), (W2, b2), (W3, b3) = f
    return [(t(W1), t(b1)), (), (t(W2), t(b2)), (), (t(W3), t(b3))]


@pytest.mark.parametrize("D,K,hidden,L", [(2, 32, 8, 8), (8, 32, 64, 8), (4, 5, 8, 1), (8, 32, 8, 2)])
def test_coupling_flow_forward_inverse(cuda, D, K, hidden, L):
    from waveflow_b200.flows.neural_splines import coupling_flow
    rng = np.random.default_rng(D * 100 + K)
    B = 3.0
    layers = _layers(rng, D, K, hidden, L)
    N = 20000
    x = rng.uniform(-3.5, 3.5, (N, D)).astype(np.float32)
    tl = [(_to_stax(f1, cuda), _to_stax(f2, cuda)) for f1, f2 in layers]
    y, ld = coupling_flow(tl, torch.from_numpy(x).to(cuda), K, B, hidden)
    l64 = [tuple([(W.astype(np.float64), b.astype(np.float64)) for W, b in f] for f in pair) for pair in layers]
    ry, rld = orqs.coupling_flow_direct(l64, x.astype(np.float64), K, B)
    r32y, r32ld = orqs.coupling_flow_direct(layers, x, K, B)
    # a bin decision taken at rounding level anywhere in the 2L half-updates changes the sample's path: compare where the
    # float32 restatement itself stayed on the float64 path
    ok = (np.abs(r32y - ry).max(-1) < 1e-3)
    assert ok.mean() > 0.99
    assert_fp32_grade(y.cpu().numpy()[ok], ry[ok], r32y[ok], 1e-5, B, "outputs", max_slack=8.0)
    assert_fp32_grade(ld.cpu().numpy()[ok], rld[ok], r32ld[ok], 1e-5, 1.0, "log_det", max_slack=8.0)
    # inverse(direct(x)) == x (tests/test_bijections.py:12-21, atol 1e-3) and log-dets cancel
    xr, ldi = coupling_flow(tl, y, K, B, hidden, inverse=True)
    assert np.allclose(xr.cpu().numpy(), x, atol=1e-3)
    assert np.median(np.abs((ld + ldi).cpu().numpy())) < 1e-4
    # identity tails: a sample entirely outside [-B, B] is untouched
    far = torch.full((4, D), 5.0, device=cuda)
    yf, lf = coupling_flow(tl, far, K, B, hidden)
    assert torch.equal(yf, far) and float(lf.abs().max()) == 0.0


def test_neural_spline_coupling_layer_protocol(cuda):
    from waveflow_b200 import flows
    from waveflow_b200.flows.neural_splines import NeuralSplineCoupling
    params, direct, inverse = NeuralSplineCoupling()(0, 4)             # defaults K=5, B=3, hidden_dim=8 (:244)
    params = tuple([tuple(t.to(cuda) for t in l) if len(l) else () for l in f] for f in params)
    x = (torch.rand(20, 4, device=cuda) * 20 - 10)                      # test_bijections.py:12-13 input range
    y, ld = direct(params, x)
    assert tuple(y.shape) == (20, 4) and tuple(ld.shape) == (20,)
    xr, _ = inverse(params, y)
    assert torch.allclose(xr, x, atol=1e-3)


# ------------------------------------------------------------------------------------------- tensor-core path (config 5)
def test_tc_dense_layer_3xtf32(cuda):
    """wf_tc_dense: tcgen05 GEMM with the hi/lo TF32 split against a float64 matmul (float32-grade accuracy)."""
    from waveflow_b200._ffi import check, lib, ptr, stream_ptr
    g = torch.Generator(device=cuda); g.manual_seed(0)
    for M, K, N in [(128, 32, 128), (300, 512, 512), (1000, 512, 576), (77, 64, 256)]:
        A = torch.randn(M, K, device=cuda, generator=g); W = torch.randn(N, K, device=cuda, generator=g) / K ** 0.5
        b = torch.randn(N, device=cuda, generator=g)
        planes = []
        for t in (A, W):
            hi, lo = torch.empty_like(t), torch.empty_like(t)
            check(lib.wf_tf32_split(ptr(t), t.numel(), ptr(hi), ptr(lo), stream_ptr()))
            assert float((t - hi - lo).abs().max()) <= 3e-7 * float(t.abs().max())
            assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))   # TF32-exact
            planes += [hi, lo]
        out = torch.empty(M, N, device=cuda)
        check(lib.wf_tc_dense(ptr(planes[0]), ptr(planes[1]), M, K, ptr(planes[2]), ptr(planes[3]), N, ptr(b), 0, ptr(out), None, stream_ptr()))
        ref = A.double() @ W.double().T + b.double()
        e32 = float(((A @ W.T + b).double() - ref).abs().max())
        err = float((out.double() - ref).abs().max())
        assert err <= 1e-5 * float(ref.abs().max()) and err <= 8 * e32 + 1e-6, (M, K, N, err, e32)
        oh, ol = torch.empty(M, N, device=cuda), torch.empty(M, N, device=cuda)
        check(lib.wf_tc_dense(ptr(planes[0]), ptr(planes[1]), M, K, ptr(planes[2]), ptr(planes[3]), N, ptr(b), 1, ptr(oh), ptr(ol), stream_ptr()))
        assert float(((oh + ol).double() - torch.tanh(ref)).abs().max()) < 5e-5


def test_tc_coupling_flow_config5(cuda):
    """D = 64, K = 64, hidden 512 (BASELINE config 5) on the tensor cores vs the numpy restatement."""
    from waveflow_b200.flows.neural_splines import coupling_flow_tc, pack_fcnn_tc
    D, K, H, B, L = 64, 64, 512, 3.0, 2
    rng = np.random.default_rng(0)
    layers = _layers(rng, D, K, H, L)
    w = torch.cat([pack_fcnn_tc(_to_stax(f, cuda), cuda) for pair in layers for f in pair]).contiguous()
    N = 1100                                                       # ragged last row tile
    x = rng.uniform(-3.3, 3.3, (N, D)).astype(np.float32)
    y, ld = coupling_flow_tc(w, L, torch.from_numpy(x).to(cuda), B)
    l64 = [tuple([(W.astype(np.float64), b.astype(np.float64)) for W, b in f] for f in pair) for pair in layers]
    ry, rld = orqs.coupling_flow_direct(l64, x.astype(np.float64), K, B)
    r32y, r32ld = orqs.coupling_flow_direct(layers, x, K, B)
    ok = np.abs(r32y - ry).max(-1) < 1e-3
    assert ok.mean() > 0.99
    assert_fp32_grade(y.cpu().numpy()[ok], ry[ok], r32y[ok], 1e-5, B, "outputs", max_slack=8.0)
    assert_fp32_grade(ld.cpu().numpy()[ok], rld[ok], r32ld[ok], 1e-5, 1.0, "log_det", max_slack=8.0)
    xr, ldi = coupling_flow_tc(w, L, y, B, inverse=True)
    assert np.allclose(xr.cpu().numpy(), x, atol=1e-3)
    assert np.median(np.abs((ld + ldi).cpu().numpy())) < 1e-3
    # chunked execution (workspace smaller than the batch) gives identical results
    y2, ld2 = coupling_flow_tc(w, L, torch.from_numpy(x).to(cuda), B, chunk_rows=256)
    assert torch.equal(y, y2) and torch.equal(ld, ld2)


@pytest.mark.parametrize("tag", ["cpl_d2", "cpl_d8"])
def test_coupling_layer_against_vectors_from_the_reference_source(cuda, tag):
    """The fused coupling kernel against NeuralSplineCoupling.direct_fun / inverse_fun of the reference's own neural_splines.py
    (executed on a numpy stand-in for its jax imports, tests/golden/make_rqs_golden.py), with the weights that layer created."""
    from pathlib import Path
    from waveflow_b200.flows.neural_splines import coupling_flow
    G = np.load(Path(__file__).resolve().parent / "golden" / "ref_rqs_vectors.npz")
    f = lambda w: [(G[f"{tag}_{w}_W{i}"], G[f"{tag}_{w}_b{i}"]) for i in range(3)]
    K, B, hidden = int(G[tag + "_K"]), float(G[tag + "_B"]), int(G[tag + "_hidden"])
    layers = [(f("f1"), f("f2"))]
    tl = [(_to_stax(f("f1"), cuda), _to_stax(f("f2"), cuda))]
    l64 = [tuple([(W.astype(np.float64), b.astype(np.float64)) for W, b in net] for net in layers[0])]
    x = G[tag + "_x"]
    y, ld = coupling_flow(tl, torch.from_numpy(x).to(cuda), K, B, hidden)
    ry, rld = orqs.coupling_flow_direct(l64, x.astype(np.float64), K, B)
    assert_fp32_grade(y.cpu().numpy(), ry, G[tag + "_y"], 1e-5, B, f"coupling vs reference source {tag} outputs", max_slack=8.0)
    assert_fp32_grade(ld.cpu().numpy(), rld, G[tag + "_ld"], 1e-5, 1.0, f"coupling vs reference source {tag} log_det", max_slack=8.0)
    assert np.abs(y.cpu().numpy() - G[tag + "_y"]).max() <= 2e-5 * B
    yin = G[tag + "_y"]
    xi, ldi = coupling_flow(tl, torch.from_numpy(yin).to(cuda), K, B, hidden, inverse=True)
    rx, rldi = orqs.coupling_flow_inverse(l64, yin.astype(np.float64), K, B)
    assert_fp32_grade(xi.cpu().numpy(), rx, G[tag + "_xi"], 1e-5, B, f"coupling vs reference source {tag} inverse", max_slack=8.0)
    assert_fp32_grade(ldi.cpu().numpy(), rldi, G[tag + "_ldi"], 1e-5, 1.0, f"coupling vs reference source {tag} inverse log_det",
                      max_slack=8.0)
