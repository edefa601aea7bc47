"""T2: rational-quadratic spline operator parity (GPU through the C ABI vs the numpy restatement)."""
import numpy as np
import pytest
import torch

from oracle import rqs as orqs
from tests.util import assert_fp32_grade, relerr

pytestmark = pytest.mark.gpu


def _run(cuda, x, uw, uh, ud, inverse, B, exact_bins=False):
    from waveflow_b200.flows.neural_splines import unconstrained_RQS
    t = lambda a: torch.from_numpy(a).to(cuda)
    out, lad, bins = unconstrained_RQS(t(x), t(uw), t(uh), t(ud), inverse=inverse, tail_bound=B, return_bin_idx=True,
                                       exact_bins=exact_bins)
    return out.cpu().numpy(), lad.cpu().numpy(), bins.cpu().numpy()


@pytest.mark.parametrize("K", [32, 64])
def test_rqs_exact_bins_bit_identical_to_float32_reference_arithmetic(cuda, K):
    """north_star: bit-exact spline bin indices.  WF_RQS_EXACT_BINS against the float32 restatement of
    neural_splines.py:11-13,98-125 on 2^20 random-parameter elements, forward and inverse, EVERY element (no mask around the
    knots), plus inputs placed exactly on / one ulp beside the oracle's float32 knots."""
    rng = np.random.default_rng(100 + K)
    N, B = 1 << 20, 3.0
    uw, uh = rng.standard_normal((2, N, K)).astype(np.float32)
    ud = rng.standard_normal((N, K - 1)).astype(np.float32)
    for inverse in (False, True):
        x = rng.uniform(-B, B, N).astype(np.float32)
        # adversarial inputs: on the float32 knots of their own row and one ulp to either side
        cw, _ = orqs._knots(uh[: 3 * 4096] if inverse else uw[: 3 * 4096], -B, B, orqs.MIN_BIN_WIDTH)
        pick = rng.integers(0, K + 1, 3 * 4096)
        kn = cw[np.arange(3 * 4096), pick]
        x[:4096] = kn[:4096]
        x[4096:8192] = np.nextafter(kn[4096:8192], np.float32(-10))
        x[8192:12288] = np.nextafter(kn[8192:12288], np.float32(10))
        x = np.clip(x, -B, B)
        out, lad, bins = _run(cuda, x, uw, uh, ud, inverse, B, exact_bins=True)
        o32, l32, b32 = orqs.unconstrained_rqs(x, uw, uh, ud, inverse, B, return_bin=True)
        assert np.array_equal(bins, b32), (K, inverse, int((bins != b32).sum()))
        # same bins and the same float32 knots: outputs agree to float32 rounding of the final rational map
        assert np.abs(out - o32).max() <= 4e-6 * B
        # the default (ex2.approx) path may only differ for inputs within float32 rounding of a knot
        _, _, fast = _run(cuda, x, uw, uh, ud, inverse, B)
        diff = fast != b32
        assert diff[12288:].mean() < 1e-4 and np.all(np.abs(fast[diff] - b32[diff]) == 1)     # [12288:] = the random inputs
        if diff.any():
            knots = orqs._knots((uh if inverse else uw)[diff], -B, B, orqs.MIN_BIN_WIDTH)[0]
            assert np.min(np.abs(knots - x[diff][:, None]), axis=-1).max() < 2e-6


@pytest.mark.parametrize("K", [5, 8, 32, 64])
def test_rqs_forward_inverse_parity(cuda, K):
    rng = np.random.default_rng(K)
    N, B = 50000, 3.0
    x = rng.uniform(-4, 4, N).astype(np.float32)
    x[:6] = [-3.0, 3.0, 0.0, -3.0000002, 3.0000002, 2.9999998]
    uw, uh = rng.standard_normal((2, N, K)).astype(np.float32)
    ud = rng.standard_normal((N, K - 1)).astype(np.float32)
    d64 = lambda a: a.astype(np.float64)
    for inverse in (False, True):
        out, lad, bins = _run(cuda, x, uw, uh, ud, inverse, B)
        ro, rl, rb = orqs.unconstrained_rqs(d64(x), d64(uw), d64(uh), d64(ud), inverse, B, return_bin=True)
        # bin indices: bit-exact away from float32 knot rounding (|x - knot| > 1e-5)
        cw, _ = orqs._knots(d64(uh if inverse else uw), -B, B, orqs.MIN_BIN_WIDTH)
        near = np.min(np.abs(cw - d64(x)[:, None]), axis=-1) < 1e-5
        inside = (x >= -B) & (x <= B)
        assert np.array_equal(bins[~near | ~inside], rb[~near | ~inside])
        assert near.mean() < 1e-3
        assert np.all(bins[~inside] == -1) and np.all(out[~inside] == x[~inside]) and np.all(lad[~inside] == 0)
        o32, l32, b32 = orqs.unconstrained_rqs(x, uw, uh, ud, inverse, B, return_bin=True)
        ok = ~near & (b32 == rb)
        assert_fp32_grade(out[ok], ro[ok], o32[ok], 1e-5, B, "outputs")
        # the extreme tail (one element in 50 000, a bin of width ~1e-3 hit next to its edge) is an order statistic: 8x head-room
        assert_fp32_grade(lad[ok], rl[ok], l32[ok], 1e-5, 1.0, "logabsdet", max_slack=8.0)


def test_rqs_bins_bit_exact_on_exact_knots(cuda):
    """Uniform bins (all unnormalised parameters equal): the float32 knot positions are reproduced operation by operation
    (sequential cumsum, neural_splines.py:98-107), so the located bin must equal the float32 oracle's everywhere."""
    rng = np.random.default_rng(0)
    N, K, B = 200000, 32, 3.0
    uw = np.zeros((N, K), dtype=np.float32); uh = np.zeros((N, K), dtype=np.float32)
    ud = rng.standard_normal((N, K - 1)).astype(np.float32)
    cw, _ = orqs._knots(uw[:1], -B, B, orqs.MIN_BIN_WIDTH)
    x = rng.uniform(-B, B, N).astype(np.float32)
    kn = cw[0]
    x[:K + 1] = kn                                                # exactly on the knots
    x[K + 1:2 * K + 2] = np.nextafter(kn, np.float32(-10))          # one ulp below
    x[2 * K + 2:3 * K + 3] = np.nextafter(kn, np.float32(10))       # one ulp above
    x = np.clip(x, -B, B)
    for inverse in (False, True):
        _, _, bins = _run(cuda, x, uw, uh, ud, inverse, B)
        _, _, rb = orqs.unconstrained_rqs(x, uw, uh, ud, inverse, B, return_bin=True)
        assert np.array_equal(bins, rb)


def test_rqs_round_trip_and_monotone(cuda):
    rng = np.random.default_rng(1)
    N, K, B = 1 << 18, 32, 3.0
    x = np.sort(rng.uniform(-B, B, N).astype(np.float32))
    uw = np.tile(rng.standard_normal((1, K)).astype(np.float32), (N, 1))
    uh = np.tile(rng.standard_normal((1, K)).astype(np.float32), (N, 1))
    ud = np.tile(rng.standard_normal((1, K - 1)).astype(np.float32), (N, 1))
    y, l1, b1 = _run(cuda, x, uw, uh, ud, False, B)
    x2, l2, b2 = _run(cuda, y, uw, uh, ud, True, B)
    assert np.all(np.diff(y) >= -1e-6)                              # monotone map
    assert np.abs(x2 - x).max() < 2e-4 and np.median(np.abs(x2 - x)) < 1e-6
    assert np.abs(l1 + l2).max() < 1e-3
    assert np.mean(b1 == b2) > 0.999


def test_rqs_argument_errors(cuda):
    from waveflow_b200 import _ffi
    from waveflow_b200.flows.neural_splines import unconstrained_RQS
    z = torch.zeros(4, device=cuda)
    with pytest.raises(_ffi.WaveflowB200Error):
        unconstrained_RQS(z, torch.zeros(4, 128, device=cuda), torch.zeros(4, 128, device=cuda), torch.zeros(4, 127, device=cuda))
    out, lad = unconstrained_RQS(torch.zeros(0, device=cuda), torch.zeros(0, 8, device=cuda), torch.zeros(0, 8, device=cuda),
                                 torch.zeros(0, 7, device=cuda))
    assert out.shape == (0,)


@pytest.mark.parametrize("tag", ["op_k8", "op_k32", "op_k32_mild", "op_k5"])
def test_rqs_against_vectors_from_the_reference_source(cuda, tag):
    """wf_rqs_apply against tests/golden/ref_rqs_vectors.npz -- inputs, bin indices and outputs produced by the reference's own
    neural_splines.py executed on a numpy stand-in for its jax imports (tests/golden/make_rqs_golden.py).  Exact-bin mode: every
    bin index the reference computed, forward and inverse; values float32-grade with the reference vector as the yardstick."""
    from pathlib import Path
    G = np.load(Path(__file__).resolve().parent / "golden" / "ref_rqs_vectors.npz")
    uw, uh, ud, B = G[tag + "_uw"], G[tag + "_uh"], G[tag + "_ud"], float(G[tag + "_B"])
    d64 = lambda a: a.astype(np.float64)
    for inverse, xin, ins, bk, yk, lk in ((False, G[tag + "_x"], G[tag + "_inside"], "_bins_fwd", "_y", "_ld"),
                                          (True, G[tag + "_y"], G[tag + "_inside_inv"], "_bins_inv", "_xi", "_ldi")):
        out, lad, bins = _run(cuda, xin, uw, uh, ud, inverse, B, exact_bins=True)
        assert np.array_equal(bins[ins], G[tag + bk]), (tag, inverse, int((bins[ins] != G[tag + bk]).sum()))
        assert np.all(bins[~ins] == -1) and np.array_equal(out[~ins], xin[~ins]) and np.all(lad[~ins] == 0)
        o64, l64 = orqs.unconstrained_rqs(d64(xin), d64(uw), d64(uh), d64(ud), inverse, B)
        assert_fp32_grade(out[ins], o64[ins], G[tag + yk][ins], 1e-5, B, f"rqs vs reference source {tag} inv={inverse} outputs")
        assert_fp32_grade(lad[ins], l64[ins], G[tag + lk][ins], 1e-5, 1.0, f"rqs vs reference source {tag} inv={inverse} logabsdet",
                          max_slack=8.0)
        _, _, fast = _run(cuda, xin, uw, uh, ud, inverse, B)
        assert np.mean(fast[ins] != G[tag + bk]) < 2e-3          # default (ex2.approx) path: may differ only next to a knot
