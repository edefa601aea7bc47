"""oracle/rqs.py against vectors produced by the REFERENCE'S OWN SOURCE (flows/bijections/neural_splines.py executed on a numpy
stand-in for its jax imports: tests/golden/make_rqs_golden.py, tests/golden/jax_numpy_shim/README.md).  CPU only."""
from pathlib import Path

import numpy as np
import pytest

from oracle import rqs as orq

G = np.load(Path(__file__).resolve().parent / "golden" / "ref_rqs_vectors.npz")
OPS = ["op_k8", "op_k32", "op_k32_mild", "op_k5"]


def _case(tag):
    return G[tag + "_x"], G[tag + "_uw"], G[tag + "_uh"], G[tag + "_ud"], float(G[tag + "_B"])


@pytest.mark.parametrize("tag", OPS)
def test_bin_indices_and_knots_equal_the_reference(tag):
    """searchsorted / knot construction (neural_splines.py:11-13, 98-125): every bin index the reference computed, forward and
    inverse, and its float32 knot vectors to the last bit."""
    x, uw, uh, ud, B = _case(tag)
    _, _, bins = orq.unconstrained_rqs(x, uw, uh, ud, inverse=False, tail_bound=B, return_bin=True)
    ins = G[tag + "_inside"]
    assert np.array_equal(ins, (x >= -B) & (x <= B))
    assert np.array_equal(bins[ins], G[tag + "_bins_fwd"]) and np.all(bins[~ins] == -1)
    y = G[tag + "_y"]
    _, _, bins_i = orq.unconstrained_rqs(y, uw, uh, ud, inverse=True, tail_bound=B, return_bin=True)
    ins_i = G[tag + "_inside_inv"]
    assert np.array_equal(bins_i[ins_i], G[tag + "_bins_inv"])
    cw, _ = orq._knots(uw[ins], -B, B, orq.MIN_BIN_WIDTH)
    ref = G[tag + "_knots_fwd"]               # as passed to searchsorted (its +1e-6 on the last knot is applied to a copy)
    # numpy's float32 exp and pairwise sum vs the restatement's correctly rounded exp + sequential sum: more than half of the
    # knots are bit-identical, the rest within a few ulp (ulp(3) = 2.4e-7) -- and no input of the set falls in between
    assert np.array_equal(cw[..., 0], ref[..., 0]) and np.array_equal(cw[..., -1], ref[..., -1])
    assert np.abs(cw - ref).max() <= 2e-6 * B
    assert np.mean(cw == ref) > 0.5


@pytest.mark.parametrize("tag", OPS)
def test_operator_values_are_float32_grade(tag):
    """unconstrained_RQS forward / inverse values and log-dets: the restatement and the reference source differ only by float32
    rounding (exp implementation, summation order) -- identical statistics against the float64 evaluation."""
    x, uw, uh, ud, B = _case(tag)
    d = lambda a: a.astype(np.float64)
    for inverse, xin, yk, lk in ((False, x, "_y", "_ld"), (True, G[tag + "_y"], "_xi", "_ldi")):
        o32, l32 = orq.unconstrained_rqs(xin, uw, uh, ud, inverse=inverse, tail_bound=B)
        o64, l64 = orq.unconstrained_rqs(d(xin), d(uw), d(uh), d(ud), inverse=inverse, tail_bound=B)
        for got, ref, truth, scale in ((o32, G[tag + yk], o64, B), (l32, G[tag + lk], l64, 1.0)):
            e_o, e_r = np.abs(got - truth), np.abs(ref - truth)
            assert np.median(e_o) <= 2 * np.median(e_r) + 1e-7 * scale
            assert np.quantile(e_o, 0.99) <= 2 * np.quantile(e_r, 0.99) + 1e-6 * scale
            assert e_o.max() <= 4 * e_r.max() + 1e-5 * scale
        outside = ~((xin >= -B) & (xin <= B))
        assert np.array_equal(o32[outside], G[tag + yk][outside]) and np.all(l32[outside] == 0)      # identity tails
    if tag.endswith("mild") or tag == "op_k5":
        y32, l32 = orq.unconstrained_rqs(x, uw, uh, ud, tail_bound=B)
        assert np.abs(y32 - G[tag + "_y"]).max() <= 2e-6 * B
        dl = np.abs(l32 - G[tag + "_ld"])
        assert np.median(dl) <= 5e-7 and np.quantile(dl, 0.99) <= 5e-5 and dl.max() <= 5e-4


@pytest.mark.parametrize("tag", ["cpl_d2", "cpl_d8"])
def test_coupling_layer_equals_the_reference(tag):
    """NeuralSplineCoupling direct_fun / inverse_fun (neural_splines.py:243-296) with the weights the reference layer created."""
    f = lambda w: [(G[f"{tag}_{w}_W{i}"], G[f"{tag}_{w}_b{i}"]) for i in range(3)]
    K, B = int(G[tag + "_K"]), float(G[tag + "_B"])
    y, ld = orq.coupling_direct(f("f1"), f("f2"), G[tag + "_x"], K, B)
    assert np.abs(y - G[tag + "_y"]).max() <= 4e-6 * B and np.abs(ld - G[tag + "_ld"]).max() <= 1e-4
    xi, ldi = orq.coupling_inverse(f("f1"), f("f2"), G[tag + "_y"], K, B)
    assert np.abs(xi - G[tag + "_xi"]).max() <= 4e-6 * B and np.abs(ldi - G[tag + "_ldi"]).max() <= 1e-4
    assert np.abs(G[tag + "_xi"] - G[tag + "_x"]).max() <= 1e-5 * B          # the reference's own round trip
