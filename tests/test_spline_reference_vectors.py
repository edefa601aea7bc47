"""oracle/live.py spline operators against vectors produced by the REFERENCE'S OWN SOURCE (splines/isplines_jax.py,
splines/bsplines_jax.py, utils/helpers.binary_search executed on a numpy stand-in for jax: tests/golden/make_spline_golden.py,
on the basis tables the reference ships).  CPU only."""
from pathlib import Path

import numpy as np

from oracle import live

GOLD = Path(__file__).resolve().parent / "golden"
G = np.load(GOLD / "ref_spline_vectors.npz")
T = np.load(GOLD / "ref_tables_deg5_k16.npz")
TAB_I = np.stack([T[f"I_nd{n}"] for n in range(4)]).astype(np.float32)
TAB_B = np.stack([T[f"B_nd{n}"] for n in range(4)]).astype(np.float32)
TAB_OB = np.stack([T[f"OB_nd{n}"] for n in range(4)]).astype(np.float32)


def test_ispline_operators_bit_identical_to_the_reference_source():
    """remove_bias, enforce_boundary_conditions ({0: 0} | {0: 1}), apply_fun_vec, apply_fun_vec_grad (= the file's own custom_jvp
    rule: the derivative table) and reverse_fun_vec (bisection, float32 loop state): every float32 bit."""
    k = int(G["k"])
    assert np.array_equal(live.remove_bias_I(G["raw"], k), G["remove_bias"])
    assert np.array_equal(live.enforce_bc(TAB_I, G["remove_bias"], {0: 0.0}, {0: 1.0}, "I"), G["enforce_bc"])
    c, x = G["enforce_bc"], G["x"]
    assert np.array_equal(live.spline_apply(TAB_I, c, x, 0), G["apply"])
    assert np.array_equal(live.spline_apply(TAB_I, c, x, 1), G["apply_grad"])
    assert np.array_equal(live.binary_search_inverse(TAB_I, c, G["apply"], float(G["tol"])), G["reverse"])
    assert np.abs(G["reverse"] - x).max() <= float(G["tol"])            # the reference's own round trip: within its tolerance


def test_bspline_prior_operators_equal_the_reference_source():
    """BSpline_fun: enforce_boundary_conditions ({0: 0, 2: 0} | {0: 0}, L2-normalised) bit-identical; apply / grad go through a
    float32 matrix product (w @ ob_to_b), identical here with the same BLAS, asserted to float32 rounding."""
    assert np.array_equal(live.enforce_bc(TAB_B, G["B_raw"], {0: 0, 2: 0}, {0: 0}, "B"), G["B_enforce_bc"])
    c = live.bspline_coeffs(G["B_enforce_bc"], T["ob_to_b"].astype(np.float32))
    for nd, key in ((0, "B_apply"), (1, "B_apply_grad")):
        got = live.spline_apply(TAB_OB, c, G["B_x"], nd)
        assert np.abs(got - G[key]).max() <= 2e-6 * np.abs(G[key]).max()
