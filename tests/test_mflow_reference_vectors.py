"""BASELINE configs[0] (2-D MFlow density): oracle/live.py::log_pdf and the product's table generator against vectors produced by
the REFERENCE'S OWN SOURCE FILES (flows.MFlow / IMADE / masked transform / isplines_jax / msplines_jax and the splines_np.M table
generator, executed on the numpy stand-in for jax: tests/golden/make_mflow_golden.py).  CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import fixtures as fx
from oracle import live

G = np.load(Path(__file__).resolve().parent / "golden" / "ref_mflow_vectors.npz")
CFG = json.loads(str(G["cfg"]))


def mflow_params(dtype):
    def rec(t):
        return tuple(rec(u) for u in t) if isinstance(t, list) else G[f"param{t:03d}"].astype(dtype)
    return rec(json.loads(str(G["treedef"])))


def mflow_model(dtype):
    return fx.mflow_model(D=CFG["D"], i_degree=CFG["k_i"], i_knots=CFG["n_i"], n_layers=CFG["L"], reg=CFG["reg"], tol=1e-6,
                          p_degree=CFG["k_p"], p_knots=CFG["n_p"], dtype=dtype)


def test_m_tables_bit_identical_to_the_reference_generator():
    """The M-spline basis tables are not shipped: the reference's own splines_np.M generated them; the product generator
    (waveflow_b200/splines/tablegen.py) reproduces every float64 bit, all four derivative orders."""
    ref = np.stack([G[f"Mtab_degree_{CFG['k_p']}_niknots_{CFG['n_p'] + CFG['k_p'] - 2}_nmp_2000_nd_{n}"] for n in range(4)])
    assert ref.dtype == np.float64 and np.array_equal(ref, fx.tables_M(CFG["k_p"], CFG["n_p"]))


@pytest.mark.parametrize("mode,dtype", [("f32", np.float32), ("f64", np.float64)])
def test_mflow_log_pdf_bit_identical_to_the_reference_source(mode, dtype):
    """MFlow.log_pdf and the flow output u (distributions.py:131-157 over made.py:66-81) in JAX's default float32 and in float64."""
    lp, u = live.log_pdf(mflow_model(dtype), mflow_params(dtype), G[mode + "_x"], return_sample=True)
    assert lp.dtype == dtype
    assert np.array_equal(u, G[mode + "_u"]) and np.array_equal(lp, G[mode + "_logpdf"])


@pytest.mark.parametrize("mode,dtype", [("f32", np.float32), ("f64", np.float64)])
def test_inverse_flow_and_box_transform_bit_identical_to_the_reference_source(mode, dtype):
    """Serial(IMADE, Reverse).inverse_fun (vmapped bisection; IMADE's inverse conditions on its input -- SURVEY quirk Q1 -- so the
    reference's own round trip is 2e-5, not its 1e-6 tolerance) and BoxTransformLayer direct / reverse for both coordinate types
    (reverse_fun_mean is quirk Q2)."""
    m = mflow_model(dtype)
    back = live.flow_inverse(m, mflow_params(dtype)[0], G[mode + "_u"])
    assert np.array_equal(back, G[mode + "_x_back"])
    assert 1e-6 < np.abs(G[mode + "_x_back"] - G[mode + "_x"]).max() < 1e-4
    for coord in ("mean", "first"):
        L = float(G[f"{mode}_box_{coord}_L"])
        u, ld = live.box_direct(G[f"{mode}_box_{coord}_x"], L, coord)
        assert np.array_equal(u, G[f"{mode}_box_{coord}_u"]) and np.array_equal(ld, G[f"{mode}_box_{coord}_ld"])
        assert np.array_equal(live.box_inverse(G[f"{mode}_box_{coord}_u"], L, coord), G[f"{mode}_box_{coord}_back"])
