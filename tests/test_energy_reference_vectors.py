"""The local-energy and gradient oracles against vectors produced by the REFERENCE'S OWN SOURCE FILES: model_factory ->
wavefunctions.Waveflow -> flows (BoxTransformLayer, IMADE, Reverse) -> isplines_jax / bsplines_jax with the custom_jvp rules they
register, utils/physics.construct_hamiltonian_function (jax.hessian) and vqmc.loss_fn_efficient with its custom_jvp estimator,
executed in float64 on the numpy stand-in for jax (tests/golden/make_energy_golden.py, tests/golden/jax_numpy_shim/README.md).
CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import fixtures as fx
from oracle import grad as ograd
from oracle import laplacian as olap
from oracle import live

G = np.load(Path(__file__).resolve().parent / "golden" / "ref_energy_vectors.npz")


def _tree(tag, get):
    def rec(t):
        return tuple(rec(u) for u in t) if isinstance(t, list) else get(t)
    return rec(json.loads(str(G[tag + "_treedef"])))


def _model(tag):
    D = int(G[tag + "_D"])
    coord = "mean" if int(G[tag + "_coord_mean"]) else "first"
    m = fx.waveflow_model(D, degree=5, n_knots=16, n_layers=int(G[tag + "_layers"]), box=float(G[tag + "_box"]), reg=0.0,
                          tol=1e-6, coord=coord)
    return m, _tree(tag, lambda i: G[f"{tag}_param{i:03d}"]), D


def _leaves(t, out):
    if isinstance(t, (tuple, list)):
        for u in t:
            _leaves(u, out)
    else:
        out.append(np.asarray(t, dtype=np.float64))
    return out


@pytest.mark.parametrize("tag", ["d2_mean", "d3_first", "d4_mean_l3"])
def test_psi_logpdf_and_local_energy_equal_the_reference_source(tag):
    """psi, log_pdf and H psi = -1/2 trace(hessian(psi)) + V psi: both restatements of oracle/laplacian.py (forward-Laplacian
    bundles, torch double autograd) against the reference's own code -- float64, 1e-12."""
    m, params, D = _model(tag)
    x, protons = G[tag + "_x"], np.zeros((D, 1))
    b = olap.local_energy_bundle(m, params, x, protons)
    a = olap.local_energy_autograd(m, params, x, protons)
    assert np.abs(b["psi"] - G[tag + "_psi"]).max() <= 1e-13 * np.abs(G[tag + "_psi"]).max()
    s = np.abs(G[tag + "_hpsi"]).max()
    assert np.abs(b["hpsi"] - G[tag + "_hpsi"]).max() <= 1e-12 * s
    assert np.abs(a["hpsi"] - G[tag + "_hpsi"]).max() <= 1e-12 * s
    assert np.abs(live.log_pdf(m, params, x) - G[tag + "_logpdf"]).max() <= 1e-11


@pytest.mark.parametrize("tag", ["d2_mean", "d4_mean_l3"])
def test_loss_and_parameter_gradient_equal_the_reference_source(tag):
    """vqmc.loss_fn_efficient (value) and <grad loss, v> for random parameter directions v: the reference's own loss code,
    with the gradient estimator it registers through custom_jvp, evaluated on parameters seeded with dual numbers."""
    m, params, D = _model(tag)
    x, ra = G[tag + "_loss_x"], float(G[tag + "_loss_running_average"])
    loss, g = ograd.loss_and_grad(m, params, x, np.zeros((D, 1)), ra)
    assert abs(loss - float(G[tag + "_loss"])) <= 1e-11 * abs(loss)
    gl = _leaves(g, [])
    n_dirs = len([k for k in G.files if k.startswith(tag + "_dloss")])
    assert n_dirs >= 1
    for k in range(n_dirs):
        d = sum(float((gl[i] * G[f"{tag}_dir{k}_{i:03d}"]).sum()) for i in range(len(gl)))
        ref = float(G[f"{tag}_dloss{k}"])
        assert abs(d - ref) <= 1e-10 * abs(ref), (k, d, ref)
