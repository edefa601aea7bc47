"""T1: the oracle against the reference's published known answers -- CPU only."""
import numpy as np
import pytest

from oracle import fixtures as fx
from oracle import laplacian as lap
from oracle import live, rqs


@pytest.fixture(scope="module")
def he():
    return fx.load_he_checkpoint()


def _grid():
    y, x = np.meshgrid(np.linspace(-10, 10, 100), np.linspace(-10, 10, 100))      # utils/helpers.py:52-54
    c = np.stack([x, y], -1).reshape(-1, 2)
    return c, (c[:, 0] > c[:, 1]).astype(int)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 2e-5), (np.float32, 3e-5)])
def test_psi_grid_known_answer(he, dtype, tol):
    params, gold = he
    m = fx.waveflow_model(2, dtype=dtype)
    c, inv = _grid()
    z = live.psi(m, fx.cast_params(params, dtype), np.sort(c, -1).astype(dtype)) * (-1.0) ** inv
    assert np.abs(z - gold["psi_grid"]).max() < tol          # published fp32 values, max |psi| = 1.53


def test_density_cuts_known_answer(he):
    params, gold = he
    m = fx.waveflow_model(2)
    for nm, tol in [("onproton", 1e-6), ("random", 2e-5)]:
        c = gold[nm + "_coord"]
        z = live.psi(m, params, np.sort(c, -1).astype(np.float64)) * (-1.0) ** (c[:, 0] > c[:, 1])
        assert np.abs(z - gold[nm + "_values"]).max() < tol


def test_normalisation_property(he):
    params, _ = he
    m = fx.waveflow_model(2)
    c, _ = _grid()
    sc = np.sort(c, -1)
    lp = live.log_pdf(m, params, sc)
    ps = live.psi(m, params, sc)
    assert np.allclose(np.exp(lp), ps ** 2, rtol=1e-4, atol=1e-6)         # log_pdf == log psi^2 away from the floor
    assert abs(np.exp(lp).sum() * (20 / 99) ** 2 - 1.0) < 0.12            # test_waveflow.py:52 prints this integral


def test_laplacian_oracles_agree(he):
    params, gold = he
    m = fx.waveflow_model(2)
    x = np.sort(gold["samples"][:48], -1).astype(np.float64)
    prot = np.array([[0.0], [0.0]])
    a = lap.local_energy_bundle(m, params, x, prot)
    b = lap.local_energy_autograd(m, params, x, prot)
    for k in ["psi", "grad", "lap", "hpsi"]:
        assert np.abs(a[k] - b[k]).max() <= 1e-10 * max(1.0, np.abs(b[k]).max()), k
    assert np.allclose(a["psi"], live.psi(m, params, x), rtol=0, atol=1e-12)


def test_energy_of_published_samples_matches_published_trace(he):
    # the shipped checkpoint is the late, diverged state of the run: E_loc on its own samples is far from -1.81 but
    # consistent with the tail of the published loss trace (mean of the last 100 steps: -441, min -809, max +612)
    params, gold = he
    m = fx.waveflow_model(2)
    e = lap.local_energy_bundle(m, params, np.sort(gold["samples"], -1).astype(np.float64), np.array([[0.0], [0.0]]))["eloc"]
    assert -900 < e.mean() < -300 and gold["loss_tail"][-100:].min() > -1000


def test_laplacian_D4_first_coords():
    m = fx.waveflow_model(3, coord="first")
    p = fx.random_params(np.random.default_rng(3), m)
    x = np.sort(np.random.default_rng(4).uniform(-9, 9, (6, 3)), -1)
    a = lap.local_energy_bundle(m, p, x, np.zeros((3, 1)))
    b = lap.local_energy_autograd(m, p, x, np.zeros((3, 1)))
    assert np.abs(a["lap"] - b["lap"]).max() <= 1e-9 * max(1.0, np.abs(b["lap"]).max())


def test_imade_inverse_is_reference_quirk_Q1():
    # D=2: dimension 0 is conditioned on nothing, so it inverts exactly; dimension 1 is conditioned on the *input*
    m = fx.mflow_model(n_layers=1)
    p = fx.random_params(np.random.default_rng(0), m, scale=8.0)
    x = np.random.default_rng(1).uniform(0.05, 0.95, (200, 2))
    y, _ = live.imade_direct(m, p[0][0], x)
    xi = live.imade_inverse(m, p[0][0], y)
    assert np.abs(xi[:, 0] - x[:, 0]).max() < 2e-6
    assert np.abs(xi[:, 1] - x[:, 1]).max() > 1e-4


def test_rqs_bijective_and_derivative():
    rng = np.random.default_rng(0)
    N, K, B = 2000, 32, 3.0
    x = rng.uniform(-4, 4, N)
    uw, uh = rng.standard_normal((2, N, K)); ud = rng.standard_normal((N, K - 1))
    y, ld, b1 = rqs.unconstrained_rqs(x, uw, uh, ud, False, B, True)
    x2, ld2, b2 = rqs.unconstrained_rqs(y, uw, uh, ud, True, B, True)
    assert np.abs(x2 - x).max() < 1e-10 and np.abs(ld + ld2).max() < 1e-9 and np.array_equal(b1, b2)
    assert np.all(y[np.abs(x) > B] == x[np.abs(x) > B]) and np.all(ld[np.abs(x) > B] == 0) and np.all(b1[np.abs(x) > B] == -1)
    h = 1e-6
    ins = np.abs(x) < B - 1e-3
    fd = np.log((rqs.unconstrained_rqs(x + h, uw, uh, ud, False, B)[0] - rqs.unconstrained_rqs(x - h, uw, uh, ud, False, B)[0]) / (2 * h))
    assert np.median(np.abs(fd - ld)[ins]) < 1e-8
    # boundary derivatives are pinned to 1 (neural_splines.py:33-42)
    e = 1e-7
    yb = rqs.unconstrained_rqs(np.full(4, -B + e), uw[:4], uh[:4], ud[:4], False, B)
    assert np.abs(yb[1]).max() < 1e-3


def test_rqs_coupling_bijective_atol_1e_3():
    # the reference's own property test: tests/test_bijections.py:12-21 (atol 1e-3), here in float32
    rng = np.random.default_rng(0)
    D, K, B = 4, 5, 3.0
    layers = [(rqs.random_fcnn(rng, D // 2, 8, (3 * K - 1) * D // 2), rqs.random_fcnn(rng, D // 2, 8, (3 * K - 1) * D // 2))]
    x = rng.uniform(-10, 10, (20, D)).astype(np.float32)
    z, l1 = rqs.coupling_flow_direct(layers, x, K, B)
    x2, l2 = rqs.coupling_flow_inverse(layers, z, K, B)
    assert z.shape == x.shape and l1.shape == (20,)
    assert np.allclose(x, x2, atol=1e-3)


def test_gradient_oracle_against_finite_differences(he):
    """oracle/grad.py (three reverse passes) vs central differences of the surrogate through the independent numpy
    bundle oracle, on prior-net parameters: their gradient does not pass through a table argument, so the custom_jvp
    table derivative (next table, Q5 clamp) and the true slope of the interpolant coincide there."""
    from oracle import grad as ograd
    params, _ = he
    m = fx.waveflow_model(2)
    p = fx.cast_params(params, np.float64)
    x = np.sort(np.random.default_rng(1).uniform(-3, 3, (6, 2)), axis=1)
    protons = np.array([[0.0], [0.0]])
    loss, g = ograd.loss_and_grad(m, p, x, protons, -1.8)
    r = lap.local_energy_bundle(m, p, x, protons)
    eloc, a, b = ograd.coefficients(r["psi"], r["hpsi"], -1.8)
    assert abs(loss - eloc.mean()) < 1e-9 * abs(loss)
    for layer, idx in [((4, 1), (5,)), ((4, 0), (3, 7)), ((2, 0), (3, 7)), ((0, 0), (0, 7))]:
        arr = p[1][0][layer[0]][layer[1]]
        orig = arr[idx]
        vals = []
        for h in (1e-6, -1e-6):
            arr[idx] = orig + h
            vals.append(ograd.surrogate_value(m, p, x, protons, a, b))
        arr[idx] = orig
        fd = (vals[0] - vals[1]) / 2e-6
        got = g[1][0][layer[0]][layer[1]][idx]
        assert abs(got - fd) <= 1e-6 * abs(fd) + 1e-7, (layer, got, fd)
    # flow-layer gradients exist and are finite; zero_params get zeros
    for net in [n for n in g[0] if len(n)]:
        assert all(np.all(np.isfinite(a_)) for lay in net[0] for a_ in lay) and not np.any(net[1])


def test_inversion_count_matches_definition():
    from waveflow_b200.utils.coordinates import get_num_inversion_count
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 5))
    brute = [sum(1 for i in range(5) for j in range(i + 1, 5) if r[i] > r[j]) for r in x]
    assert get_num_inversion_count(x).tolist() == brute
    assert get_num_inversion_count(np.array([[2.0, 1.0]])).tolist() == [1]
