"""ORACLE helpers (test infrastructure): golden-fixture loading and model construction.

Tables come from waveflow_b200.splines.tablegen, which is itself pinned bit-for-bit against the
reference generator (tests/test_tablegen_golden.py), so the oracle does not carry a second generator.
"""
from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

from waveflow_b200.splines import tablegen as tg
from . import live

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"


@lru_cache(maxsize=None)
def tables_I(k, n):
    return tg.build_I_tables(k, n)[0]


@lru_cache(maxsize=None)
def tables_M(k, n):
    return tg.build_M_tables(k, n)[0]


@lru_cache(maxsize=None)
def _tables_B(k, n):
    return tg.build_B_tables(k, n)


def tables_B(k, n):
    return _tables_B(k, n)


def net_from_arrays(W1, b1, W2, b2, W3, b3, zero):
    """The reference's pytree for one masked_transform: (stax.serial params, zero_params)."""
    return ([(W1, b1), (), (W2, b2), (), (W3, b3)], zero)


def load_he_checkpoint():
    """-> (params pytree in the reference's structure, dict of golden outputs)."""
    z = np.load(GOLDEN / "he_checkpoint_epoch100000.npz")
    names = ["W1", "b1", "W2", "b2", "W3", "b3", "zero"]
    tp = [()]
    for li in range(int(z["n_imade"])):
        tp.append(net_from_arrays(*[z[f"imade{li}_{n}"] for n in names]))
        tp.append(())
    sp = net_from_arrays(*[z[f"prior_{n}"] for n in names])
    gold = {k: z[k] for k in ["psi_grid", "onproton_coord", "onproton_values", "random_coord", "random_values",
                              "samples", "loss_tail"]}
    return (tp, sp), gold


def waveflow_model(D, degree=6, n_knots=23, n_layers=3, box=10.0, reg=0.05, tol=1e-6, coord="mean",
                   dtype=np.float64) -> live.LiveModel:
    """model_factory.get_waveflow_model as configured by vqmc.create_train_state (vqmc.py:128-132)."""
    B = tables_B(degree, n_knots)
    m = live.LiveModel(D=D, n_layers=n_layers, k_i=degree, tab_I=tables_I(degree, n_knots), reg=reg, tol=tol,
                       bc_i_left={0: 0}, bc_i_right={0: 1}, prior="B", k_p=degree, tab_P=B["b"], tab_OB=B["ob"],
                       ob_to_b=B["ob_to_b"], b_to_ob=B["b_to_ob"], bc_p_left={0: 0}, bc_p_right={0: 0},
                       box=box, coord=coord)
    return m.cast(dtype)


def mflow_model(D=2, i_degree=5, i_knots=23, n_layers=3, reg=0.02, tol=1e-6, p_degree=3, p_knots=15,
                bc_i_left=None, bc_i_right=None, bc_p_left=None, bc_p_right=None, dtype=np.float64) -> live.LiveModel:
    """benchmark_tests.get_model('MFlow') (benchmark_tests.py:67-72): IMADE defaults {0:0}|{0:1}, MFlow {0:0}|{0:0}."""
    m = live.LiveModel(D=D, n_layers=n_layers, k_i=i_degree, tab_I=tables_I(i_degree, i_knots), reg=reg, tol=tol,
                       bc_i_left={0: 0.0} if bc_i_left is None else bc_i_left,
                       bc_i_right={0: 1.0} if bc_i_right is None else bc_i_right,
                       prior="M", k_p=p_degree, tab_P=tables_M(p_degree, p_knots),
                       bc_p_left={0: 0} if bc_p_left is None else bc_p_left,
                       bc_p_right={0: 0} if bc_p_right is None else bc_p_right, box=None)
    return m.cast(dtype)


def random_net(rng: np.random.Generator, D, P, hidden=64, scale=1.0):
    """W, b ~ U(+-1/sqrt(fan_in)) (model_factory.py:25-28), zero_params ~ U(-.5,.5) (:84); float32."""
    def u(shape, fan_in):
        b = scale / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)
    return net_from_arrays(u((D, hidden), D), u((hidden,), D), u((hidden, hidden), hidden), u((hidden,), hidden),
                           u((hidden, D * P), hidden), u((D * P,), hidden),
                           rng.uniform(-0.5, 0.5, size=(D, P)).astype(np.float32))


def random_params(rng, m: live.LiveModel, scale=1.0):
    tp = [()] if m.box is not None else []
    for _ in range(m.n_layers):
        tp.append(random_net(rng, m.D, m.P_I, scale=scale))
        tp.append(())
    sp = random_net(rng, m.D, m.P_P, scale=scale) if m.prior in ("B", "M") else ()
    return (tp, sp)


def cast_params(params, dtype):
    def c(o):
        if isinstance(o, np.ndarray):
            return o.astype(dtype)
        if isinstance(o, (list, tuple)):
            return type(o)(c(x) for x in o)
        return o
    return c(params)
