"""ORACLE (test infrastructure, never imported by the product path).

Literal CPU restatement (numpy) of the reference's live spline-flow path: table-interpolated
I-/M-/B-spline operators, MADE conditioner, IMADE layers, box transform, MFlow density and the
square-normalised Waveflow wavefunction.  Each function cites the reference lines it follows
(paths relative to /root/reference/waveflow).  ``dtype`` selects float32 (the reference's default,
JAX x64 disabled) or float64 (used as the high-precision yardstick in parity tests).

Pinned by: tests/golden/ref_tables_deg5_k16.npz (shipped basis tables, bit-exact), tests/golden/he_checkpoint_epoch100000.npz
(published parameters + psi grids) -- see tests/test_oracle_golden.py -- and, for the spline operators (remove_bias,
enforce_boundary_conditions, apply / apply_grad, bisection), tests/golden/ref_spline_vectors.npz: outputs of the reference's OWN
isplines_jax.py / bsplines_jax.py / helpers.binary_search executed on a numpy stand-in for jax (tests/golden/make_spline_golden.py),
reproduced bit for bit in float32 (tests/test_spline_reference_vectors.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------- A3: table lookup
def table_lookup(tab: np.ndarray, nd: int, x: np.ndarray) -> np.ndarray:
    """{I,M,B}_cached for ALL bases at once: returns [P, M].   splines/isplines_jax.py:45-56.

    tab: [4, P, T] already in the working dtype.  JAX gather semantics: an index above
    T-1 clamps to T-1, index -1 wraps to T-1.  nd above 3 clamps to 3 (quirk Q5).
    """
    dt = tab.dtype.type
    T = tab.shape[-1]
    np_ = T - 1
    nd = min(nd, tab.shape[0] - 1)
    x = np.asarray(x, dtype=tab.dtype)
    xs = x * dt(np_)
    x_l = np.floor(xs).astype(np.int32)
    x_r = np.ceil(xs).astype(np.int32)

    def fix(i):
        i = np.where(i < 0, i + T, i)
        return np.clip(i, 0, T - 1)
    y_l = tab[nd][:, fix(x_l)]
    y_r = tab[nd][:, fix(x_r)]
    dx = x - (x_l.astype(tab.dtype) / dt(np_))
    slope = (y_r - y_l) * dt(np_)
    return y_l + slope * dx


def spline_apply(tab: np.ndarray, c: np.ndarray, x: np.ndarray, nd: int = 0) -> np.ndarray:
    """apply_fun_vec (nd=0) / apply_fun_vec_grad (nd=1):  sum_i c[:, i] * basis_i^{(nd)}(x).

    isplines_jax.py:69-79,139-149 (zero_border=False), msplines_jax.py:53-62,116-126, bsplines_jax.py:42-45.
    The derivative w.r.t. x is *the same lookup in table nd+1* (custom_jvp, isplines_jax.py:60-66; quirk Q4).
    Python's ``sum`` adds left to right starting from 0.
    """
    basis = table_lookup(tab, nd, x)            # [P, M]
    acc = np.zeros_like(basis[0])
    for i in range(c.shape[-1]):
        acc = acc + c[:, i] * basis[i]
    return acc


def bspline_coeffs(w: np.ndarray, ob_to_b: np.ndarray) -> np.ndarray:
    """bsplines_jax.py:134-135:  c = w @ ob_to_b;  c /= ||c||_2."""
    c = w @ ob_to_b
    return c / np.sqrt(np.sum(c ** 2, axis=-1, keepdims=True))


# --------------------------------------------------------------------------- a3: remove_bias
def remove_bias_I(p: np.ndarray, k: int) -> np.ndarray:
    """isplines_jax.py:196-200."""
    p = p.copy()
    dt = p.dtype.type
    for i in range(k):
        p[:, i + 1] = p[:, i + 1] * dt(i + 1) / dt(k)
        p[:, -(i + 2)] = p[:, -(i + 2)] * dt(i + 1) / dt(k)
    return p / p.sum(-1, keepdims=True)


def remove_bias_M(p: np.ndarray, k: int) -> np.ndarray:
    """msplines_jax.py:186-190."""
    p = p.copy()
    dt = p.dtype.type
    for i in range(k):
        p[:, i] = p[:, i] * dt(i + 1) / dt(k)
        p[:, -(i + 1)] = p[:, -(i + 1)] * dt(i + 1) / dt(k)
    return p / p.sum(-1, keepdims=True)


# --------------------------------------------------------------------------- a4: boundary conditions
def _at(tab, nd, j, x):
    """scalar basis value {I,M,B}_cached(x, j, tables, n_derivative=nd)."""
    return table_lookup(tab, nd, np.array([x], dtype=tab.dtype))[j, 0]


def enforce_bc(tab: np.ndarray, w: np.ndarray, left: dict, right: dict, kind: str) -> np.ndarray:
    """enforce_boundary_conditions for kind in {'I','M','B'}.

    isplines_jax.py:158-192 (right {0: 1.0} special-cased to w[-1] = 0, :174-176),
    msplines_jax.py:156-183, bsplines_jax.py:173-198 (B normalises by ||w||_2, :198).
    ``tab`` is the *plain* basis table (for 'B' the non-orthogonalised B tables).
    """
    w = w.copy()
    dt = w.dtype.type
    P = w.shape[-1]
    for nd, val in left.items():
        prev = [_at(tab, nd, j, 0.0) for j in range(nd)]
        value = _at(tab, nd, nd, 0.0)
        s = np.zeros_like(w[:, 0])
        for pv, j in zip(prev, range(nd)):
            s = s + pv * w[:, j]
        w[:, nd] = (dt(val) - s) / value
    for nd, val in right.items():
        if kind == "I" and nd == 0:
            if val != 1:
                raise ValueError("Only constraint value of 1.0 is supported (isplines_jax.py:177-179)")
            w[:, P - 1] = dt(0.0)
            continue
        prev = [_at(tab, nd, P - j - 1, 1.0) for j in range(nd)]
        value = _at(tab, nd, P - nd - 1, 1.0)
        s = np.zeros_like(w[:, 0])
        for pv, j in zip(prev, range(nd)):
            s = s + pv * w[:, P - 1 - j]          # zip(prev, flip(w))
        w[:, P - nd - 1] = (dt(val) - s) / value
    if kind == "B":
        return w / np.sqrt(np.sum(w ** 2, axis=-1, keepdims=True))
    return w / w.sum(-1, keepdims=True)


# --------------------------------------------------------------------------- a7: bisection
def binary_search_inverse(tab: np.ndarray, c: np.ndarray, y: np.ndarray, tol: float) -> np.ndarray:
    """reverse_fun_vec: utils/helpers.py:150-166 on f(x) = ispline(x) - y, vmapped (isplines_jax.py:153-156).

    Returns the LOWER bracket.  A vmapped lax.while_loop runs every lane until its own predicate fails.
    """
    dt = tab.dtype.type
    lo = np.zeros_like(y)
    hi = np.ones_like(y)
    half_tol = dt(tol) / dt(2)
    while True:
        mid = dt(0.5) * (lo + hi)
        active = (lo + half_tol < mid) & (mid < hi - half_tol)
        if not active.any():
            return lo
        upper = (spline_apply(tab, c, mid, 0) - y) > 0
        lo = np.where(active & ~upper, mid, lo)
        hi = np.where(active & upper, mid, hi)


# --------------------------------------------------------------------------- A4: MADE conditioner
def made_masks(D: int, hidden: int = 64):
    """model_factory.py:8-19 (num_hidden=1)."""
    deg = [np.arange(D), np.arange(hidden) % (D - 1), np.arange(hidden) % (D - 1), np.arange(D) % D - 1]
    return [(d1[:, None] >= d0[None, :]).T.astype(np.float32) for d0, d1 in zip(deg[:-1], deg[1:])]


def conditioner(net, x: np.ndarray, P: int, allow_negative: bool, grad_to_zero: bool = False) -> np.ndarray:
    """masked_transform.calculate_bijection_params  (model_factory.py:56-70, MaskedDense :31-33).

    net = ([(W1,b1),(),(W2,b2),(),(W3,b3)], zero_params).  Returns [N, D, P].
    """
    nn, zero = net
    (W1, b1), _, (W2, b2), _, (W3, b3) = nn
    dt = x.dtype
    D = x.shape[-1]
    m1, m2, m3 = made_masks(D, W1.shape[1])
    m3 = np.tile(m3, P)
    h = np.tanh(x @ (W1.astype(dt) * m1.astype(dt)) + b1.astype(dt))
    h = np.tanh(h @ (W2.astype(dt) * m2.astype(dt)) + b2.astype(dt))
    o = h @ (W3.astype(dt) * m3.astype(dt)) + b3.astype(dt)
    p = np.stack(np.split(o, o.shape[-1] // D, axis=-1), axis=-1)       # p[n, d, q] = o[n, q*D + d]
    zero = zero.astype(dt)
    if not allow_negative:
        p = 1 / (1 + np.exp(-p))
        zero = np.abs(zero)
    if grad_to_zero:
        g = np.roll(np.cumprod(x ** 3, axis=-1), 1, axis=-1)
        g[:, 0] = 1
        p = g[..., None] * p + zero
    return p / p.sum(-1, keepdims=True)


# --------------------------------------------------------------------------- model description
@dataclass
class LiveModel:
    """Static description + tables of one model built by model_factory.get_model / get_waveflow_model."""
    D: int
    n_layers: int
    k_i: int
    tab_I: np.ndarray                  # [4, P_I, T]
    reg: float = 0.0
    tol: float = 1e-6
    bc_i_left: dict = field(default_factory=dict)
    bc_i_right: dict = field(default_factory=dict)
    grad_to_zero: bool = False
    # prior
    prior: str = "B"                   # 'B' (Waveflow) | 'M' (MFlow) | 'uniform' (Flow/IFlow with Uniform prior)
    k_p: int = 0
    tab_P: np.ndarray | None = None    # M tables, or plain B tables (for the BC)
    tab_OB: np.ndarray | None = None
    ob_to_b: np.ndarray | None = None
    b_to_ob: np.ndarray | None = None
    bc_p_left: dict = field(default_factory=dict)
    bc_p_right: dict = field(default_factory=dict)
    # box
    box: float | None = None           # None: no BoxTransformLayer (density models)
    coord: str = "mean"

    def cast(self, dtype):
        import copy
        m = copy.copy(self)
        for nm in ["tab_I", "tab_P", "tab_OB", "ob_to_b", "b_to_ob"]:
            a = getattr(self, nm)
            if a is not None:
                setattr(m, nm, a.astype(dtype))
        return m

    @property
    def P_I(self):
        return self.tab_I.shape[1]

    @property
    def P_P(self):
        return (self.tab_OB if self.prior == "B" else self.tab_P).shape[1]


# --------------------------------------------------------------------------- a8: box transform
def box_direct(x: np.ndarray, L: float, coord: str):
    """BoxTransformLayer.direct_fun_{mean,first}  (flows/bijections/made.py:156-183, :118-137)."""
    dt = x.dtype.type
    tolr = dt(1e-7)
    N, D = x.shape
    out = np.ones_like(x)
    if coord == "mean":
        mean = x.mean(-1)
        l = mean - x[:, 0]
        w = x[:, -1] - x[:, 0]
        space_left = np.full(N, 2 * L, dtype=x.dtype)
        ld = np.zeros(N, dtype=x.dtype)
        for i in range(D - 1):
            diff = x[:, i + 1] - x[:, i]
            out[:, i] = diff / (space_left + tolr)
            ld = ld - np.log(space_left + tolr)
            space_left = space_left - diff
        out[:, -1] = (mean + dt(L) - l) / (dt(2 * L) - w + tolr)
        ld = ld - np.log(dt(2 * L) - w + tolr)
        return out, ld
    out[:, 0] = (x[:, 0] + dt(L)) / dt(2 * L)
    for i in range(1, D):
        out[:, i] = (x[:, i] - x[:, i - 1]) / (dt(L) - x[:, i - 1] + tolr)
    ld = -np.log(dt(2 * L)) - np.log(dt(L) - x[:, :-1] + tolr).sum(-1)
    return out, ld


def box_inverse(u: np.ndarray, L: float, coord: str) -> np.ndarray:
    """reverse_fun_{mean,first}  (made.py:186-197 -- only correct for D=2, quirk Q2; :139-154)."""
    dt = u.dtype.type
    if coord == "mean":
        out = np.zeros_like(u)
        out[:, 1:] = np.cumsum(u[:, :-1], axis=-1)
        mean = out.mean(-1)
        w = out[:, -1]
        pm = u[:, -1] * (1 - w) - (dt(0.5) - mean)
        return (out - mean[:, None] + pm[:, None]) * dt(2) * dt(L)
    x = u.copy()
    x[:, 0] = (x[:, 0] - dt(0.5)) * dt(2 * L)
    for i in range(1, u.shape[-1]):
        x[:, i] = x[:, i] * (dt(L) - x[:, i - 1]) + x[:, i - 1]
    return x


# --------------------------------------------------------------------------- a6/a7: IMADE
def imade_coeffs(m: LiveModel, net, x: np.ndarray) -> np.ndarray:
    """made.py:67-72: conditioner -> +reg -> remove_bias -> enforce_boundary_conditions.  [N*D, P]."""
    p = conditioner(net, x, m.P_I, allow_negative=False, grad_to_zero=m.grad_to_zero)
    p = p + x.dtype.type(m.reg)
    p = remove_bias_I(p.reshape(-1, m.P_I), m.k_i)
    return enforce_bc(m.tab_I, p, m.bc_i_left, m.bc_i_right, "I")


def imade_direct(m: LiveModel, net, x: np.ndarray):
    """IMADE.direct_fun  (made.py:66-81)."""
    N, D = x.shape
    c = imade_coeffs(m, net, x)
    y = spline_apply(m.tab_I, c, x.reshape(-1), 0).reshape(N, D)
    dy = spline_apply(m.tab_I, c, x.reshape(-1), 1).reshape(N, D)
    return y, np.log(dy + x.dtype.type(1e-7)).sum(-1)


def imade_inverse(m: LiveModel, net, y: np.ndarray) -> np.ndarray:
    """IMADE.inverse_fun (made.py:85-100): coefficients come from the conditioner of the *inputs* y (quirk Q1)."""
    N, D = y.shape
    c = imade_coeffs(m, net, y).reshape(N, D, -1)
    out = np.zeros_like(y)
    for d in range(D):
        out[:, d] = binary_search_inverse(m.tab_I, c[:, d, :], y[:, d], m.tol)
    return out


def flow_direct(m: LiveModel, transform_params, x: np.ndarray):
    """Serial(BoxTransformLayer?, (IMADE, Reverse) * L).direct_fun  (bijections.py:452-460,336-345)."""
    ld = np.zeros(x.shape[0], dtype=x.dtype)
    nets = [p for p in transform_params if len(p)]
    u = x
    if m.box is not None:
        u, l0 = box_direct(u, m.box, m.coord)
        ld = ld + l0
    for net in nets:
        u, l1 = imade_direct(m, net, u)
        ld = ld + l1
        u = u[:, ::-1]
    return np.ascontiguousarray(u), ld


def flow_inverse(m: LiveModel, transform_params, u: np.ndarray) -> np.ndarray:
    """Serial.inverse_fun  (bijections.py:462-463): reversed layers: Reverse, IMADE.inverse, ..., Box inverse."""
    nets = [p for p in transform_params if len(p)]
    x = u
    for net in reversed(nets):
        x = np.ascontiguousarray(x[:, ::-1])
        x = imade_inverse(m, net, x)
    if m.box is not None:
        x = box_inverse(x, m.box, m.coord)
    return x


# --------------------------------------------------------------------------- a10/a11: densities, psi
def prior_coeffs(m: LiveModel, sp_params, u: np.ndarray) -> np.ndarray:
    """Coefficients of the prior spline, [N*D, P].  wavefunctions.py:58-62 / distributions.py:146-154."""
    P = m.P_P
    if m.prior == "B":
        w = conditioner(sp_params, u, P, allow_negative=True, grad_to_zero=m.grad_to_zero).reshape(-1, P)
        return enforce_bc(m.tab_P, w, m.bc_p_left, m.bc_p_right, "B")
    w = conditioner(sp_params, u, P, allow_negative=False, grad_to_zero=m.grad_to_zero).reshape(-1, P)
    w = remove_bias_M(w, m.k_p)
    return enforce_bc(m.tab_P, w, m.bc_p_left, m.bc_p_right, "M")


def prior_factors(m: LiveModel, sp_params, u: np.ndarray) -> np.ndarray:
    """phi_d(u_d) for 'B' (bsplines_jax.py:127-137) or the M-spline density for 'M'.  [N, D]."""
    N, D = u.shape
    w = prior_coeffs(m, sp_params, u)
    uc = np.clip(u, 0.0, 1.0).reshape(-1)
    if m.prior == "B":
        return spline_apply(m.tab_OB, bspline_coeffs(w, m.ob_to_b), uc, 0).reshape(N, D)
    return spline_apply(m.tab_P, w, uc, 0).reshape(N, D)


def psi(m: LiveModel, params, x: np.ndarray) -> np.ndarray:
    """Waveflow.psi  (wavefunctions.py:54-71); constrained dims 0..D-2 divided by sqrt(2) ('mean' coords)."""
    tp, sp = params
    u, ld = flow_direct(m, tp, x)
    phi = prior_factors(m, sp, u)
    dt = x.dtype.type
    cons = _constrained(m)
    phi[:, cons] = phi[:, cons] / np.sqrt(dt(2))
    return np.prod(phi, axis=-1) * np.exp(dt(0.5) * ld)


def _constrained(m: LiveModel):
    """model_factory.py:124-129."""
    return np.arange(0, m.D - 1) if m.coord == "mean" else np.arange(1, m.D)


def log_pdf(m: LiveModel, params, x: np.ndarray, return_sample: bool = False):
    """Waveflow.log_pdf (wavefunctions.py:33-52) / MFlow.log_pdf (distributions.py:139-163)."""
    tp, sp = params
    u, ld = flow_direct(m, tp, x)
    dt = x.dtype.type
    if m.prior == "uniform":
        uc = np.clip(u, 0.0, 1.0)                           # Flow(prior_support=(0,1)), distributions.py:97-99
        lp = np.zeros(x.shape[0], dtype=x.dtype) + ld       # uniform.logpdf == 0 on [0,1]
        return (lp, uc) if return_sample else lp
    pr = prior_factors(m, sp, u)
    if m.prior == "B":
        pr = pr ** 2
        cons = _constrained(m)
        pr[:, cons] = pr[:, cons] / dt(2)
    lp = np.log(pr + dt(1e-7)).sum(-1) + ld
    if return_sample:
        return lp, np.clip(u, 0.0, 1.0)
    return lp


# --------------------------------------------------------------------------- f2: samplers (statistical parity only)
def mspline_rejection_sample(m: LiveModel, c: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """MSpline_fun.sample_fun_vec with num_samples = 1 (msplines_jax.py:129-154): per row, propose x ~ U(0,1), y ~ U(0, ymax)
    with the convex-hull bound ymax = max(c) * n_knots (:145-148) until y < sum_q c_q M_q(x).  c [N, P] -> x [N].
    The random stream is numpy's, not JAX's threefry: the draws agree with the reference in distribution only."""
    N, P = c.shape
    ymax = c.max(-1) * c.dtype.type(P + m.k_p)              # n_knots = len(knots) = P + k (msplines_jax.py:72-75)
    x = np.zeros(N, dtype=c.dtype)
    todo = np.arange(N)
    while todo.size:
        xs = rng.uniform(0.0, 1.0, todo.size).astype(c.dtype)
        ys = rng.uniform(0.0, 1.0, todo.size).astype(c.dtype) * ymax[todo]
        ok = ys < spline_apply(m.tab_P, c[todo], xs, 0)
        x[todo[ok]] = xs[ok]
        todo = todo[~ok]
    return x


def mflow_sample(m: LiveModel, params, n: int, rng: np.random.Generator, dtype=np.float32):
    """MFlow.sample (distributions.py:165-190): D rounds of (prior conditioner on the columns drawn so far -> rejection
    sample of column i), then Serial.inverse_fun.  -> (x [n, D] data space, u [n, D] prior space)."""
    tp, sp = params
    u = np.zeros((n, m.D), dtype=dtype)
    for i in range(m.D):
        c = prior_coeffs(m, sp, u).reshape(n, m.D, m.P_P)[:, i, :]
        u[:, i] = mspline_rejection_sample(m, np.ascontiguousarray(c), rng)
    return flow_inverse(m, tp, u), u


# --------------------------------------------------------------------------- a12: potential
def potential(x: np.ndarray, protons: np.ndarray) -> np.ndarray:
    """get_potential (utils/physics.py:60-76): soft-Coulomb, 1 space dimension per particle."""
    pe = -(1 / np.sqrt(1 + (protons[None] - x[:, None]) ** 2)).sum(-1).sum(-1)
    diff = x[:, :, None] - x[:, None, :]
    il = np.tril_indices(x.shape[1], k=-1)
    ee = (1 / np.sqrt(1 + diff ** 2)[:, il[0], il[1]]).sum(-1)
    return pe + ee
