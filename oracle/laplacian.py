"""ORACLE (test infrastructure): local energy of the Waveflow wavefunction.

The reference computes  H psi = -1/2 tr(d^2 psi / dx^2) + V psi  with jax.hessian (utils/physics.py:50-52,
79-93) and forms  E_loc = H psi / (psi + 1e-8)  in the loss (vqmc.py:198-200).  JAX differentiates the
table lookup through its custom_jvp: d/dx lookup(nd) := lookup(nd+1) (splines/isplines_jax.py:60-66, quirk Q4)
-- NOT the slope of the linear interpolant.  Two independent restatements live here:

* ``local_energy_bundle``  numpy float64 forward-Laplacian propagation of (value, grad[D], laplacian) bundles
  through the restated forward pass (same algebra as the CUDA kernel, written independently in numpy);
* ``local_energy_autograd`` torch float64 double reverse-mode autodiff of the restated psi, where the table lookup
  is an autograd.Function whose backward is the lookup in table nd+1 (what jax.hessian sees).

PINNED BY THE REFERENCE'S OWN SOURCE: the reference stores no H psi anywhere and JAX cannot be installed here,
but tests/golden/make_energy_golden.py executes the unmodified model_factory -> wavefunctions.Waveflow -> flows -> isplines_jax /
bsplines_jax -> utils/physics.construct_hamiltonian_function on a numpy stand-in for jax (float64): jax.hessian becomes forward
over forward mode with nested, level-tagged dual numbers carried through the reference's own arithmetic, and every table lookup
on a dual number applies the custom_jvp rule the reference registered, at both levels.  Both restatements reproduce the stored
psi / log_pdf / H psi to 1e-12 for a D = 2 ('mean' coordinates), a D = 3 ('first') and a three-layer D = 4 ('mean') model (tests/test_energy_reference_vectors.py).
Further pins: the two restatements agree with each other, the psi KAT, the published energy plateau.
"""
from __future__ import annotations

import numpy as np

from . import live


# =========================================================================== bundles
class Bundle:
    """(v [N], g [N, D], l [N]) = value, gradient and Laplacian w.r.t. the D walker coordinates."""
    __slots__ = ("v", "g", "l")

    def __init__(self, v, g, l):
        self.v, self.g, self.l = v, g, l

    @staticmethod
    def const(c, N, D):
        return Bundle(np.full(N, float(c)), np.zeros((N, D)), np.zeros(N))

    @staticmethod
    def coord(x, d):
        N, D = x.shape
        g = np.zeros((N, D)); g[:, d] = 1.0
        return Bundle(x[:, d].astype(np.float64).copy(), g, np.zeros(N))

    def _lift(self, o):
        if isinstance(o, Bundle):
            return o
        return Bundle(np.broadcast_to(np.asarray(o, dtype=np.float64), self.v.shape), np.zeros_like(self.g),
                      np.zeros_like(self.l))

    def __add__(self, o):
        o = self._lift(o); return Bundle(self.v + o.v, self.g + o.g, self.l + o.l)
    __radd__ = __add__

    def __neg__(self):
        return Bundle(-self.v, -self.g, -self.l)

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        if not isinstance(o, Bundle):
            o = np.asarray(o, dtype=np.float64)
            return Bundle(self.v * o, self.g * (o[..., None] if o.ndim else o), self.l * o)
        return Bundle(self.v * o.v, self.v[:, None] * o.g + o.v[:, None] * self.g,
                      self.v * o.l + o.v * self.l + 2.0 * (self.g * o.g).sum(-1))
    __rmul__ = __mul__

    def apply(self, f0, f1, f2):
        """Unary f with f(v)=f0, f'(v)=f1, f''(v)=f2."""
        return Bundle(f0, f1[:, None] * self.g, f1 * self.l + f2 * (self.g ** 2).sum(-1))

    def recip(self):
        r = 1.0 / self.v
        return self.apply(r, -r * r, 2.0 * r * r * r)

    def __truediv__(self, o):
        if not isinstance(o, Bundle):
            return self * (1.0 / o)
        return self * o.recip()

    def log(self):
        r = 1.0 / self.v
        return self.apply(np.log(self.v), r, -r * r)

    def exp(self):
        e = np.exp(self.v); return self.apply(e, e, e)

    def tanh(self):
        t = np.tanh(self.v); d = 1.0 - t * t
        return self.apply(t, d, -2.0 * t * d)

    def sigmoid(self):
        s = 1.0 / (1.0 + np.exp(-self.v)); d = s * (1.0 - s)
        return self.apply(s, d, d * (1.0 - 2.0 * s))

    def sqrt(self):
        s = np.sqrt(self.v)
        return self.apply(s, 0.5 / s, -0.25 / (s * self.v))

    def clip01(self):
        """jnp.clip(u, 0, 1): derivative 1 strictly inside, 0 outside (ties are measure-zero)."""
        inside = ((self.v > 0.0) & (self.v < 1.0)).astype(np.float64)
        return Bundle(np.clip(self.v, 0.0, 1.0), self.g * inside[:, None], self.l * inside)


def _bsum(bs):
    acc = bs[0]
    for b in bs[1:]:
        acc = acc + b
    return acc


def _spline_bundle(tab, c, x: Bundle, nd: int) -> Bundle:
    """sum_q c_q * basis_q^{(nd)}(x) with d/dx basis^{(nd)} := basis^{(nd+1)} (table semantics, clamped at 3)."""
    f0 = live.table_lookup(tab, nd, x.v)
    f1 = live.table_lookup(tab, nd + 1, x.v)
    f2 = live.table_lookup(tab, nd + 2, x.v)
    return _bsum([c[q] * x.apply(f0[q], f1[q], f2[q]) for q in range(len(c))])


def _conditioner_bundle(net, xs, P, allow_negative):
    """model_factory.py:56-70 on bundles. xs: list of D bundles -> [D][P] bundles (sum-normalised)."""
    nn, _zero = net
    (W1, b1), _, (W2, b2), _, (W3, b3) = [tuple(np.asarray(a, dtype=np.float64) for a in l) if len(l) else () for l in nn]
    D = len(xs)
    m1, m2, m3 = live.made_masks(D, W1.shape[1])
    m3 = np.tile(m3, P)
    W1, W2, W3 = W1 * m1, W2 * m2, W3 * m3
    H = W1.shape[1]
    h1 = [(_bsum([xs[d] * W1[d, j] for d in range(D)]) + b1[j]).tanh() for j in range(H)]
    h2 = [(_bsum([h1[i] * W2[i, j] for i in range(H)]) + b2[j]).tanh() for j in range(H)]
    out = []
    for d in range(D):
        row = []
        for q in range(P):
            col = q * D + d
            o = _bsum([h2[i] * W3[i, col] for i in range(H)]) + b3[col]
            row.append(o if allow_negative else o.sigmoid())
        s = _bsum(row).recip()
        out.append([r * s for r in row])
    return out


def _sumnorm(row):
    s = _bsum(row).recip()
    return [r * s for r in row]


def psi_bundle(m: live.LiveModel, params, x: np.ndarray):
    """Forward-Laplacian restatement of Waveflow.psi for the get_waveflow_model configuration
    (I-BC {0:0}|{0:1}, B-BC {0:0}|{0:0}, no grad_to_zero).  Returns the psi bundle."""
    assert m.prior == "B" and m.bc_i_left == {0: 0} and m.bc_i_right == {0: 1}
    assert m.bc_p_left == {0: 0} and m.bc_p_right == {0: 0} and not m.grad_to_zero
    m = m.cast(np.float64)
    tp, sp = params
    x = np.asarray(x, dtype=np.float64)
    N, D = x.shape
    xs = [Bundle.coord(x, d) for d in range(D)]
    L = float(m.box)
    tol = 1e-7
    # ---- box transform (made.py:156-183 / :118-137)
    ld = Bundle.const(0.0, N, D)
    if m.coord == "mean":
        mean = _bsum(xs) * (1.0 / D)
        l = mean - xs[0]
        w = xs[-1] - xs[0]
        us = []
        space = Bundle.const(2 * L, N, D)
        for i in range(D - 1):
            diff = xs[i + 1] - xs[i]
            us.append(diff / (space + tol))
            ld = ld - (space + tol).log()
            space = space - diff
        den = (2 * L - w) + tol
        us.append((mean + L - l) / den)
        ld = ld - den.log()
    else:
        us = [(xs[0] + L) * (1.0 / (2 * L))]
        ld = ld - np.log(2 * L)
        for i in range(1, D):
            den = (L - xs[i - 1]) + tol
            us.append((xs[i] - xs[i - 1]) / den)
            ld = ld - den.log()
    # ---- IMADE layers (made.py:66-81) + Reverse
    k = m.k_i
    P = m.P_I
    for net in [p for p in tp if len(p)]:
        cs = _conditioner_bundle(net, us, P, allow_negative=False)
        new = []
        for d in range(D):
            row = [c + m.reg for c in cs[d]]
            for i in range(k):                                      # remove_bias, isplines_jax.py:196-200
                row[i + 1] = row[i + 1] * ((i + 1) / k)
                row[P - (i + 2)] = row[P - (i + 2)] * ((i + 1) / k)
            row = _sumnorm(row)
            row[0] = row[0] * 0.0                                   # {0:0}: w[0] = (0 - 0)/I_0(0)
            row[P - 1] = row[P - 1] * 0.0                           # {0:1}: w[-1] = 0
            row = _sumnorm(row)
            y = _spline_bundle(m.tab_I, row, us[d], 0)
            dy = _spline_bundle(m.tab_I, row, us[d], 1)
            ld = ld + (dy + 1e-7).log()
            new.append(y)
        us = new[::-1]
    # ---- prior (wavefunctions.py:58-71, bsplines_jax.py:127-137,173-198)
    PB = m.P_P
    ws = _conditioner_bundle(sp, us, PB, allow_negative=True)
    psi = None
    cons = set(live._constrained(m).tolist())
    for d in range(D):
        row = list(ws[d])
        row[0] = row[0] * 0.0
        row[PB - 1] = row[PB - 1] * 0.0
        nrm = _bsum([r * r for r in row]).sqrt().recip()
        row = [r * nrm for r in row]
        c = [_bsum([row[i] * m.ob_to_b[i, j] for i in range(PB)]) for j in range(PB)]
        nrm = _bsum([r * r for r in c]).sqrt().recip()
        c = [r * nrm for r in c]
        phi = _spline_bundle(m.tab_OB, c, us[d].clip01(), 0)
        if d in cons:
            phi = phi * (1.0 / np.sqrt(2.0))
        psi = phi if psi is None else psi * phi
    return psi * (ld * 0.5).exp()


def local_energy_bundle(m: live.LiveModel, params, x: np.ndarray, protons: np.ndarray):
    """-> dict(psi, lap, hpsi, eloc)  (physics.py:79-93 with eps=0, vqmc.py:198-200)."""
    b = psi_bundle(m, params, x)
    V = live.potential(np.asarray(x, dtype=np.float64), np.asarray(protons, dtype=np.float64))
    hpsi = -0.5 * b.l + V * b.v
    return dict(psi=b.v, grad=b.g, lap=b.l, hpsi=hpsi, eloc=hpsi / (b.v + 1e-8), V=V)


# =========================================================================== torch double-autograd
def autograd_psi_lap(m: live.LiveModel, params, x: np.ndarray, param_grad: bool = False, dtype=np.float64):
    """torch float64 (psi, grad, lap) with the table lookup differentiated as jax does (custom_jvp -> next table).

    param_grad=True keeps the graph and makes every weight a leaf, so that oracle/grad.py can take one more reverse
    pass w.r.t. the parameters (what value_and_grad(loss_fn_efficient) does, vqmc.py:214-221).
    -> (psi, g, lap, leaves) with leaves = {id(numpy leaf): torch leaf}.
    """
    import torch

    tdt = torch.float64 if dtype == np.float64 else torch.float32
    m = m.cast(dtype)
    leaves = {}

    def tt(a):
        if param_grad and id(a) in leaves:
            return leaves[id(a)]
        return torch.as_tensor(np.asarray(a, dtype=dtype))

    if param_grad:
        for net in [p for p in params[0] if len(p)] + [params[1]]:
            for lay in net[0]:
                for a in lay:
                    leaves[id(a)] = torch.tensor(np.asarray(a, dtype=dtype), requires_grad=True)
    tabs = {"I": torch.as_tensor(np.asarray(m.tab_I, dtype=dtype)), "OB": torch.as_tensor(np.asarray(m.tab_OB, dtype=dtype))}

    class Lookup(torch.autograd.Function):
        """value = lerp(table[nd]); d/dx := Lookup(nd+1) (isplines_jax.py:60-66); nd clamps at 3."""
        @staticmethod
        def forward(ctx, xv, name, nd):
            ctx.name, ctx.nd = name, nd
            ctx.save_for_backward(xv)
            tab = tabs[name][min(nd, 3)]
            T = tab.shape[-1]
            xs = xv.detach() * (T - 1)
            il = torch.floor(xs).long().clamp(0, T - 1)
            ir = torch.ceil(xs).long().clamp(0, T - 1)
            yl, yr = tab[:, il], tab[:, ir]
            return yl + (yr - yl) * (T - 1) * (xv.detach() - il.to(xv.dtype) / (T - 1))      # [P, N]

        @staticmethod
        def backward(ctx, gout):
            (xv,) = ctx.saved_tensors
            return (gout * Lookup.apply(xv, ctx.name, ctx.nd + 1)).sum(0), None, None

    def cond(net, u, P, allow_negative):
        nn, _ = net
        (W1, b1), _, (W2, b2), _, (W3, b3) = nn
        D = u.shape[1]
        m1, m2, m3 = live.made_masks(D, W1.shape[1])
        h = torch.tanh(u @ (tt(W1) * tt(m1)) + tt(b1))
        h = torch.tanh(h @ (tt(W2) * tt(m2)) + tt(b2))
        o = h @ (tt(W3) * tt(np.tile(m3, P))) + tt(b3)
        p = o.reshape(-1, P, D).transpose(1, 2)                 # p[n, d, q] = o[n, q*D + d]
        if not allow_negative:
            p = torch.sigmoid(p)
        return p / p.sum(-1, keepdim=True)

    def psi_fn(xx):
        N, D = xx.shape
        L = float(m.box)
        cols = [xx[:, d] for d in range(D)]
        ld = torch.zeros(N, dtype=tdt)
        if m.coord == "mean":
            mean = xx.mean(-1)
            l = mean - cols[0]; w = cols[-1] - cols[0]
            us = []; space = torch.full((N,), 2 * L, dtype=tdt)
            for i in range(D - 1):
                diff = cols[i + 1] - cols[i]
                us.append(diff / (space + 1e-7)); ld = ld - torch.log(space + 1e-7); space = space - diff
            us.append((mean + L - l) / (2 * L - w + 1e-7)); ld = ld - torch.log(2 * L - w + 1e-7)
        else:
            us = [(cols[0] + L) / (2 * L)]
            ld = ld - np.log(2 * L)
            for i in range(1, D):
                us.append((cols[i] - cols[i - 1]) / (L - cols[i - 1] + 1e-7)); ld = ld - torch.log(L - cols[i - 1] + 1e-7)
        u = torch.stack(us, -1)
        k, P = m.k_i, m.P_I
        scale = torch.ones(P, dtype=tdt)
        for i in range(k):
            scale[i + 1] = scale[i + 1] * (i + 1) / k
            scale[P - (i + 2)] = scale[P - (i + 2)] * (i + 1) / k
        mask = torch.ones(P, dtype=tdt); mask[0] = 0; mask[-1] = 0
        for net in [p for p in params[0] if len(p)]:
            c = cond(net, u, P, False) + m.reg
            c = c * scale; c = c / c.sum(-1, keepdim=True)
            c = c * mask; c = c / c.sum(-1, keepdim=True)
            ys = []
            for d in range(D):
                f0 = Lookup.apply(u[:, d], "I", 0); f1 = Lookup.apply(u[:, d], "I", 1)
                ys.append((c[:, d, :].T * f0).sum(0))
                ld = ld + torch.log((c[:, d, :].T * f1).sum(0) + 1e-7)
            u = torch.stack(ys[::-1], -1)
        PB = m.P_P
        w = cond(params[1], u, PB, True)
        maskb = torch.ones(PB, dtype=tdt); maskb[0] = 0; maskb[-1] = 0
        w = w * maskb; w = w / torch.sqrt((w ** 2).sum(-1, keepdim=True))
        c = w @ tt(m.ob_to_b); c = c / torch.sqrt((c ** 2).sum(-1, keepdim=True))
        uc = torch.clamp(u, 0.0, 1.0)
        psi = torch.ones(N, dtype=tdt)
        cons = set(live._constrained(m).tolist())
        for d in range(D):
            phi = (c[:, d, :].T * Lookup.apply(uc[:, d], "OB", 0)).sum(0)
            if d in cons:
                phi = phi / np.sqrt(2.0)
            psi = psi * phi
        return psi * torch.exp(0.5 * ld)

    xx = torch.tensor(np.asarray(x, dtype=dtype), requires_grad=True)
    psi = psi_fn(xx)
    (g,) = torch.autograd.grad(psi.sum(), xx, create_graph=True)
    lap = torch.zeros_like(psi)
    for d in range(xx.shape[1]):
        (g2,) = torch.autograd.grad(g[:, d].sum(), xx, retain_graph=True, create_graph=param_grad)
        lap = lap + g2[:, d]
    return psi, g, lap, leaves


def local_energy_autograd(m: live.LiveModel, params, x: np.ndarray, protons: np.ndarray):
    """Independent check: torch float64, two autograd.grad passes per dimension (what jax.hessian's trace is)."""
    psi, g, lap, _ = autograd_psi_lap(m, params, x)
    V = live.potential(np.asarray(x, dtype=np.float64), np.asarray(protons, dtype=np.float64))
    psi_n, lap_n = psi.detach().numpy(), lap.detach().numpy()
    hpsi = -0.5 * lap_n + V * psi_n
    return dict(psi=psi_n, grad=g.detach().numpy(), lap=lap_n, hpsi=hpsi, eloc=hpsi / (psi_n + 1e-8), V=V)
