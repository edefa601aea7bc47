"""ORACLE (test infrastructure): a *vectorised* CPU port of the Waveflow local energy, used as the timed CPU baseline.

Same arithmetic as oracle/laplacian.py::local_energy_bundle (forward-mode propagation of value / gradient / Laplacian,
table derivatives = next table), but on whole [N, G, ...] torch-CPU tensors so that it runs on all host cores through
MKL/OpenMP -- the fairest stand-in available for the reference's jitted XLA-CPU path (JAX is not installable in this
image).  Checked against the scalar-bundle oracle in tests/test_oracle_golden.py.

Layout: a "bundle" is a tensor [N, G, *S] with G = D + 2 components (value, D gradient entries, Laplacian).
"""
from __future__ import annotations

import numpy as np
import torch

from . import live


def _unary(a, f0, f1, f2, D):
    out = torch.empty_like(a)
    out[:, 0] = f0
    out[:, 1:D + 1] = f1.unsqueeze(1) * a[:, 1:D + 1]
    out[:, D + 1] = f1 * a[:, D + 1] + f2 * (a[:, 1:D + 1] ** 2).sum(1)
    return out


def _mul(a, b, D):
    out = torch.empty_like(a + b)
    av, bv = a[:, 0], b[:, 0]
    out[:, 0] = av * bv
    out[:, 1:D + 1] = av.unsqueeze(1) * b[:, 1:D + 1] + bv.unsqueeze(1) * a[:, 1:D + 1]
    out[:, D + 1] = av * b[:, D + 1] + bv * a[:, D + 1] + 2 * (a[:, 1:D + 1] * b[:, 1:D + 1]).sum(1)
    return out


def _recip(a, D):
    r = 1.0 / a[:, 0]
    return _unary(a, r, -r * r, 2 * r * r * r, D)


def _log(a, D):
    r = 1.0 / a[:, 0]
    return _unary(a, torch.log(a[:, 0]), r, -r * r, D)


def _tanh(a, D):
    t = torch.tanh(a[:, 0]); d = 1 - t * t
    return _unary(a, t, d, -2 * t * d, D)


def _sigmoid(a, D):
    s = torch.sigmoid(a[:, 0]); d = s * (1 - s)
    return _unary(a, s, d, d * (1 - 2 * s), D)


def _const_add(a, c):
    a = a.clone(); a[:, 0] += c; return a


def _lookup(tab, nd, xv):
    """[P, N] interpolated basis values (isplines_jax.py:45-56), tab: torch [4, P, T]."""
    T = tab.shape[-1]
    xs = xv * (T - 1)
    il = torch.floor(xs).long(); ir = torch.ceil(xs).long()
    dx = xv - il.to(xv.dtype) / (T - 1)
    il = torch.where(il < 0, il + T, il).clamp(0, T - 1); ir = torch.where(ir < 0, ir + T, ir).clamp(0, T - 1)
    t = tab[min(nd, 3)]
    yl, yr = t[:, il], t[:, ir]
    return yl + (yr - yl) * (T - 1) * dx


def _spline(tab, c, x, nd, D):
    """sum_q c[..., q] (x) basis_q^{(nd)}(x).  c: [N, G, P], x: [N, G] -> [N, G]."""
    f0, f1, f2 = (_lookup(tab, nd + k, x[:, 0]).T for k in range(3))        # [N, P]
    N, G, P = c.shape
    bas = torch.empty_like(c)
    bas[:, 0] = f0
    bas[:, 1:D + 1] = f1.unsqueeze(1) * x[:, 1:D + 1].unsqueeze(-1)
    bas[:, D + 1] = f1 * x[:, D + 1].unsqueeze(-1) + f2 * (x[:, 1:D + 1] ** 2).sum(1).unsqueeze(-1)
    return _mul(c, bas, D).sum(-1)


def _conditioner(net, u, P, allow_negative, D, masks):
    """u: [N, G, D] -> normalised coefficients [N, G, D, P]  (model_factory.py:56-70)."""
    W1, b1, W2, b2, W3, b3 = net
    m1, m2, m3 = masks
    h = u @ (W1 * m1); h[:, 0] += b1; h = _tanh(h, D)
    h = h @ (W2 * m2); h[:, 0] += b2; h = _tanh(h, D)
    o = h @ (W3 * m3.repeat(1, P)); o[:, 0] += b3
    N, G, _ = o.shape
    o = o.reshape(N, G, P, D).transpose(2, 3).contiguous()                  # [N, G, D, P]
    if not allow_negative:
        o = _sigmoid(o, D)
    return _mul(o, _recip(o.sum(-1, keepdim=True), D), D)


class FastLocalEnergy:
    """Pre-converted model (tables, masks, weights) + __call__(x) -> dict(psi, hpsi, eloc)."""

    def __init__(self, m: live.LiveModel, params, protons, dtype=torch.float32, threads: int | None = None):
        assert m.prior == "B" and m.bc_i_left == {0: 0} and m.bc_i_right == {0: 1} and m.box is not None
        if threads:
            torch.set_num_threads(threads)
        self.m, self.dtype, self.D = m, dtype, m.D
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
        self.tab_I, self.tab_OB, self.ob_to_b = t(m.tab_I), t(m.tab_OB), t(m.ob_to_b)
        self.masks = [t(a) for a in live.made_masks(m.D)]
        conv = lambda net: tuple(t(a) for lay in net[0] if len(lay) for a in lay)
        self.nets = [conv(p) for p in params[0] if len(p)]
        self.prior = conv(params[1])
        self.protons = np.asarray(protons, dtype=np.float64)
        k, P = m.k_i, m.P_I
        sc = torch.ones(P, dtype=dtype)
        for i in range(k):
            sc[i + 1] = sc[i + 1] * (i + 1) / k
            sc[P - (i + 2)] = sc[P - (i + 2)] * (i + 1) / k
        sc[0] = 0; sc[P - 1] = 0
        self.wq = sc
        mb = torch.ones(m.P_P, dtype=dtype); mb[0] = 0; mb[-1] = 0
        self.mb = mb

    @torch.no_grad()
    def __call__(self, x: np.ndarray):
        m, D, dt = self.m, self.D, self.dtype
        x = torch.as_tensor(np.asarray(x), dtype=dt)
        N = x.shape[0]
        G = D + 2
        X = torch.zeros(N, G, D, dtype=dt)
        X[:, 0] = x
        for d in range(D):
            X[:, 1 + d, d] = 1.0
        L, tol = float(m.box), 1e-7
        ld = torch.zeros(N, G, dtype=dt)
        U = torch.zeros(N, G, D, dtype=dt)
        if m.coord == "mean":
            mean = X.sum(-1) / D
            l = mean - X[..., 0]; w = X[..., -1] - X[..., 0]
            space = torch.zeros(N, G, dtype=dt); space[:, 0] = 2 * L
            for i in range(D - 1):
                diff = X[..., i + 1] - X[..., i]
                den = _const_add(space, tol)
                U[..., i] = _mul(diff, _recip(den, D), D)
                ld = ld - _log(den, D)
                space = space - diff
            den = _const_add(-w, 2 * L + tol)
            U[..., -1] = _mul(_const_add(mean - l, L), _recip(den, D), D)
            ld = ld - _log(den, D)
        else:
            U[..., 0] = _const_add(X[..., 0], L) / (2 * L)
            ld[:, 0] -= float(np.log(2 * L))
            for i in range(1, D):
                den = _const_add(-X[..., i - 1], L + tol)
                U[..., i] = _mul(X[..., i] - X[..., i - 1], _recip(den, D), D)
                ld = ld - _log(den, D)
        for net in self.nets:
            c = _conditioner(net, U, m.P_I, False, D, self.masks)           # [N, G, D, P]
            c = _const_add(c, m.reg) * self.wq
            c = _mul(c, _recip(c.sum(-1, keepdim=True), D), D)
            Y = torch.empty_like(U)
            for d in range(D):
                Y[..., d] = _spline(self.tab_I, c[:, :, d], U[..., d], 0, D)
                ld = ld + _log(_const_add(_spline(self.tab_I, c[:, :, d], U[..., d], 1, D), 1e-7), D)
            U = Y.flip(-1)
        wgt = _conditioner(self.prior, U, m.P_P, True, D, self.masks) * self.mb
        nrm = _unary_rsqrt(_mul(wgt, wgt, D).sum(-1, keepdim=True), D)
        wgt = _mul(wgt, nrm, D)
        c = wgt @ self.ob_to_b
        c = _mul(c, _unary_rsqrt(_mul(c, c, D).sum(-1, keepdim=True), D), D)
        psi = None
        cons = set(live._constrained(m).tolist())
        for d in range(D):
            ud = U[..., d]
            inside = ((ud[:, 0] > 0) & (ud[:, 0] < 1)).to(dt)
            uc = ud * inside.unsqueeze(1)
            uc[:, 0] = ud[:, 0].clamp(0, 1)
            phi = _spline(self.tab_OB, c[:, :, d], uc, 0, D)
            if d in cons:
                phi = phi / float(np.sqrt(2.0))
            psi = phi if psi is None else _mul(psi, phi, D)
        half = 0.5 * ld
        e = torch.exp(half[:, 0])
        psi = _mul(psi, _unary(half, e, e, e, D), D)
        V = torch.as_tensor(live.potential(x.numpy().astype(np.float64), self.protons), dtype=dt)
        hpsi = -0.5 * psi[:, D + 1] + V * psi[:, 0]
        return dict(psi=psi[:, 0].numpy(), hpsi=hpsi.numpy(), eloc=(hpsi / (psi[:, 0] + 1e-8)).numpy(),
                    lap=psi[:, D + 1].numpy(), grad=psi[:, 1:D + 1].numpy())


def _unary_rsqrt(a, D):
    s = torch.rsqrt(a[:, 0])
    return _unary(a, s, -0.5 * s / a[:, 0], 0.75 * s / (a[:, 0] ** 2), D)
