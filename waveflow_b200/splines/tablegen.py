"""Basis-table generation for the table-interpolated M-/I-/B-spline operators.

Vectorised (over the mesh) float64 Cox-de Boor recursions that follow the
reference's scalar generator operation-for-operation, so the resulting tables
are bit-identical to the ones the reference caches on disk:

* M-spline + derivatives   reference waveflow/splines/splines_np.py:42-62
* I-spline + derivatives   reference waveflow/splines/splines_np.py:79-93
* B-spline + derivatives   reference waveflow/splines/splines_np.py:101-137
* symmetric Gram-Schmidt   reference waveflow/splines/ortho_splines.py:43-161
* knot vectors / file names reference waveflow/splines/isplines_jax.py:91-131,
  msplines_jax.py:72-107, bsplines_jax.py:58-116

Everything here is host-side, run once per (degree, n_knots) and cached as
``.npy`` files that are name-compatible with the reference's cache.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

N_DERIV_TABLES = 4  # reference builds nd = 0..3 (isplines_jax.py:113)


# --------------------------------------------------------------------------- knots
def knots_I(k: int, n_internal_knots: int) -> np.ndarray:
    """isplines_jax.py:91-93 -- end knots repeated k+1 times."""
    ik = np.linspace(0, 1, n_internal_knots)
    ik = np.repeat(ik, ((ik == ik[0]) * (k + 1)).clip(min=1))
    return np.repeat(ik, ((ik == ik[-1]) * (k + 1)).clip(min=1))


def knots_M(k: int, n_internal_knots: int) -> np.ndarray:
    """msplines_jax.py:72-74 -- end knots repeated k times."""
    ik = np.linspace(0, 1, n_internal_knots)
    ik = np.repeat(ik, ((ik == ik[0]) * k).clip(min=1))
    return np.repeat(ik, ((ik == ik[-1]) * k).clip(min=1))


def knots_B(k: int, n_internal_knots: int) -> np.ndarray:
    """bsplines_jax.py:58-60 -- end knots repeated k+1 times."""
    ik = np.linspace(0, 1, n_internal_knots)
    ik = np.repeat(ik, ((ik == ik[0]) * k + 1).clip(min=1))
    return np.repeat(ik, ((ik == ik[-1]) * k + 1).clip(min=1))


# --------------------------------------------------------------------------- M
class _MRec:
    """Memoised M(x, k, i, t, max_k, nd) over a whole mesh (splines_np.py:42-62)."""

    def __init__(self, x: np.ndarray, t: np.ndarray, max_k: int):
        self.x, self.t, self.max_k = x, t, max_k
        self.memo: dict = {}
        self.zero = np.zeros_like(x)

    def __call__(self, k: int, i: int, nd: int) -> np.ndarray:
        key = (k, i, nd)
        if key not in self.memo:
            self.memo[key] = self._eval(k, i, nd)
        return self.memo[key]

    def _eval(self, k, i, nd):
        x, t, max_k = self.x, self.t, self.max_k
        if k == 1:
            inside = ((x >= t[i]) & (x < t[i + 1])) | (
                (i >= len(t) - (max_k + 1)) & (x >= t[i]) & (x <= t[i + 1]))
            if t[i + 1] - t[i] == 0 or nd != 0:
                return self.zero
            return np.where(inside, 1 / (t[i + 1] - t[i]), 0.0)
        if t[i + k] - t[i] == 0:
            return self.zero
        if nd == 0:
            return k * ((x - t[i]) * self(k - 1, i, 0) + (t[i + k] - x) * self(k - 1, i + 1, 0)) / (
                (k - 1) * (t[i + k] - t[i]))
        pref = k / ((k - 1) * (t[i + k] - t[i]))
        if nd == 1:
            return pref * ((x - t[i]) * self(k - 1, i, 1) + (t[i + k] - x) * self(k - 1, i + 1, 1)
                           + self(k - 1, i, 0) - self(k - 1, i + 1, 0))
        return pref * ((x - t[i]) * self(k - 1, i, nd) + (t[i + k] - x) * self(k - 1, i + 1, nd)
                       + nd * (self(k - 1, i, nd - 1) - self(k - 1, i + 1, nd - 1)))


def m_table(k: int, t: np.ndarray, mesh: np.ndarray, nd: int) -> np.ndarray:
    """table[i, m] = M_i^{(nd)}(mesh[m]);  M(x, k, i, t, max_k=k)  (msplines_jax.py:101-102)."""
    rec = _MRec(mesh, t, k)
    return np.stack([rec(k, i, nd) + 0.0 for i in range(len(t) - k)])


# --------------------------------------------------------------------------- I
def i_table(k: int, t: np.ndarray, mesh: np.ndarray, nd: int) -> np.ndarray:
    """table[i, m] = I_i^{(nd)}(mesh[m]);  I(x, k, i, t, max_k=k+1)  (isplines_jax.py:124-125)."""
    n_bases = len(t) - k
    rec = _MRec(mesh, t, k + 1)
    j = np.searchsorted(t, mesh, 'left') - 1          # splines_np.py:83
    j = np.where(mesh == 0.0, k, j)                   # splines_np.py:80-81
    out = np.zeros((n_bases, len(mesh)))
    for i in range(n_bases):
        # splines_np.py:93 -- sum over m = i..j of (t[m+k+1]-t[m]) * M(x,k+1,m) / (k+1); j-i < k there,
        # numpy's .sum() over <8 terms is a plain left-to-right sum.
        acc = np.zeros_like(mesh)
        for m in range(i, min(i + k - 1, len(t) - k - 2) + 1):
            term = (t[m + k + 1] - t[m]) * rec(k + 1, m, nd) / (k + 1)
            acc = acc + np.where(m <= j, term, 0.0)
        one = 1.0 if nd == 0 else 0.0
        val = np.where(i <= j - k, one, acc)
        val = np.where((i > j) | (i == len(t) - (k + 1)), 0.0, val)   # splines_np.py:85-86
        out[i] = val
    return out


# --------------------------------------------------------------------------- B
class _BRec:
    """Memoised B(x, k, i, t, max_k, nd) over a mesh (splines_np.py:101-137)."""

    def __init__(self, x, t, max_k):
        self.x, self.t, self.max_k = x, t, max_k
        self.memo: dict = {}
        self.zero = np.zeros_like(x)

    def __call__(self, k, i, nd):
        key = (k, i, nd)
        if key not in self.memo:
            self.memo[key] = self._eval(k, i, nd)
        return self.memo[key]

    def _eval(self, k, i, nd):
        x, t, max_k = self.x, self.t, self.max_k
        if nd == 0:
            if k == 0:
                inside = ((t[i] <= x) & (x < t[i + 1])) | (
                    (i >= len(t) - (max_k + 2)) & (x >= t[i]) & (x <= t[i + 1]))
                return np.where(inside, 1.0, 0.0)
            c1 = self.zero if t[i + k] == t[i] else (x - t[i]) / (t[i + k] - t[i]) * self(k - 1, i, 0)
            c2 = self.zero if t[i + k + 1] == t[i + 1] else (
                (t[i + k + 1] - x) / (t[i + k + 1] - t[i + 1]) * self(k - 1, i + 1, 0))
            return c1 + c2
        c1 = self.zero if t[i + k] - t[i] == 0 else self(k - 1, i, nd - 1) / (t[i + k] - t[i])
        c2 = self.zero if t[i + k + 1] - t[i + 1] == 0 else self(k - 1, i + 1, nd - 1) / (
            t[i + k + 1] - t[i + 1])
        return k * (c1 - c2)


def b_table(k: int, t: np.ndarray, mesh: np.ndarray, nd: int) -> np.ndarray:
    """table[i, m] = B_i^{(nd)}(mesh[m]);  B(x, k, i, t, max_k=k)  (bsplines_jax.py:93-94)."""
    rec = _BRec(mesh, t, k)
    return np.stack([rec(k, i, nd) + 0.0 for i in range(len(t) - k - 1)])


# --------------------------------------------------------------------------- orthonormalisation
def gram_schmidt_l2r(cols: np.ndarray, gram: np.ndarray) -> np.ndarray:
    """Gram-Schmidt on the columns of `cols`, left to right, driven by their Gram matrix `gram` (= cols.T @ cols, consumed).

    The orthonormal B basis is part of the numerical contract (a different orthonormalisation changes psi for the published
    checkpoint), so this follows the reference's elimination step for step -- same operations in the same order, derived from
    ortho_splines.py:140-161 -- and is pinned bit for bit by tests/test_tablegen_golden.py: at step i the remaining columns and
    the trailing block of the Gram matrix are deflated by column i (outer product, THEN the division by the pivot, THEN the
    subtraction), and column i is scaled by the square root of its current pivot."""
    work = np.copy(cols)
    m = work.shape[1]
    out = np.zeros_like(work, dtype=np.float64)
    for i in range(m):
        pivot = gram[i, i]
        out[:, i] = work[:, i] / np.sqrt(pivot)
        if i + 1 < m:
            row = gram[i, (i + 1):]
            work[:, (i + 1):] -= np.outer(work[:, i], row) / pivot
            gram[(i + 1):, (i + 1):] -= np.outer(row, row) / pivot
    return out


def gram_schmidt_symm(imat: np.ndarray) -> np.ndarray:
    """Symmetrised Gram-Schmidt of the columns of imat (the scheme of ortho_splines.py:43-112, restated with explicit column
    orders; bit-identical tables, tests/test_tablegen_golden.py).

    Two left-to-right passes over interleaved column orders -- one starting at the left end (0, m-1, 1, m-2, ...), one at the
    right end (m-1, 0, m-2, 1, ...) -- give, at the even positions, vectors that are mirror images of each other; each mirror
    pair (v1, v2) with overlap ov is then combined symmetrically (Loewdin for a 2 x 2 block) into the basis functions i and
    m-1-i."""
    mat = np.copy(imat)
    n, m = mat.shape
    if m % 2:
        raise ValueError("only an even number of bases can be orthogonalised (ortho_splines.py:59-63)")
    half = m // 2
    gram = np.dot(mat.T, mat)
    lo, hi = np.arange(half), np.arange(m - 1, m - half - 1, -1)
    order_l = np.stack([lo, hi], axis=1).reshape(-1)          # 0, m-1, 1, m-2, ...
    order_r = np.stack([hi, lo], axis=1).reshape(-1)          # m-1, 0, m-2, 1, ...
    from_left = gram_schmidt_l2r(mat[:, order_l], gram[np.ix_(order_l, order_l)].copy())
    from_right = gram_schmidt_l2r(mat[:, order_r], gram[np.ix_(order_r, order_r)].copy())
    omat = np.zeros((n, m))
    for i in range(half):
        v1, v2 = from_left[:, 2 * i], from_right[:, 2 * i]
        ov = np.dot(v1, v2)
        assert 0 <= ov <= 1
        s1, s2 = 1. / np.sqrt(1 + ov), 1. / np.sqrt(1 - ov)
        a1, a2 = 0.5 * (s1 + s2), 0.5 * (s1 - s2)
        omat[:, i] = a1 * v1 + a2 * v2
        omat[:, m - i - 1] = a2 * v1 + a1 * v2
    return omat * np.sqrt(n)


# --------------------------------------------------------------------------- table sets (+ cache)
def _cache(path: str | None, fn):
    if path is not None and os.path.exists(path):
        return np.load(path)
    arr = fn()
    if path is not None:
        Path(path).parent.mkdir(exist_ok=True, parents=True)
        np.save(path, arr)
    return arr


def build_I_tables(k: int, n_internal_knots: int, n_mesh_points: int = 2000, cache_root: str | None = None):
    """-> (tables float64 [4, n+k, T], knots).  File names as isplines_jax.py:115."""
    t = knots_I(k, n_internal_knots)
    mesh = np.linspace(0, 1, n_mesh_points)
    n_bases = len(t) - k
    tabs = []
    for nd in range(N_DERIV_TABLES):
        p = None if cache_root is None else f"{cache_root}/degree_{k}_niknots_{n_bases}_nmp_{n_mesh_points}_nd_{nd}.npy"
        tabs.append(_cache(p, lambda nd=nd: i_table(k, t, mesh, nd)))
    return np.stack(tabs), t


def build_M_tables(k: int, n_internal_knots: int, n_mesh_points: int = 2000, cache_root: str | None = None):
    """-> (tables float64 [4, n+k-2, T], knots).  File names as msplines_jax.py:93."""
    t = knots_M(k, n_internal_knots)
    mesh = np.linspace(0, 1, n_mesh_points)
    tabs = []
    for nd in range(N_DERIV_TABLES):
        p = None if cache_root is None else f"{cache_root}/degree_{k}_niknots_{len(t) - k}_nmp_{n_mesh_points}_nd_{nd}.npy"
        tabs.append(_cache(p, lambda nd=nd: m_table(k, t, mesh, nd)))
    return np.stack(tabs), t


def build_B_tables(k: int, n_internal_knots: int, n_mesh_points: int = 2000, cache_root: str | None = None):
    """-> dict(b=[4,P,T], ob=[4,P,T], b_to_ob=[P,P], ob_to_b=[P,P], knots).  bsplines_jax.py:74-116."""
    t = knots_B(k, n_internal_knots)
    mesh = np.linspace(0, 1, n_mesh_points)
    nk = len(t)
    stem = None if cache_root is None else f"{cache_root}/degree_{k}_niknots_{nk - k}_nmp_{n_mesh_points}"
    b, ob = [], []
    b_to_ob = ob_to_b = None
    for nd in range(N_DERIV_TABLES):
        pb = None if cache_root is None else f"{cache_root}/b_degree_{k}_niknots_{nk - k}_nmp_{n_mesh_points}_nd_{nd}.npy"
        po = None if cache_root is None else f"{cache_root}/ob_degree_{k}_niknots_{nk - k}_nmp_{n_mesh_points}_nd_{nd}.npy"
        if pb is not None and os.path.exists(pb) and os.path.exists(po):
            b.append(np.load(pb)); ob.append(np.load(po))
            if nd == 0:
                b_to_ob = np.load(f"{stem}_b_to_ob.npy"); ob_to_b = np.load(f"{stem}_ob_to_b.npy")
            continue
        bt = b_table(k, t, mesh, nd)
        if nd == 0:
            obt = gram_schmidt_symm(bt.T).T
            obt = obt / np.sqrt((obt ** 2).sum(-1)[0] / n_mesh_points)
            b_to_ob = obt @ np.linalg.pinv(bt)
            ob_to_b = bt @ np.linalg.pinv(obt)
            if stem is not None:
                Path(cache_root).mkdir(exist_ok=True, parents=True)
                np.save(f"{stem}_b_to_ob.npy", b_to_ob); np.save(f"{stem}_ob_to_b.npy", ob_to_b)
        else:
            obt = b_to_ob @ bt
        if pb is not None:
            np.save(pb, bt); np.save(po, obt)
        b.append(bt); ob.append(obt)
    return dict(b=np.stack(b), ob=np.stack(ob), b_to_ob=b_to_ob, ob_to_b=ob_to_b, knots=t)
