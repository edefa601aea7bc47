"""Core of the numpy stand-in for jax (test infrastructure, see ../README.md): an array class with jax's dtype rule, functional
`.at[...]` updates, immutable in-place operators and clamped out-of-range indexing; level-tagged dual numbers for forward-mode
differentiation (nested for second order); and the function transformations the reference's files use -- jit, vmap, grad,
hessian, custom_jvp (the registered rule is applied whenever a dual number reaches the function)."""
import math as _math
import os as _os

import numpy as _np

X64 = _os.environ.get("JAX_SHIM_X64", "0") == "1"        # float64 everywhere (jax_enable_x64): used for the local-energy vectors


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        arr = self._arr

        class _Upd:
            @staticmethod
            def set(v):
                out = _np.array(arr, copy=True).view(JArr)
                out[idx] = v
                return out

            @staticmethod
            def add(v):
                out = _np.array(arr, copy=True)
                out[idx] = out[idx] + v
                return out.view(JArr)

        return _Upd


class JArr(_np.ndarray):
    """numpy array with `.at[idx].set(v)` / `.add(v)` returning an updated copy, and JAX's default dtype rule on every ufunc
    (operators included): a float64 result becomes float32 ("x64 disabled": int32 / int is a float32 there).  0-d results stay
    arrays, so the rule keeps applying along a scalar computation."""

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        """jax's out-of-bounds rule for retrieval: integer indices are CLAMPED into the axis (negative ones wrap first) instead of
        raising -- `tables[n_derivative + 1]` with n_derivative = 3 reads table 3, a mesh index above T - 1 reads
        the last node (Q3)."""
        parts = idx if isinstance(idx, tuple) else (idx,)
        if not any(p is Ellipsis for p in parts):
            fixed, axis = [], 0
            for p_ in parts:
                if p_ is None:
                    fixed.append(p_)
                    continue
                n = self.shape[axis] if axis < self.ndim else 1
                if isinstance(p_, _np.ndarray) and p_.ndim == 0 and p_.dtype.kind in "iu":
                    p_ = int(p_)
                if isinstance(p_, (int, _np.integer)) and not isinstance(p_, (bool, _np.bool_)):
                    q = int(p_) + n if p_ < 0 else int(p_)
                    p_ = min(max(q, 0), n - 1)
                elif isinstance(p_, _np.ndarray) and p_.dtype.kind in "iu":
                    q = p_.view(_np.ndarray)
                    p_ = _np.clip(_np.where(q < 0, q + n, q), 0, n - 1)
                fixed.append(p_)
                axis += 1
            idx = tuple(fixed) if isinstance(idx, tuple) else fixed[0]
        return super().__getitem__(idx)

    def __iter__(self):                      # by length (the sequence protocol would wait for an IndexError that clamping never raises)
        for i in range(len(self)):
            yield super().__getitem__(i)

    # jax arrays are immutable: `a += b` rebinds the name to a new array (and may change its dtype)
    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [i.view(_np.ndarray) if isinstance(i, JArr) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(o.view(_np.ndarray) if isinstance(o, JArr) else o for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out

        def fix(r):
            r = _np.asarray(r)
            if r.dtype == _np.float64 and not X64:
                r = r.astype(_np.float32)
            return r.view(JArr)
        return tuple(fix(r) for r in res) if isinstance(res, tuple) else fix(res)


def jarr(x):
    return x.view(JArr) if isinstance(x, _np.ndarray) and not isinstance(x, JArr) else x


def _unbox(a):
    """0-d object arrays (what vmap hands out) -> the element itself."""
    if isinstance(a, _np.ndarray) and a.dtype == object and a.ndim == 0:
        return a.item()
    return a


def _f(name, v):
    """elementary function on a plain number or a (nested) Dual"""
    return getattr(v, name)() if isinstance(v, Dual) else getattr(_math, name)(float(v))


_TAG = [0]


def new_tag():
    _TAG[0] += 1
    return _TAG[0]


class Dual:
    """value + tangent along ONE direction (forward mode), with a TAG that names the differentiation level it belongs to.
    Components may be Duals of lower tags, which nests the modes (forward over forward: second derivatives).  In a binary
    operation the operand of the lower tag is a constant with respect to the higher one (no perturbation confusion: the
    `grad` inside the reference's apply_fun_vec_grad runs under the outer levels of `hessian`).  Arithmetic and the elementary
    functions are generic over the nesting; a `custom_jvp` function called on a Dual applies the rule the reference registered."""
    __array_ufunc__ = None          # numpy scalars / arrays defer to the reflected operators below
    __array_priority__ = 1000

    def __init__(self, v, t, tag):
        self.v, self.t, self.tag = v, t, tag

    def _vt(self, o):
        """(value, tangent) of the other operand AT THIS LEVEL, or None if the other operand lives on a higher level."""
        o = _unbox(o)
        if isinstance(o, Dual):
            if o.tag == self.tag:
                return o.v, o.t
            if o.tag > self.tag:
                return None
        return o, 0.0

    def __add__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__radd__(self)
        return Dual(self.v + vt[0], self.t + vt[1], self.tag)

    __radd__ = __add__

    def __sub__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__rsub__(self)
        return Dual(self.v - vt[0], self.t - vt[1], self.tag)

    def __rsub__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__sub__(self)
        return Dual(vt[0] - self.v, vt[1] - self.t, self.tag)

    def __mul__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__rmul__(self)
        v, t = vt
        return Dual(self.v * v, self.t * v + self.v * t, self.tag)

    __rmul__ = __mul__

    def __truediv__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__rtruediv__(self)
        v, t = vt
        return Dual(self.v / v, (self.t * v - self.v * t) / (v * v), self.tag)

    def __rtruediv__(self, o):
        vt = self._vt(o)
        if vt is None:
            return _unbox(o).__truediv__(self)
        v, t = vt
        return Dual(v / self.v, (t * self.v - v * self.t) / (self.v * self.v), self.tag)

    def __neg__(self):
        return Dual(-self.v, -self.t, self.tag)

    def __pos__(self):
        return self

    def __abs__(self):
        return self if self >= 0 else -self

    def __pow__(self, n):
        if isinstance(n, (int, _np.integer)) and n >= 0:
            out = 1.0
            for _ in range(int(n)):
                out = out * self
            return out
        if float(n) == 0.5:
            return self.sqrt()
        return (self.log() * float(n)).exp()

    # comparisons act on the primal value (the branch taken), as in jax
    def _p(self):
        v = self.v
        while isinstance(v, Dual):
            v = v.v
        return v

    @staticmethod
    def _pv(o):
        o = _unbox(o)
        return o._p() if isinstance(o, Dual) else o

    def __lt__(self, o): return self._p() < Dual._pv(o)
    def __le__(self, o): return self._p() <= Dual._pv(o)
    def __gt__(self, o): return self._p() > Dual._pv(o)
    def __ge__(self, o): return self._p() >= Dual._pv(o)
    def __eq__(self, o): return self._p() == Dual._pv(o)
    def __ne__(self, o): return self._p() != Dual._pv(o)
    __hash__ = None

    # numpy ufuncs on object arrays call the method of the same name
    def exp(self):
        e = _f("exp", self.v)
        return Dual(e, e * self.t, self.tag)

    def log(self):
        return Dual(_f("log", self.v), self.t / self.v, self.tag)

    def sqrt(self):
        r = _f("sqrt", self.v)
        return Dual(r, self.t / (2.0 * r), self.tag)

    def tanh(self):
        th = _f("tanh", self.v)
        return Dual(th, (1.0 - th * th) * self.t, self.tag)

    def conjugate(self):
        return self


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        args = [jarr(_np.asarray(a)) if ax is not None else a for a, ax in zip(args, axes)]
        n = next(a.shape[0] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a[i, ...] if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]   # 0-d views, not scalars
        if isinstance(outs[0], tuple):
            return tuple(jarr(_np.stack([_np.asarray(o[j]) for o in outs])) for j in range(len(outs[0])))
        return jarr(_np.stack([_np.asarray(o) for o in outs]))
    return mapped


def grad(fn, argnums=0):
    """d fn / d args[argnums] for a scalar argument, by first-order forward mode through custom_jvp rules and arithmetic."""
    def g(*args):
        a = list(args)
        x = a[argnums]
        x = _unbox(x)
        tag = new_tag()
        a[argnums] = Dual(x, 1.0 if X64 else _np.float32(1.0), tag)
        out = _unbox(fn(*a))
        return out.t if isinstance(out, Dual) and out.tag == tag else (0.0 if X64 else _np.float32(0.0)) * x
    return g


class custom_jvp:
    """jax.custom_jvp: the function itself on plain inputs; on a Dual input the registered rule (primals, tangents) -> (out, t)."""

    def __init__(self, fn):
        self.fn, self.rule = fn, None
        self.__name__ = getattr(fn, "__name__", "custom_jvp")

    def defjvp(self, rule):
        self.rule = rule
        return rule

    def __call__(self, *args, **kwargs):
        import inspect
        bound = inspect.signature(self.fn).bind(*args, **kwargs)      # as jax does: defaults and keywords become positional
        bound.apply_defaults()
        args = tuple(_unbox(a) for a in bound.arguments.values())
        if any(isinstance(a, _np.ndarray) and a.dtype == object for a in args):
            return self._call_arrays(args)
        tags = [a.tag for a in args if isinstance(a, Dual)]
        if tags:
            tag = max(tags)                              # differentiate the outermost level; lower levels ride along as values
            top = lambda a: isinstance(a, Dual) and a.tag == tag
            primals = tuple(a.v if top(a) else a for a in args)
            tangents = tuple(a.t if top(a) else 0.0 for a in args)
            out, t = self.rule(primals, tangents)
            return Dual(out, t, tag)
        return self.fn(*args)


def _custom_jvp_call_arrays(self, args):
    """array arguments holding Duals (the loss estimator of vqmc.py): the rule sees primal and tangent ARRAYS of the top level"""
    flat = [e for a in args if isinstance(a, _np.ndarray) and a.dtype == object for e in a.reshape(-1) if isinstance(e, Dual)]
    if not flat:
        return self.fn(*args)
    tag = max(e.tag for e in flat)
    top = lambda e: isinstance(e, Dual) and e.tag == tag

    def split(a):
        if isinstance(a, _np.ndarray) and a.dtype == object:
            p = _np.empty(a.shape, dtype=object)
            t = _np.empty(a.shape, dtype=object)
            for idx in _np.ndindex(a.shape):
                e = a[idx]
                p[idx], t[idx] = (e.v, e.t) if top(e) else (e, 0.0)
            return p.view(JArr), t.view(JArr)
        return a, (_np.zeros_like(a) if isinstance(a, _np.ndarray) else 0.0)

    pt = [split(a) for a in args]
    out, tan = self.rule(tuple(p for p, _ in pt), tuple(t for _, t in pt))
    out, tan = _np.asarray(out, dtype=object), _np.asarray(tan, dtype=object)
    res = _np.empty(out.shape, dtype=object)
    for idx in _np.ndindex(out.shape):
        res[idx] = Dual(out[idx], tan[idx], tag)
    return res.view(JArr)


custom_jvp._call_arrays = _custom_jvp_call_arrays


def value_and_grad(*a, **k):
    raise NotImplementedError("reverse mode is outside this stand-in (directional derivatives: seed the parameters with Duals)")


def tree_map(f, tree):
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, t) for t in tree)
    return f(tree)


def hessian(fn, argnums=0):
    """jax.hessian for a vector argument: H[..., i, j] = d2 out / dx_i dx_j by forward over forward mode (nested Duals): the
    outer level carries e_j, the inner one e_i; custom_jvp rules are applied at both levels, as jacfwd(jacfwd) would."""
    def h(*args):
        x = _np.asarray(args[argnums])
        D = x.shape[0]
        t1, t2 = new_tag(), new_tag()
        rows = []
        for i in range(D):
            cols = []
            for j in range(D):
                xo = _np.empty(D, dtype=object)
                for k in range(D):
                    xo[k] = Dual(Dual(x[k].item(), 1.0 if k == i else 0.0, t1), Dual(1.0 if k == j else 0.0, 0.0, t1), t2)
                a = list(args)
                a[argnums] = xo.view(JArr)
                out = _np.asarray(fn(*a), dtype=object).reshape(-1)
                cols.append([o.t.t if isinstance(o, Dual) and o.tag == t2 and isinstance(o.t, Dual) and o.t.tag == t1 else 0.0
                             for o in out])
            rows.append(cols)
        plain = not any(isinstance(e, Dual) for r in rows for c in r for e in c)
        H = _np.array(rows, dtype=_np.float64 if plain else object)      # [i, j, out] (Duals of lower levels stay objects)
        return jarr(_np.moveaxis(H, -1, 0))                               # [out, i, j]
    return h


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
