"""Array wrapper with jax's functional `.at[...]` updates, a one-level dual number for `grad` of scalar functions, and the
function transformations the reference's spline modules use (jit, vmap, grad, custom_jvp).  Test infrastructure (see README.md)."""
import numpy as _np


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
                out = _np.array(arr, copy=True).view(JArr)
                out[idx] += v
                return out

        return _Upd


class JArr(_np.ndarray):
    """numpy array with `.at[idx].set(v)` / `.add(v)` returning an updated copy, and JAX's default dtype rule on every ufunc
    (operators included): a float64 result becomes float32 ("x64 disabled": int32 / int is a float32 there).  0-d results stay
    arrays, so the rule keeps applying along a scalar computation."""

    @property
    def at(self):
        return _At(self)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [i.view(_np.ndarray) if isinstance(i, JArr) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(o.view(_np.ndarray) if isinstance(o, JArr) else o for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out

        def fix(r):
            r = _np.asarray(r)
            if r.dtype == _np.float64:
                r = r.astype(_np.float32)
            return r.view(JArr)
        return tuple(fix(r) for r in res) if isinstance(res, tuple) else fix(res)


def jarr(x):
    return x.view(JArr) if isinstance(x, _np.ndarray) and not isinstance(x, JArr) else x


class Dual:
    """value + tangent of a scalar (first-order forward mode): what `grad(f, argnums)` of a scalar function needs."""
    __array_ufunc__ = None          # numpy scalars / arrays defer to the reflected operators below

    def __init__(self, v, t):
        self.v, self.t = v, t

    @staticmethod
    def _vt(o):
        return (o.v, o.t) if isinstance(o, Dual) else (o, 0.0)

    def __add__(self, o):
        v, t = Dual._vt(o)
        return Dual(self.v + v, self.t + t)

    __radd__ = __add__

    def __sub__(self, o):
        v, t = Dual._vt(o)
        return Dual(self.v - v, self.t - t)

    def __rsub__(self, o):
        v, t = Dual._vt(o)
        return Dual(v - self.v, t - self.t)

    def __mul__(self, o):
        v, t = Dual._vt(o)
        return Dual(self.v * v, self.t * v + self.v * t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        v, t = Dual._vt(o)
        return Dual(self.v / v, (self.t * v - self.v * t) / (v * v))

    def __neg__(self):
        return Dual(-self.v, -self.t)


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
        a[argnums] = Dual(x, _np.float32(1.0))
        out = fn(*a)
        return out.t if isinstance(out, Dual) else _np.float32(0.0) * x
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
        args = tuple(bound.arguments.values())
        if any(isinstance(a, Dual) for a in args):
            primals = tuple(a.v if isinstance(a, Dual) else a for a in args)
            tangents = tuple(a.t if isinstance(a, Dual) else 0.0 for a in args)
            out, t = self.rule(primals, tangents)
            return Dual(out, t)
        return self.fn(*args)


def hessian(*a, **k):
    raise NotImplementedError("second-order transformations are outside this stand-in")


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
