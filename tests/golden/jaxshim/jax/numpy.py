"""jax.numpy stand-in: numpy functions behind dtype canonicalisation (see ../README.md)."""
import builtins

import numpy as _np

from . import _core
from ._core import Array, FLOAT, INT, _plain, asarray, canon_dtype, wrap

ndarray = Array
newaxis = None
pi, inf, nan, e = _np.pi, _np.inf, _np.nan, _np.e
float32, float64 = canon_dtype(_np.float32).type, canon_dtype(_np.float64).type
int32, int64 = canon_dtype(_np.int32).type, canon_dtype(_np.int64).type
uint32, uint8, int8, bool_ = _np.uint32, _np.uint8, _np.int8, _np.bool_
float_, int_ = FLOAT.type, INT.type


def array(x, dtype=None, copy=True):
    return asarray(x, dtype).copy()


def _default_float(dtype):
    return FLOAT if dtype is None else canon_dtype(dtype)


def zeros(shape, dtype=None):
    return wrap(_np.zeros(shape, _default_float(dtype)))


def ones(shape, dtype=None):
    return wrap(_np.ones(shape, _default_float(dtype)))


def full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = canon_dtype(_np.asarray(_plain(fill_value)).dtype)
    return wrap(_np.full(shape, _plain(fill_value), canon_dtype(dtype)))


def arange(*args, dtype=None):
    a = _np.arange(*[_plain(x) for x in args])
    return wrap(a.astype(canon_dtype(dtype)) if dtype is not None else a)


def linspace(start, stop, num=50, endpoint=True, dtype=None):
    """jnp.linspace computes in the result type (float32 by default): start*(1-t) + stop*t with t = iota/(num-1)
    (jax/_src/numpy/lax_numpy.py::_linspace) — not numpy's double evaluation."""
    dt = _default_float(dtype)
    start, stop = _np.asarray(_plain(start), dtype=dt), _np.asarray(_plain(stop), dtype=dt)
    div = (num - 1) if endpoint else num
    if num > 1:
        step = (_np.arange(div, dtype=dt) / dt.type(div)).astype(dt) if False else (_np.arange(num, dtype=dt)[:div] / dt.type(div)).astype(dt)
        out = (start * (dt.type(1) - step) + stop * step).astype(dt)
        if endpoint:
            out = _np.concatenate([out, _np.asarray([stop], dtype=dt)])
    elif num == 1:
        out = _np.asarray([start], dtype=dt)
    else:
        out = _np.zeros((0,), dt)
    return wrap(out)


def take(a, indices, axis=None, mode=None, fill_value=None, **kw):
    a = _np.asarray(_plain(a))
    idx = _np.asarray(_plain(indices)).astype(_np.int64)
    if axis is None:
        a = a.reshape(-1)
        axis = 0
    size = a.shape[axis]
    if mode == "wrap":
        return wrap(_np.take(a, idx % size, axis=axis))
    if mode == "clip":
        return wrap(_np.take(a, _np.clip(idx, 0, size - 1), axis=axis))
    # default ("fill"): negative indices count from the end, out-of-bounds positions return the fill value
    idx = _np.where(idx < 0, idx + size, idx)
    oob = (idx < 0) | (idx >= size)
    out = _np.take(a, _np.clip(idx, 0, builtins.max(size - 1, 0)), axis=axis)
    if oob.any():
        fv = fill_value if fill_value is not None else (_np.nan if out.dtype.kind == "f" else _np.iinfo(out.dtype).min)
        sel = [slice(None)] * out.ndim
        mask = oob.reshape((1,) * axis + oob.shape + (1,) * (out.ndim - axis - oob.ndim))
        out = _np.where(_np.broadcast_to(mask, out.shape), _np.asarray(fv, dtype=out.dtype), out)
    return wrap(out)


def take_along_axis(arr, indices, axis, **kw):
    a = _np.asarray(_plain(arr))
    idx = _np.asarray(_plain(indices)).astype(_np.int64)
    size = a.shape[axis]
    idx = _np.clip(_np.where(idx < 0, idx + size, idx), 0, size - 1)
    return wrap(_np.take_along_axis(a, idx, axis))


def searchsorted(a, v, side="left", sorter=None, method=None):
    return wrap(_np.searchsorted(_plain(a), _plain(v), side=side).astype(INT))


def unique(ar, return_index=False, return_inverse=False, return_counts=False, axis=None, size=None, fill_value=None, **kw):
    res = _np.unique(_plain(ar), return_index=return_index, return_inverse=return_inverse, return_counts=return_counts, axis=axis)
    if size is not None:
        u = res[0] if isinstance(res, tuple) else res
        fv = fill_value if fill_value is not None else (u[0] if u.size else 0)
        u = _np.concatenate([u[:size], _np.full(builtins.max(size - u.size, 0), fv, dtype=u.dtype)])
        res = (u,) + tuple(res[1:]) if isinstance(res, tuple) else u
    return wrap(res)


def where(condition, x=None, y=None, size=None, fill_value=None):
    if x is None and y is None:
        res = _np.nonzero(_np.asarray(_plain(condition)))
        if size is not None:
            fvs = fill_value if isinstance(fill_value, (tuple, list)) else (fill_value,) * len(res)
            res = tuple(_np.concatenate([r[:size], _np.full(builtins.max(size - r.size, 0), 0 if fv is None else fv, dtype=r.dtype)])
                        for r, fv in zip(res, fvs))
        return wrap(tuple(r.astype(INT) for r in res))
    c, x, y = _plain(condition), _plain(x), _plain(y)
    if isinstance(x, float) and isinstance(y, float):
        x = FLOAT.type(x)
    if isinstance(x, int) and isinstance(y, int) and not isinstance(x, bool):
        x = INT.type(x)
    return wrap(_np.where(c, x, y))


def nonzero(a, size=None, fill_value=None):
    return where(a, size=size, fill_value=fill_value)


def argwhere(a, size=None, fill_value=None):
    res = _np.argwhere(_np.asarray(_plain(a)))
    if size is not None:
        pad = _np.full((builtins.max(size - res.shape[0], 0), res.shape[1]), 0 if fill_value is None else fill_value, dtype=res.dtype)
        res = _np.concatenate([res[:size], pad])
    return wrap(res.astype(INT))


def argsort(a, axis=-1, descending=False, stable=True, **kw):
    a = _np.asarray(_plain(a))
    if axis is None:
        a, axis = a.reshape(-1), 0
    if descending:
        # stable descending order: ties keep their original order (jnp.argsort(..., descending=True))
        return wrap(_np.argsort(-a if a.dtype.kind != "b" else ~a, axis=axis, kind="stable").astype(INT))
    return wrap(_np.argsort(a, axis=axis, kind="stable").astype(INT))


def sort(a, axis=-1, **kw):
    return wrap(_np.sort(_plain(a), axis=axis, kind="stable"))


def cumsum(a, axis=None, dtype=None):
    a = _np.asarray(_plain(a))
    dt = canon_dtype(dtype) if dtype is not None else (INT if a.dtype.kind == "b" else a.dtype)
    return wrap(_np.cumsum(a, axis=axis, dtype=dt))


def sum(a, axis=None, dtype=None, keepdims=False, **kw):     # noqa: A001
    a = _np.asarray(_plain(a))
    dt = canon_dtype(dtype) if dtype is not None else (INT if a.dtype.kind in "b" else a.dtype)
    return wrap(_np.sum(a, axis=axis, dtype=dt, keepdims=keepdims))


def clip(a, a_min=None, a_max=None, **kw):
    a_min = kw.get("min", a_min)
    a_max = kw.get("max", a_max)
    a = _np.asarray(_plain(a))
    out = a
    if a_min is not None:
        out = _np.maximum(out, _plain(a_min))
    if a_max is not None:
        out = _np.minimum(out, _plain(a_max))
    return wrap(out)


class _MGrid:
    def __getitem__(self, key):
        return wrap(_np.mgrid[key])


mgrid = _MGrid()


def frompyfunc(*a, **k):
    raise NotImplementedError("jaxshim: jnp.frompyfunc")


def _generic(name):
    fn = getattr(_np, name)

    def call(*args, **kwargs):
        args = tuple(_plain(a) for a in args)
        kwargs = {k: _plain(v) for k, v in kwargs.items()}
        if kwargs.get("dtype") is not None:
            kwargs["dtype"] = canon_dtype(kwargs["dtype"])
        with _np.errstate(all="ignore"):
            return wrap(fn(*args, **kwargs))
    call.__name__ = name
    return call


def __getattr__(name):
    if hasattr(_np, name):
        obj = getattr(_np, name)
        if callable(obj) and not isinstance(obj, type):
            f = _generic(name)
            globals()[name] = f
            return f
        return obj
    raise AttributeError("jaxshim.jax.numpy has no attribute %r" % name)
