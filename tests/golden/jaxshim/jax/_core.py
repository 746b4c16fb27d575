"""Core of the numpy stand-in for jax (see ../README.md): the array type, dtype canonicalisation, indexing semantics."""
import os

import numpy as np

X64 = os.environ.get("JAXSHIM_X64", "0") == "1"     # jax_enable_x64


def canon_dtype(dt):
    """jax's dtype canonicalisation: without x64 every 64-bit type becomes its 32-bit sibling; with JAXSHIM_X64 the
    shim goes one step further than jax and also widens explicit 32-bit requests, so that the WHOLE program runs in
    double (that mode exists for finite differences and for a rounding-free structural comparison)."""
    if dt is None:
        return None
    dt = np.dtype(dt)
    if X64:
        if dt == np.float32:
            return np.dtype(np.float64)
        if dt == np.int32:
            return np.dtype(np.int64)
        return dt
    if dt == np.float64:
        return np.dtype(np.float32)
    if dt == np.int64:
        return np.dtype(np.int32)
    if dt == np.uint64:
        return np.dtype(np.uint32)
    if dt == np.complex128:
        return np.dtype(np.complex64)
    return dt


FLOAT = canon_dtype(np.float64)
INT = canon_dtype(np.int64)


def _plain(x):
    """Operand -> plain numpy (canonical dtype); Python scalars stay weak."""
    if isinstance(x, np.ndarray):
        a = x.view(np.ndarray)
        c = canon_dtype(a.dtype)
        return a if c == a.dtype else a.astype(c)
    if isinstance(x, np.generic):
        return np.asarray(x).astype(canon_dtype(x.dtype))
    if isinstance(x, (list, tuple)) and any(isinstance(e, (np.ndarray, np.generic)) for e in x):
        return type(x)(_plain(e) for e in x)
    return x


def wrap(x):
    """Result -> Array (canonical dtype), through tuples / lists."""
    if isinstance(x, (tuple, list)):
        return type(x)(wrap(e) for e in x)
    if isinstance(x, (np.ndarray, np.generic)):
        a = np.asarray(x)
        c = canon_dtype(a.dtype)
        if c != a.dtype:
            a = a.astype(c)
        return a.view(Array)
    return x


def _float_divmod(a, b):
    """jnp.floor_divide / jnp.remainder for floats (jax/_src/numpy/ufuncs.py::_float_divmod)."""
    mod = np.fmod(a, b)
    div = (a - mod) / b
    ind = (mod != 0) & (np.sign(b) != np.sign(mod))
    mod = np.where(ind, mod + b, mod)
    div = np.where(ind, div - 1, div)
    # lax.round(x): half away from zero
    rdiv = np.sign(div) * np.floor(np.abs(div) + 0.5)
    rdiv = rdiv.astype(div.dtype) if isinstance(rdiv, np.ndarray) else div.dtype.type(rdiv)
    return rdiv, mod


def _weak_promote(ins):
    """A Python float next to integer arrays computes in the default float type (weak-type promotion), not in double."""
    has_pyfloat = any(isinstance(i, float) for i in ins)
    arrays = [i for i in ins if isinstance(i, np.ndarray)]
    if has_pyfloat and arrays and all(a.dtype.kind in "iub" for a in arrays):
        return tuple(i.astype(FLOAT) if isinstance(i, np.ndarray) else i for i in ins)
    return ins


def normalize_index(idx, shape, clamp):
    """Python-style negative wrap for integer-array indices, then clamp (gather) — returns the index and, for scatters
    (clamp=False), a mask of the positions whose index is in bounds (jax drops the others)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    idx = tuple(np.asarray(i) if isinstance(i, list) else i for i in idx)
    n_real = sum(1 for i in idx if i is not None and i is not Ellipsis and not (isinstance(i, np.ndarray) and i.dtype == bool))
    n_real += sum(i.ndim for i in idx if isinstance(i, np.ndarray) and i.dtype == bool)
    out, axis, ok = [], 0, None
    for i in idx:
        if i is None:
            out.append(i)
            continue
        if i is Ellipsis:
            axis += len(shape) - n_real
            out.append(i)
            continue
        if isinstance(i, np.ndarray) and i.dtype == bool:
            out.append(i.view(np.ndarray))
            axis += i.ndim
            continue
        if isinstance(i, (np.ndarray, np.generic)) and np.asarray(i).dtype.kind in "iu":
            a = np.asarray(i).view(np.ndarray).astype(np.int64)
            size = shape[axis]
            a = np.where(a < 0, a + size, a)
            inb = (a >= 0) & (a < size)
            if clamp:
                a = np.clip(a, 0, max(size - 1, 0))
            else:
                ok = inb if ok is None else (ok & inb)     # broadcast together like the index arrays themselves
                a = np.where(inb, a, size)                   # the padding slot of _AtRef._scatter / .get
            out.append(a)
            axis += 1
            continue
        out.append(i)
        axis += 1
    return tuple(out), ok


class _AtRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _scatter(self, vals, op):
        base = np.array(self.arr.view(np.ndarray), copy=True)
        idx, ok = normalize_index(self.idx, base.shape, clamp=False)
        vals = _plain(vals)
        if not isinstance(vals, np.ndarray):
            vals = np.asarray(vals, dtype=base.dtype)
        if ok is None:                                   # slices / integers / boolean masks only
            if op == "set":
                base[idx] = vals
            elif op == "add":
                base[idx] += vals.astype(base.dtype)
            else:
                base[idx] *= vals.astype(base.dtype)
            return wrap(base)
        # advanced indices: flat target position of every update, -1 where an index is out of bounds (jax drops those).
        # Every axis is padded by one slot that holds -1; out-of-bounds indices were routed to that slot.
        posarr = np.full(tuple(s + 1 for s in base.shape), -1, dtype=np.int64)
        posarr[tuple(slice(0, s) for s in base.shape)] = np.arange(base.size, dtype=np.int64).reshape(base.shape)
        pos = posarr[_padded_index(idx, base.shape)]
        v = np.broadcast_to(vals.astype(base.dtype), pos.shape)
        keep = pos >= 0
        flat = base.reshape(-1)
        if op == "set":
            flat[pos[keep]] = v[keep]
        elif op == "add":
            # XLA leaves the order of a float scatter-add unspecified: take the exactly rounded sum (accumulate in double)
            if flat.dtype == np.float32:
                acc = flat.astype(np.float64)
                np.add.at(acc, pos[keep], v[keep].astype(np.float64))
                flat = acc.astype(np.float32)
            else:
                np.add.at(flat, pos[keep], v[keep])
        else:
            np.multiply.at(flat, pos[keep], v[keep])
        return wrap(flat.reshape(base.shape))

    def set(self, vals, **kw):
        return self._scatter(vals, "set")

    def add(self, vals, **kw):
        return self._scatter(vals, "add")

    def subtract(self, vals, **kw):
        return self._scatter(-np.asarray(_plain(vals)), "add")

    def multiply(self, vals, **kw):
        return self._scatter(vals, "mul")

    def get(self, mode=None, fill_value=None, **kw):
        base = self.arr.view(np.ndarray)
        if (mode is None and fill_value is None) or mode == "clip":
            idx, _ = normalize_index(self.idx, base.shape, clamp=True)
            return wrap(base[idx])
        idx, ok = normalize_index(self.idx, base.shape, clamp=False)
        if ok is None:
            return wrap(base[idx])
        fv = fill_value if fill_value is not None else (np.nan if base.dtype.kind == "f" else np.iinfo(base.dtype).min)
        padded = np.pad(base, [(0, 1)] * base.ndim, constant_values=fv)
        return wrap(padded[_padded_index(idx, base.shape)])


def _padded_index(idx, shape):
    """The index tuple re-targeted at an array padded by one slot per axis: the Ellipsis is expanded and every slice is
    bounded by the ORIGINAL axis length, so only routed out-of-bounds integers can reach the padding."""
    n_real = sum(1 for i in idx if i is not None and i is not Ellipsis)
    full = []
    for i in idx:
        if i is Ellipsis:
            full.extend([slice(None)] * (len(shape) - n_real))
        else:
            full.append(i)
    full.extend([slice(None)] * (len(shape) - sum(1 for i in full if i is not None)))
    out, axis = [], 0
    for i in full:
        if i is None:
            out.append(i)
            continue
        if isinstance(i, slice):
            out.append(slice(*i.indices(shape[axis])) if (i.step or 1) > 0 else i)
        else:
            out.append(i)
        axis += 1
    return tuple(out)


class _AtHelper:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


def _canon_method(name):
    """ndarray methods that do not dispatch through __array_ufunc__ and would hand back 64-bit results."""
    def method(self, *a, **k):
        return wrap(getattr(self.view(np.ndarray), name)(*a, **k))
    method.__name__ = name
    return method


class Array(np.ndarray):
    """jax.Array stand-in: immutable-style `.at[]` updates, clamped gathers, canonical dtypes."""
    __array_priority__ = 100.0

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = _weak_promote(tuple(_plain(i) for i in inputs))
        if "dtype" in kwargs and kwargs["dtype"] is not None:
            kwargs["dtype"] = canon_dtype(kwargs["dtype"])
        if method == "__call__" and ufunc in (np.floor_divide, np.remainder):
            a, b = ins
            dts = [np.asarray(i).dtype for i in ins if isinstance(i, (np.ndarray, np.generic))]
            if any(d.kind == "f" for d in dts) or any(isinstance(i, float) for i in ins):
                ft = np.result_type(*[d for d in dts if d.kind == "f"]) if any(d.kind == "f" for d in dts) else FLOAT
                a = np.asarray(a, dtype=ft)
                b = np.asarray(b, dtype=ft)
                d, m = _float_divmod(a, b)
                return wrap(np.asarray(d if ufunc is np.floor_divide else m, dtype=ft))
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
        return wrap(getattr(ufunc, method)(*ins, **kwargs))

    def __array_function__(self, func, types, args, kwargs):
        # numpy.<function>(jax array): numpy converts the operand (np.asarray) and applies ITS OWN semantics, returning
        # numpy arrays — e.g. np.unique(event[mask]) in optimize/simulate.py:139 yields hashable numpy integers
        def plain(a):
            if isinstance(a, np.ndarray):
                return a.view(np.ndarray)
            if isinstance(a, (list, tuple)):
                return type(a)(plain(e) for e in a)
            return a
        return func(*plain(args), **{k: plain(v) for k, v in kwargs.items()})

    argmax, argmin, cumsum, cumprod, nonzero, argsort, searchsorted = [_canon_method(_n) for _n in (
        "argmax", "argmin", "cumsum", "cumprod", "nonzero", "argsort", "searchsorted")]

    @property
    def at(self):
        return _AtHelper(self)

    def __getitem__(self, idx):
        base = self.view(np.ndarray)
        idx = tuple(_plain(i) for i in idx) if isinstance(idx, tuple) else _plain(idx)
        nidx, _ = normalize_index(idx, base.shape, clamp=True)
        return wrap(base[nidx])

    def astype(self, dtype, *a, **k):
        return wrap(self.view(np.ndarray).astype(canon_dtype(dtype)))

    def take(self, indices, axis=None, mode=None, fill_value=None, **k):
        from .numpy import take
        return take(self, indices, axis=axis, mode=mode, fill_value=fill_value)

    def copy(self, *a, **k):
        return wrap(np.array(self.view(np.ndarray), copy=True))

    def block_until_ready(self):
        return self

    def __hash__(self):
        return id(self)


def asarray(x, dtype=None):
    """jnp.asarray: Python ints / floats (and lists of them) get the default 32-bit types."""
    if isinstance(x, np.ndarray):
        a = x.view(np.ndarray)
        dt = canon_dtype(dtype) if dtype is not None else canon_dtype(a.dtype)
        return (a.astype(dt) if dt != a.dtype else a).view(Array)
    a = np.asarray(_plain(x) if isinstance(x, (list, tuple)) else x)
    dt = canon_dtype(dtype) if dtype is not None else canon_dtype(a.dtype)
    return a.astype(dt).view(Array)
