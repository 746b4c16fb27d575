"""Minimal pytree helpers (tuples / lists / None / arrays) for scan and vmap."""
import numpy as _np

from ._core import wrap


def tree_map(f, t):
    if isinstance(t, (tuple, list)):
        return type(t)(tree_map(f, e) for e in t)
    if t is None:
        return None
    return f(t)


def tree_stack(items, axis=0):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return type(first)(tree_stack([it[k] for it in items], axis) for k in range(len(first)))
    if first is None:
        return None
    return wrap(_np.stack([_np.asarray(i) for i in items], axis=axis))
