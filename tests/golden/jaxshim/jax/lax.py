"""jax.lax stand-in (scan / cond / cummax / top_k / stop_gradient), eager numpy."""
import numpy as _np

from ._core import _plain, wrap
from .tree_util import tree_map, tree_stack


def stop_gradient(x):
    return x


def cummax(x, axis=0, reverse=False):
    a = _np.asarray(_plain(x))
    if reverse:
        return wrap(_np.flip(_np.maximum.accumulate(_np.flip(a, axis), axis=axis), axis))
    return wrap(_np.maximum.accumulate(a, axis=axis))


def top_k(x, k):
    """Largest k entries along the last axis, ties -> lower index first."""
    a = _np.asarray(_plain(x))
    order = _np.argsort(-a, axis=-1, kind="stable")[..., :k]
    return wrap(_np.take_along_axis(a, order, -1)), wrap(order)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def scan(f, init, xs=None, length=None):
    n = length if xs is None else len(_first_leaf(xs))
    carry, ys = init, []
    for i in range(n):
        x = None if xs is None else tree_map(lambda a: a[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (tree_stack(ys) if ys else None)


def _first_leaf(t):
    while isinstance(t, (tuple, list)):
        t = t[0]
    return t
