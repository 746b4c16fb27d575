"""numpy stand-in for jax — TEST INFRASTRUCTURE for running the unmodified reference source (see ../README.md)."""
import contextlib
import functools

import numpy as _np

from . import _core
from ._core import Array  # noqa: F401
from . import numpy, lax, random, nn, scipy, ops, debug, profiler, tree_util  # noqa: F401,E402
from .tree_util import tree_map as _tree_map, tree_stack as _tree_stack

__version__ = "0.0-numpy-shim"


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def checkpoint(fun=None, **kw):
    return jit(fun, **kw)


remat = checkpoint


def vmap(fun, in_axes=0, out_axes=0, **kw):
    """Loop over the mapped axis, stack the outputs."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaf = a
                while isinstance(leaf, (tuple, list)):
                    leaf = leaf[0]
                n = leaf.shape[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _tree_map(lambda t: _np.take(_core._plain(t), i, axis=ax).view(Array), a) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        oa = out_axes if isinstance(out_axes, int) else 0
        return _tree_stack(outs, axis=oa)
    return mapped


@contextlib.contextmanager
def named_scope(name):
    yield


def grad(*a, **k):
    raise NotImplementedError("jaxshim has no automatic differentiation: use finite differences in JAXSHIM_X64 mode")


value_and_grad = grad


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def devices(*a):
    return ["cpu:numpy-shim"]


def default_backend():
    return "cpu"


def device_put(x, device=None, **kw):
    from ._core import asarray
    return asarray(x)


def device_get(x):
    return x


def block_until_ready(x):
    return x
