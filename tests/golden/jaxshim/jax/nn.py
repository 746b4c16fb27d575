"""jax.nn stand-in."""
import numpy as _np
from scipy import special as _sps

from ._core import _plain, wrap


def sigmoid(x):
    return wrap(_sps.expit(_np.asarray(_plain(x))))


def softplus(x):
    a = _np.asarray(_plain(x))
    return wrap(_np.logaddexp(a, a.dtype.type(0)))


def logsumexp(a, axis=None, b=None, keepdims=False, **kw):
    return wrap(_sps.logsumexp(_np.asarray(_plain(a)), axis=axis, b=_plain(b), keepdims=keepdims))


def softmax(x, axis=-1):
    return wrap(_sps.softmax(_np.asarray(_plain(x)), axis=axis))
