"""jax.scipy.special stand-in (scipy's implementations in the operand's precision)."""
import numpy as _np
from scipy import special as _sps

from .._core import FLOAT, _plain, wrap


def _f(fn):
    def call(x):
        a = _np.asarray(_plain(x))
        if a.dtype.kind != "f":
            a = a.astype(FLOAT)
        return wrap(fn(a).astype(a.dtype))
    return call


erf, erfc, log_ndtr, ndtr, erfinv, gammaln = (_f(_sps.erf), _f(_sps.erfc), _f(_sps.log_ndtr), _f(_sps.ndtr), _f(_sps.erfinv),
                                              _f(_sps.gammaln))


def logsumexp(a, axis=None, b=None, keepdims=False, **kw):
    return wrap(_sps.logsumexp(_np.asarray(_plain(a)), axis=axis, b=_plain(b), keepdims=keepdims))
