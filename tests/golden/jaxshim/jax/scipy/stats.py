"""jax.scipy.stats stand-in: norm.pdf / cdf / logpdf evaluated in the operand's precision."""
import numpy as _np
from scipy import special as _sps

from .._core import FLOAT, _plain, wrap


class _Norm:
    @staticmethod
    def _prep(x, loc, scale):
        x, loc, scale = (_np.asarray(_plain(v)) for v in (x, loc, scale))
        dt = _np.result_type(*[v.dtype if v.dtype.kind == "f" else FLOAT for v in (x, loc, scale)])
        return x.astype(dt), loc.astype(dt), scale.astype(dt), dt

    def logpdf(self, x, loc=0, scale=1):
        """jax.scipy.stats.norm.logpdf: -(log(2 pi)/2 + log(scale)) ... in jax: -(z^2)/2 - log(scale) - log(2 pi)/2 with
        z = (x - loc) / scale, formed as written in jax/_src/scipy/stats/norm.py."""
        x, loc, scale, dt = self._prep(x, loc, scale)
        scale_sqrd = scale * scale
        log_normalizer = _np.log(dt.type(2 * _np.pi) * scale_sqrd)
        quadratic = (x - loc) ** 2 / scale_sqrd
        return wrap((-(log_normalizer + quadratic) / dt.type(2)).astype(dt))

    def pdf(self, x, loc=0, scale=1):
        return wrap(_np.exp(_np.asarray(self.logpdf(x, loc, scale))))

    def cdf(self, x, loc=0, scale=1):
        x, loc, scale, dt = self._prep(x, loc, scale)
        return wrap(_sps.ndtr((x - loc) / scale).astype(dt))


norm = _Norm()
