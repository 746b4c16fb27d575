"""jax.ops stand-in."""
import numpy as _np

from ._core import _plain, wrap


def segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False, unique_indices=False, bucket_size=None, mode=None):
    """Values whose segment id is outside [0, num_segments) are dropped; negative ids are NOT wrapped
    (jax/_src/ops/scatter.py::_segment_update scatters with normalize_indices=False)."""
    d = _np.asarray(_plain(data))
    ids = _np.asarray(_plain(segment_ids)).astype(_np.int64)
    n = int(ids.max()) + 1 if num_segments is None else int(num_segments)
    # XLA leaves the order of a float scatter-add unspecified: take the exactly rounded sum (accumulate in double)
    acc_dt = _np.float64 if d.dtype == _np.float32 else d.dtype
    out = _np.zeros((n,) + d.shape[1:], dtype=acc_dt)
    ok = (ids >= 0) & (ids < n)
    _np.add.at(out, ids[ok], d[ok].astype(acc_dt))
    return wrap(out.astype(d.dtype))
