"""Input preparation on the device: host-side mirror of the reference's ``optimize/dataio.py`` helpers that sit
directly in front of the simulation (chop_tracks :63-106, pad_batch / _invalidate_rows :340-373).  The reference
chops on the host with a Python loop per raw segment and uploads 104 B per chopped segment; here the raw rows are
uploaded and expanded by the k_chop_* kernels (csrc/chop.cu), bit-identical to numpy's evaluation of the reference
expressions.  File reading and event batching stay host-side Python in the caller (they touch a few thousand rows)."""
import ctypes as C

import torch

from . import _lib

_CHOP_COLS = ("x", "y", "z", "x_start", "y_start", "z_start", "x_end", "y_end", "z_end", "dx", "dE")


def make_chop_columns(fields):
    fields = tuple(fields)
    c = _lib.ChopColumns()
    c.ncols = len(fields)
    for name in _CHOP_COLS:
        if name not in fields:
            raise ValueError("tracks are missing the '%s' column" % name)
        setattr(c, name, fields.index(name))
    return c


def chop_offsets(tracks, fields, precision=0.001):
    """Exclusive prefix (int64, length M+1) of the number of pieces of every raw row; the last entry is the total."""
    if not torch.is_tensor(tracks) or not tracks.is_cuda:
        raise _lib.LarndError("tracks must be a CUDA torch tensor (larndsim_b200 has no CPU path)")
    if tracks.dtype != torch.float32 or tracks.dim() != 2:
        raise ValueError("tracks must be a float32 (N, n_fields) tensor")
    tracks = tracks.contiguous()
    cols = make_chop_columns(fields)
    off = torch.empty(tracks.shape[0] + 1, dtype=torch.int64, device=tracks.device)
    with torch.cuda.device(tracks.device):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.get_lib().larnd_chop_count(C.c_void_p(tracks.data_ptr()), tracks.shape[0], C.byref(cols), float(precision),
                                                   C.c_void_p(off.data_ptr()), st))
    return off


def chop_tracks(tracks, fields, precision=0.001, out=None, offsets=None):
    """Same contract as the reference's chop_tracks(tracks, fields, precision) (optimize/dataio.py:63): every segment is
    cut into ceil(length/precision) pieces.  ``out`` (capacity rows >= total) keeps the call asynchronous; without it
    the total is read back once (8-byte D2H) to size the result.  Returns the (total, n_fields) tensor (a view of
    ``out`` when given: rows beyond the total are untouched)."""
    tracks = tracks.contiguous()
    off = chop_offsets(tracks, fields, precision) if offsets is None else offsets
    cols = make_chop_columns(fields)
    total = None
    if out is None:
        total = int(off[-1].item())
        out = torch.empty((total, tracks.shape[1]), dtype=torch.float32, device=tracks.device)
    with torch.cuda.device(tracks.device):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.get_lib().larnd_chop_tracks(C.c_void_p(tracks.data_ptr()), tracks.shape[0], C.byref(cols), float(precision),
                                                    C.c_void_p(off.data_ptr()), C.c_void_p(out.data_ptr()), out.shape[0], st))
    return out if total is None else out[:total]


def pad_batch(batch, target_len, fields):
    """Pads a chopped batch with invalid rows the way TracksDataset.pad_batch does (optimize/dataio.py:340-373):
    eventID -1, zero n_electrons/dE/dEdx/dx/long_diff/tran_diff, trackID/pixel_plane -1."""
    fields = tuple(fields)
    n = batch.shape[0]
    if target_len <= n:
        return batch
    out = torch.zeros((target_len, batch.shape[1]), dtype=batch.dtype, device=batch.device)
    out[:n] = batch
    out[n:, fields.index("eventID")] = -1
    for name in ("trackID", "pixel_plane"):
        if name in fields:
            out[n:, fields.index(name)] = -1
    return out
