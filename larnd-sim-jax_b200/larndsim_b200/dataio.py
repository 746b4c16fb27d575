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


# The ten columns the simulation reads (include/larnd_b200.h, larnd_columns_t); everything else in the 26-column file
# layout is bookkeeping the hot path never touches.
PACKED_FIELDS = ("eventID", "x", "y", "z", "z_start", "z_end", "dx", "dEdx", "dE", "t0")


def pack_columns(tracks, fields):
    """(packed (N, 10) array, PACKED_FIELDS): the columns the simulation reads, in the ABI's order.  ``fields`` is an
    argument of every simulate_* entry point, so a packed batch is a valid input as it is: a loader that holds chopped
    batches on the host uploads 40 instead of 104 bytes per segment.  Works on numpy arrays and torch tensors."""
    idx = [tuple(fields).index(n) for n in PACKED_FIELDS]
    if torch.is_tensor(tracks):
        return tracks[:, idx].contiguous(), PACKED_FIELDS
    import numpy as np
    return np.ascontiguousarray(tracks[:, idx]), PACKED_FIELDS


def simulate_from_raw(params, response_template, raw_tracks, fields, precision=None, rngseed=0, device=None,
                      npix_capacity=None, n_events=None):
    """The reference-facing entry for production batches: RAW (un-chopped) segment rows as they come out of the input file
    (optimize/dataio.py:133-141) -> hits.  The reference chops on the host (a Python loop per raw row, :63-106) and uploads
    104 B per CHOPPED segment; here the raw rows are uploaded (100-700x fewer) and expanded on the device by the chop
    kernels, then simulate_wfs + simulate_stochastic run as usual.  ``raw_tracks``: (M, n_fields) float32, a pinned host
    tensor / numpy array (uploaded here) or a CUDA tensor; event ids must be batch-local.  Returns the 8-tuple of
    simulate_stochastic."""
    from . import sim
    if precision is None:
        precision = float(params.electron_sampling_resolution)
    if not torch.is_tensor(raw_tracks):
        raw_tracks = torch.from_numpy(raw_tracks)
    if not raw_tracks.is_cuda:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        raw_tracks = raw_tracks.to(dev, non_blocking=True)
    chopped = chop_tracks(raw_tracks, fields, precision)
    wfs, upix = sim.simulate_wfs(params, response_template, chopped, fields, npix_capacity=npix_capacity, n_events=n_events)
    return sim.simulate_stochastic(params, wfs, upix, rngseed)
