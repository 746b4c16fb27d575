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
                      npix_capacity=None, n_events=None, n_segments=None, fused=True):
    """The reference-facing entry for production batches: RAW (un-chopped) segment rows as they come out of the input file
    (optimize/dataio.py:133-141) -> hits.  The reference chops on the host (a Python loop per raw row, :63-106) and uploads
    104 B per CHOPPED segment; here the raw rows are uploaded (100-700x fewer) and expanded on the device by the chop
    kernels, then simulate_wfs + simulate_stochastic run as usual.  ``raw_tracks``: (M, n_fields) float32, a pinned host
    tensor / numpy array (uploaded here) or a CUDA tensor; event ids must be batch-local.  ``fused`` (default): the chop runs
    INSIDE the prepare kernel, piece by piece in registers; ``n_segments`` = segment slots of the batch (None: the piece count is
    read back once).  Returns the 8-tuple of simulate_stochastic."""
    from . import sim
    if precision is None:
        precision = float(params.electron_sampling_resolution)
    if not torch.is_tensor(raw_tracks):
        raw_tracks = torch.from_numpy(raw_tracks)
    if not raw_tracks.is_cuda:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        raw_tracks = raw_tracks.to(dev, non_blocking=True)
    if n_events is None:
        n_events = sim.n_events_of(raw_tracks, fields)
    if fused:   # chop_tracks inside the prepare kernel: the chopped (n, 26) batch is never written (larnd_lut_prepare_raw)
        return sim.simulate_hits(params, response_template, raw_tracks.contiguous(), fields, rngseed=rngseed,
                                 npix_capacity=npix_capacity, n_events=n_events, raw=(precision, n_segments))
    else:
        chopped = chop_tracks(raw_tracks, fields, precision)
        wfs, upix = sim.simulate_wfs(params, response_template, chopped, fields, npix_capacity=npix_capacity, n_events=n_events)
    return sim.simulate_stochastic(params, wfs, upix, rngseed)


# ------------------------------------------------------------------------------------------ event-aligned batching
class TracksDataset:
    """Device-resident counterpart of the reference's ``TracksDataset`` (optimize/dataio.py:107-418), same constructor
    arguments and accessors.  The bookkeeping — trajectories = unique (eventID, trackID), trajectories longer than
    ``max_batch_len`` dropped, WHOLE events packed into batches by floor-divide of the cumulative event length, per-batch
    global event ids — is a few thousand raw rows and stays on the host (numpy, vectorised; the reference loops in
    Python).  The rows themselves are uploaded ONCE; ``dataset[i]`` assembles batch i on the device: gather + event-id
    remap (k_batch_gather), chop (k_chop_*), padding with invalid rows to the largest batch (k_pad_rows, driven by the
    device-side row count — no host synchronisation).  Returns CUDA float32 tensors of shape (rows, n_fields).

    ``source``: path of an HDF5 file with a ``segments`` table (read with larndsim_b200.h5io) or the structured array."""

    def __init__(self, filename, nevents=None, max_nbatch=None, swap_xz=True, random_nevents=False, data_seed=42, track_len_sel=2.,
                 max_abs_costheta_sel=0.966, min_abs_segz_sel=15., track_z_bound=28., max_batch_len=50, print_input=False,
                 chopped=True, pad=True, electron_sampling_resolution=0.1, live_selection=False, device=None):
        import numpy as np
        from numpy.lib import recfunctions as rfn
        if isinstance(filename, str):
            from .h5io import read_dataset
            tracks = read_dataset(filename, "segments")
        else:
            tracks = np.array(filename, copy=True)
        if swap_xz:                                   # the drift axis is z in the simulation (dataio.py:117-128)
            for a, b in (("x_start", "z_start"), ("x_end", "z_end"), ("x", "z")):
                tmp = tracks[a].copy()
                tracks[a] = tracks[b]
                tracks[b] = tmp
        if "t0" not in tracks.dtype.names:
            tracks = rfn.append_fields(tracks, "t0", np.zeros(tracks.shape[0]), usemask=False)
        rename = {"event_id": "eventID", "traj_id": "trackID"}
        self.track_fields = tuple(rename.get(f, f) for f in tracks.dtype.names)
        tracks.dtype.names = self.track_fields
        if max_batch_len is not None and max_nbatch is not None and max_nbatch > 0:
            # load only what max_nbatch batches can need, cut at an event boundary (dataio.py:147-157)
            cutoff = int(np.searchsorted(np.cumsum(tracks["dx"]), max_batch_len * (max_nbatch + 2), side="left"))
            if cutoff < len(tracks):
                other = np.nonzero(tracks["eventID"][:cutoff + 1] != tracks["eventID"][cutoff])[0]
                if other.size:
                    tracks = tracks[:int(other[-1]) + 1]
        self.tracks_struct = tracks
        ev, trk = tracks["eventID"].astype(np.int64), tracks["trackID"].astype(np.int64)
        if live_selection:                            # dataio.py:162-186
            sel_rows = np.nonzero(np.abs(tracks["z"]) < min_abs_segz_sel)[0]
            sel = tracks[sel_rows]
            _, first = np.unique(sel[["eventID", "trackID"]], return_index=True)
            first = np.sort(first)
            last = np.r_[first[1:] - 1, len(sel) - 1]
            a, b = sel[first], sel[last]
            d = np.column_stack((a["x_start"] - b["x_end"], a["y_start"] - b["y_end"], a["z_start"] - b["z_end"]))
            cos_theta = np.abs(d[:, 2]) / (np.linalg.norm(d, axis=1) + 1e-10)
            ok = (np.sqrt((d ** 2).sum(axis=1)) > track_len_sel) & (cos_theta < max_abs_costheta_sel) & \
                 (np.maximum(np.abs(a["z"]), np.abs(b["z"])) < track_z_bound)
            key_rows = sel_rows[np.repeat(ok, last - first + 1)]
        else:
            key_rows = np.arange(len(tracks))
        # trajectories: rows grouped by (eventID, trackID) in key order, rows in file order inside a trajectory.
        # NOTE (reference quirk, kept): with live_selection the inverse index is over the SELECTED rows but is used to
        # address rows of the full table (dataio.py:189-207); without it (every production script) they coincide.
        keys = np.ascontiguousarray(tracks[key_rows][["eventID", "trackID"]])
        self.traj_keys, inverse = np.unique(keys, return_inverse=True)
        order = np.argsort(inverse, kind="stable")
        starts = np.r_[0, np.nonzero(np.diff(inverse[order]))[0] + 1, len(order)]
        traj_rows = [order[s:e] for s, e in zip(starts[:-1], starts[1:])] if len(order) else []
        traj_event = np.array([ev[r[0]] for r in traj_rows], dtype=np.int64)
        uniq_ev, first_ev = np.unique(traj_event, return_index=True)
        ordered_events = uniq_ev[np.argsort(first_ev)]
        if nevents is not None and nevents > 0:
            if random_nevents and nevents < len(ordered_events):
                chosen = np.random.default_rng(seed=data_seed).choice(ordered_events, size=nevents, replace=False)
            else:
                chosen = ordered_events if random_nevents else ordered_events[:nevents]
            keep = np.nonzero(np.isin(traj_event, chosen))[0]
            traj_rows = [traj_rows[i] for i in keep]
        self.trajectory_row_indices = traj_rows
        if max_batch_len is not None:
            traj_len = np.array([tracks["dx"][r].sum() for r in traj_rows])
            valid = np.nonzero(traj_len <= max_batch_len)[0]
            if valid.size == 0:
                raise ValueError("All tracks are longer than the batch size! Please check.")
            uev, inv = np.unique(np.array([ev[traj_rows[v][0]] for v in valid]), return_inverse=True)
            ev_len = np.zeros(len(uev))
            np.add.at(ev_len, inv, traj_len[valid])            # same accumulation order as the reference's loop
            cum = np.cumsum(ev_len)
            split = np.nonzero(np.diff(np.floor_divide(cum, max_batch_len)) > 0)[0] + 1
            split = np.r_[0, split, len(cum)]
            if max_nbatch and max_nbatch > 0:
                split = split[:max_nbatch + 1]
            by_event = [valid[inv == e] for e in range(len(uev))]
            self.batch_traj_indices = [[int(t) for e in range(split[i], split[i + 1]) for t in by_event[e]] for i in range(len(split) - 1)]
            self.tot_data_length = cum[split[-1] - 1]
        else:
            self.batch_traj_indices = [[i] for i in range(len(traj_rows))]
            self.tot_data_length = float(sum(tracks["dx"][r].sum() for r in traj_rows))
        if min(len(b) for b in self.batch_traj_indices) == 0:
            raise ValueError("There exist some empty batch in the simulation input!")
        self.batch_row_indices, self.batch_event_global_ids, self.batch_row_keys, self.batch_nsteps = [], [], [], []
        for trajs in self.batch_traj_indices:
            rows = np.concatenate([traj_rows[t] for t in trajs]).astype(np.int64)
            self.batch_row_indices.append(rows)
            self.batch_row_keys.append(np.ascontiguousarray(np.unique(tracks[rows][["eventID", "trackID"]])))
            self.batch_event_global_ids.append(np.unique(ev[rows]))
            self.batch_nsteps.append(int(np.maximum(np.ceil(tracks["dx"][rows] / electron_sampling_resolution), 1).astype(int).sum()))
        self.max_batch_nsteps = max(self.batch_nsteps) if self.batch_nsteps else 0
        self.chopped, self.pad, self.print_input = chopped, pad, print_input
        self.electron_sampling_resolution = electron_sampling_resolution
        # ---- device side (uploaded on first use): the float32 rows once, per batch the row list and the local event id of
        # every row (remap_event_ids_to_local, dataio.py:47-61)
        self.device = device
        self._host_rows = np.stack([tracks[n].astype(np.float32) for n in self.track_fields], axis=1)   # structured_to_unstructured
        self.batch_local_event_ids = [np.searchsorted(g, ev[r]).astype(np.int32)
                                      for r, g in zip(self.batch_row_indices, self.batch_event_global_ids)]
        self._raw = None
        f = self.track_fields
        self._pad_cols = _lib.PadColumns(f.index("eventID"), f.index("trackID") if "trackID" in f else -1,
                                         f.index("pixel_plane") if "pixel_plane" in f else -1)

    def _upload(self):
        if self._raw is None:
            if not torch.cuda.is_available():
                raise _lib.LarndError("TracksDataset batches are assembled on the GPU (larndsim_b200 has no CPU path)")
            self.device = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
            self._raw = torch.from_numpy(self._host_rows).to(self.device)
            self._rows_d = [torch.from_numpy(r).to(self.device) for r in self.batch_row_indices]
            self._local_d = [torch.from_numpy(l).to(self.device) for l in self.batch_local_event_ids]

    def __len__(self):
        return len(self.batch_traj_indices)

    def get_track_fields(self):
        return self.track_fields

    def get_batch_global_event_ids(self, idx=None):
        return self.batch_event_global_ids if idx is None else self.batch_event_global_ids[idx]

    def get_batch_row_keys(self):
        return self.batch_row_keys

    def get_batch_row_indices(self, idx=None):
        return self.batch_row_indices if idx is None else self.batch_row_indices[idx]

    def raw_batch(self, idx):
        """Un-chopped rows of batch idx with batch-local event ids, on the device."""
        self._upload()
        rows, local = self._rows_d[idx], self._local_d[idx]
        out = torch.empty((rows.numel(), self._raw.shape[1]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.get_lib().larnd_batch_gather(C.c_void_p(self._raw.data_ptr()), self._raw.shape[1], C.c_void_p(rows.data_ptr()),
                                                         C.c_void_p(local.data_ptr()), rows.numel(), self.track_fields.index("eventID"),
                                                         C.c_void_p(out.data_ptr()), st))
        return out

    def device_batch(self, idx, capacity=None):
        """Batch idx chopped and padded to ``capacity`` rows (default: its own size with pad=False, the largest batch with
        pad=True) without a host synchronisation: the row counts come from the host-side bookkeeping (batch_nsteps, exact
        by construction) and the padding kernel reads the chop's device-side total."""
        if idx < 0 or idx >= len(self):
            raise IndexError("Batch index out of range")
        raw = self.raw_batch(idx)
        if not self.chopped:
            return raw if capacity is None or capacity <= raw.shape[0] else pad_batch(raw, capacity, self.track_fields)
        if capacity is None:
            capacity = self.max_batch_nsteps if self.pad else self.batch_nsteps[idx]
        capacity = max(int(capacity), self.batch_nsteps[idx])
        off = chop_offsets(raw, self.track_fields, self.electron_sampling_resolution)
        out = torch.empty((capacity, raw.shape[1]), dtype=torch.float32, device=self.device)
        chop_tracks(raw, self.track_fields, self.electron_sampling_resolution, out=out, offsets=off)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.get_lib().larnd_pad_rows(C.c_void_p(out.data_ptr()), out.shape[1], C.c_void_p(off[-1:].data_ptr()), capacity,
                                                     C.byref(self._pad_cols), st))
        return out

    def __getitem__(self, idx):
        return self.device_batch(idx)

    def pad_batch(self, batch_arr, target_len, idx=None):
        """TracksDataset.pad_batch (dataio.py:351-368) on a device batch: invalid rows up to target_len; rows whose local event
        id is not one of the batch's events are invalidated too."""
        out = pad_batch(batch_arr, target_len, self.track_fields)
        if idx is not None:
            f = self.track_fields
            evc = out[:, f.index("eventID")]
            bad = (evc >= 0) & (evc >= len(self.batch_event_global_ids[idx]))
            if bool(bad.any()):
                out = out.clone()
                out[bad, f.index("eventID")] = -1
                for name in ("n_electrons", "dE", "dEdx", "dx", "long_diff", "tran_diff"):
                    if name in f:
                        out[bad, f.index(name)] = 0
                for name in ("trackID", "pixel_plane"):
                    if name in f:
                        out[bad, f.index(name)] = -1
        return out
