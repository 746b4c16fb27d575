"""Simulation drivers: host-side mirror of the reference's ``larndsim.sim_jax``.

Same public names and call signatures — ``simulate_wfs`` (sim_jax.py:689), ``simulate_stochastic`` (:738),
``simulate_parametrized`` (:339), ``simulate_drift_new`` (:375), ``simulate_signals`` (:142), ``parse_output``
(:620), ``pad_size`` (:61), ``shift_tracks`` (:109) — but arrays are CUDA ``torch`` tensors and the work is done
by the hand-written sm_100a kernels of liblarnd_b200.so, reached through the C ABI (include/larnd_b200.h) and
wrapped in ``torch.autograd.Function`` (the role ``jax.custom_vjp`` plays for the jax.ffi binding, see
INTEGRATION.md).  There is no CPU path: a missing library or a CPU tensor raises.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .consts import RecombinationMode, linspace_f32, vdrift_and_derivative

size_history_dict = {}


def pad_size(cur_size, tag, pad_threshold=0.05):
    """Shape bucketing with the reference's semantics (sim_jax.py:61-102): reuse a known size within
    ``pad_threshold`` or create ``cur*(1+thr/2)``.  Kept so that output shapes seen by callers are the
    reference's; the kernels themselves never recompile."""
    single = isinstance(cur_size, (int, np.integer))
    cur = (int(cur_size),) if single else tuple(int(c) for c in cur_size)
    hist = size_history_dict.setdefault(tag, [])
    if cur in hist:
        return cur[0] if single else cur
    for size in hist:
        if all(c <= s <= c * (1 + pad_threshold) for c, s in zip(cur, size)):
            return size[0] if single else size
    new = tuple(int(c * (1 + pad_threshold / 2) + 0.5) for c in cur)
    hist.append(new)
    hist.sort()
    return new[0] if single else new


def get_size_history():
    return size_history_dict


def shift_tracks(params, tracks, fields):
    """Subtracts (shift_x, shift_y, shift_z) from the nine coordinate columns (reference: sim_jax.py:109-119)."""
    f = tuple(fields)
    out = tracks.clone()
    for ax in "xyz":
        sh = params.value("shift_" + ax)
        for col in (ax, ax + "_start", ax + "_end"):
            out[:, f.index(col)] -= sh
    return out


# ------------------------------------------------------------------------------------------ parameter block
def _f(params, name):
    return params.value(name)


def fill_pod_leaves(P, params):
    """The fields of the C parameter block that depend on fittable Params fields (everything a fit step changes)."""
    P.Ab, P.kb, P.alpha, P.beta = _f(params, "Ab"), _f(params, "kb"), _f(params, "alpha"), _f(params, "beta")
    P.inv_R2 = 1.0 / _f(params, "R_param") ** 2
    P.efield_rho = _f(params, "eField") * _f(params, "lArDensity")
    P.MeVToElectrons = _f(params, "MeVToElectrons")
    v, dv = vdrift_and_derivative(params)
    P.vdrift, P.dvdrift_dEfield = v, dv
    P.lifetime, P.long_diff, P.tran_diff = _f(params, "lifetime"), _f(params, "long_diff"), _f(params, "tran_diff")
    P.shift_x, P.shift_y, P.shift_z = _f(params, "shift_x"), _f(params, "shift_y"), _f(params, "shift_z")
    P.eField, P.lArDensity, P.R_param = _f(params, "eField"), _f(params, "lArDensity"), _f(params, "R_param")
    P.ts_vdrift = float(np.float32(params.t_sampling) * np.float32(v))
    return P


def make_pod(params, lut_shape=None):
    """The C parameter block of ``params`` (a fresh copy per call).  Building it costs ~0.1 ms of Python, several times per
    batch; Params objects are immutable, so the block is cached on the object, keyed on the identity / version counters of the
    tensor-valued fields (a fit updates those in place).  ``lut_shape``: the bank the block will be used with (validated)."""
    d = params.__dict__
    tf = d.get("_tensor_fields")
    if tf is None:      # the object is immutable: which of its fields are tensors never changes
        tf = tuple(k for k, v in d.items() if not k.startswith("_") and torch.is_tensor(v))
        object.__setattr__(params, "_tensor_fields", tf)
    ver = tuple((id(d[k]), d[k]._version) for k in tf)
    cache = d.get("_pod_cache")
    if cache is None or cache[0] != ver:
        cache = (ver, _build_pod(params, None))
        object.__setattr__(params, "_pod_cache", cache)
    pod = cache[1]
    # the block does not depend on the bank (n_templates is long_diff_template's length); the bank only has to fit it
    if lut_shape is not None and int(lut_shape[0]) > pod.n_templates:
        raise ValueError("response_template has %d templates, params.long_diff_template only %d" % (int(lut_shape[0]), pod.n_templates))
    return _lib.ParamsPOD.from_buffer_copy(pod)


def _build_pod(params, lut_shape=None):
    """Fill the C parameter block.  Python-float constant expressions are evaluated in double and rounded to
    float32 once, which is what the reference's weak-typed constants do under jit."""
    P = _lib.ParamsPOD()
    mode = params.recombination_mode
    P.recombination_mode = mode.value if isinstance(mode, RecombinationMode) else int(mode)
    fill_pod_leaves(P, params)
    P.size_margin = params.size_margin
    borders = np.asarray(params.tpc_borders, dtype=np.float64)
    if borders.ndim != 3 or borders.shape[0] > _lib.MAX_TPC:
        raise ValueError("tpc_borders must have shape (n_tpc<=%d, 3, 2)" % _lib.MAX_TPC)
    P.n_tpc = borders.shape[0]
    for i in range(borders.shape[0]):
        for j in range(3):
            for k in range(2):
                P.tpc_borders[i][j][k] = borders[i, j, k]
    nb = int(params.nb_sampling_bins_per_pixel)
    P.pixel_pitch = params.pixel_pitch
    P.bin_width = params.pixel_pitch / nb
    P.half_pitch = params.pixel_pitch / 2
    P.nb_sampling_bins_per_pixel = nb
    P.n_pixels_x, P.n_pixels_y = int(params.n_pixels_x), int(params.n_pixels_y)
    P.number_pix_neighbors = int(params.number_pix_neighbors)
    if int(params.nb_tran_diff_bins) != _lib.NB_TRAN_BINS:
        raise ValueError("nb_tran_diff_bins must be %d" % _lib.NB_TRAN_BINS)
    sym = (_lib.NB_TRAN_BINS - 1) // 2
    w = params.pixel_pitch / nb
    for i, e in enumerate(linspace_f32(np.float32(-sym * w), np.float32((sym + 1) * w), _lib.NB_TRAN_BINS + 1)):
        P.tran_bin_edges[i] = e
    P.t_sampling = params.t_sampling
    P.n_ticks = int(params.time_interval[1] / params.t_sampling) + 1
    P.signal_length = int(params.signal_length)
    tpl = np.asarray(params.long_diff_template, dtype=np.float32)
    if tpl.shape[0] > _lib.MAX_TEMPLATES:
        raise ValueError("too many longitudinal-diffusion templates")
    if lut_shape is not None and int(lut_shape[0]) > tpl.shape[0]:
        raise ValueError("response_template has %d templates, params.long_diff_template only %d" % (int(lut_shape[0]), tpl.shape[0]))
    # the reference searches / clips with long_diff_template.shape[0] (sim_jax.py:163-164).  A bank with FEWER rows (tests
    # truncate it to bound memory) is accepted; a segment that needs a missing row is flagged on the device (check_state)
    P.n_templates = tpl.shape[0]
    for i, t in enumerate(tpl):
        P.long_diff_template[i] = t
    P.discrimination_threshold = params.DISCRIMINATION_THRESHOLD
    P.reset_noise_charge = params.RESET_NOISE_CHARGE
    P.uncorrelated_noise_charge = params.UNCORRELATED_NOISE_CHARGE
    P.gain, P.v_cm, P.v_pedestal = params.GAIN, params.V_CM, params.V_PEDESTAL
    P.v_ref_minus_cm = params.V_REF - params.V_CM
    P.adc_counts = params.ADC_COUNTS
    P.hit_prob_threshold = params.hit_prob_threshold
    P.hold_interval = round((3 * params.CLOCK_CYCLE + params.ADC_HOLD_DELAY * params.CLOCK_CYCLE) / params.t_sampling)
    P.max_adc_values = int(params.MAX_ADC_VALUES)
    P.diffusion_in_current_sim = int(bool(params.diffusion_in_current_sim))
    return P


def make_columns(fields):
    fields = tuple(fields)
    c = _lib.Columns()
    c.ncols = len(fields)
    for name in ("eventID", "x", "y", "z", "z_start", "z_end", "dx", "dEdx", "dE", "t0"):
        if name not in fields:
            raise ValueError("tracks are missing the '%s' column" % name)
        setattr(c, name, fields.index(name))
    return c


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_cuda(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.LarndError("%s must be a CUDA torch tensor (larndsim_b200 has no CPU path)" % name)


# ------------------------------------------------------------------------------------------ LUT handle cache
class _LutHandle:
    def __init__(self, bank, L, nb, n_neigh):
        self.bank = bank  # keeps the pointer alive / unique
        self.L = L
        self.shape = tuple(bank.shape)
        h = C.c_void_p()
        ntpl, nx, ny, nt = self.shape
        lib = _lib.get_lib()
        _lib.check(lib.larnd_lut_create(_ptr(bank), ntpl, nx, ny, nt, L, _stream(), C.byref(h)))
        self.handle = h
        # the neighbourhood-sum tables belong to the table build, not to the per-batch calls (which take a const handle)
        _lib.check(lib.larnd_lut_prepare_neighbours(h, int(nb), int(n_neigh), _stream()))

    def __del__(self):
        try:
            if self.handle:
                _lib.get_lib().larnd_lut_destroy(self.handle)
        except Exception:
            pass


_lut_cache = {}


def get_lut(response_template, signal_length, nb=10, n_neigh=0):
    """Device tables (compacted response rows + running sums + neighbourhood sums) for (response_template, signal_length,
    nb_sampling_bins_per_pixel, number_pix_neighbors); built once and cached — the reference recomputes
    cumsum(response_template) on every call (sim_jax.py:228)."""
    _check_cuda(response_template, "response_template")
    if response_template.dtype != torch.float32 or response_template.dim() != 4:
        raise ValueError("response_template must be a float32 (n_templates, Nx, Ny, Nt) tensor")
    rt = response_template.contiguous()
    key = (rt.data_ptr(), tuple(rt.shape), int(signal_length), rt.device.index, rt._version, int(nb), int(n_neigh))
    h = _lut_cache.get(key)
    if h is None:
        with torch.cuda.device(rt.device):
            h = _LutHandle(rt, int(signal_length), nb, n_neigh)
        if len(_lut_cache) > 8:
            _lut_cache.pop(next(iter(_lut_cache)))
        _lut_cache[key] = h
    return h


# ------------------------------------------------------------------------------------------ LUT waveform simulation
class LutState:
    """Everything a backward pass (or a kernel-level test) needs from one forward call."""
    __slots__ = ("workspace", "counts", "n", "n_events", "npix", "pod", "lut", "unique_pixels", "wfs_full", "wfs_buf", "flags")


def n_events_of(tracks, fields):
    """Upper bound (max local event id + 1) used to size the pixel bitmap; one tiny device reduction + sync, the
    counterpart of the host-side event-id validation the reference runs per batch (optimize/simulate.py:111-113)."""
    ev = tracks[:, tuple(fields).index("eventID")]
    return max(int(ev.max().item()) + 1, 0) if ev.numel() else 0


_deterministic = False


def set_deterministic(on=True):
    """Process-wide switch for bitwise reproducible waveforms and hit lists — the counterpart of running the reference with
    XLA_FLAGS=--xla_gpu_deterministic_ops (optimize/example_run.py:44-47).  The default kernels add float32 window sums with
    red.global.add in arrival order (the low bits vary from run to run, ~1e-7 relative); with this switch simulate_wfs
    accumulates 64-bit fixed-point integers (order-independent) through the chunk kernels.  Slower at spill-sized batches
    and needs 8 more bytes per waveform sample.  Gradients are not covered (the backward kernels reduce with float atomics)."""
    global _deterministic
    _deterministic = bool(on)


def lut_forward(params, response_template, tracks, fields, npix_capacity=None, n_events=None, flags=0, out=None, deterministic=None,
                raw=None, wfs_zero=False):
    """prepare -> unique/renumber -> accumulate.  ``npix_capacity=None`` reproduces the reference's padded size
    pad_size(n_unique+1,'unique_pixels',0.2) (one 16-byte D2H read, like jnp.unique's sync); an explicit
    capacity keeps the whole call asynchronous.  Returns a LutState.

    ``raw=(precision, n_segments)``: ``tracks`` are RAW (un-chopped) rows; chop_tracks(tracks, fields, precision) is fused into
    the prepare kernel (larnd_lut_prepare_raw) for a batch of ``n_segments`` segment slots (>= the number of pieces; the
    rest are pad_batch's invalid rows) — same records, waveforms and gradients as chopping first, without the chopped
    (n, 26) batch ever being written.  ``n_segments=None`` reads the piece count back once (8-byte D2H).

    ``wfs_zero``: the caller's ``out`` waveform buffer is already all-zero (fresh torch.zeros, or cleaned by
    fee_forward(clear_wfs=True)): the 8 KB-per-row memset is skipped (LARND_FLAG_WFS_ZERO)."""
    _check_cuda(tracks, "tracks")
    if tracks.dtype != torch.float32 or tracks.dim() != 2:
        raise ValueError("tracks must be a float32 (N, n_fields) tensor")
    tracks = tracks.contiguous()
    lib = _lib.get_lib()
    with torch.cuda.device(tracks.device):
        lut = get_lut(response_template, params.signal_length, params.nb_sampling_bins_per_pixel, params.number_pix_neighbors)
        pod = make_pod(params, lut.shape)
        cols = make_columns(fields)
        n = tracks.shape[0]
        if n_events is None:
            n_events = n_events_of(tracks, fields)
        flags = int(flags) | env_flags()
        offsets = None
        if raw is not None:
            from . import dataio
            precision, n_seg = raw
            offsets = dataio.chop_offsets(tracks, fields, precision)
            n = int(offsets[-1].item()) if n_seg is None else int(n_seg)
        ws_bytes = lib.larnd_workspace_bytes(n, n_events, pod.n_tpc, pod.n_pixels_x, pod.n_pixels_y)
        st = LutState()
        st.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=tracks.device)
        st.counts = torch.zeros(4, dtype=torch.int32, device=tracks.device)
        st.n, st.n_events, st.pod, st.lut, st.flags = n, n_events, pod, lut, int(flags)
        if raw is not None:
            ccols = dataio.make_chop_columns(fields)
            _lib.check(lib.larnd_lut_prepare_raw(_ptr(tracks), tracks.shape[0], C.byref(ccols), C.byref(cols), float(precision),
                                                 _ptr(offsets), n, C.byref(pod), lut.handle, n_events, _ptr(st.workspace), ws_bytes,
                                                 _ptr(st.counts), _stream()))
        else:
            _lib.check(lib.larnd_lut_prepare(_ptr(tracks), n, C.byref(cols), C.byref(pod), lut.handle, n_events,
                                             _ptr(st.workspace), ws_bytes, _ptr(st.counts), _stream()))
        if npix_capacity is None:
            cnt = st.counts.cpu()
            if int(cnt[2]) & 4:
                raise ValueError("a segment's longitudinal diffusion needs a template row beyond the (truncated) response_template bank")
            if int(cnt[2]) & 8:
                raise _lib.LarndError("the raw rows chop into more pieces than the batch's n_segments slots")
            if int(cnt[2]) != 0:
                raise ValueError("eventID outside [-1, n_events) found in tracks")
            npix_capacity = pad_size(int(cnt[0]) + 1, "unique_pixels", 0.2)
        st.npix = int(npix_capacity)
        if out is None:
            # -1 everywhere: if the device-side capacity / event-id check trips (counts[2] != 0) the kernels bail out and
            # the caller still sees a well-defined "no pixel" list and all-zero waveforms (see check_state)
            st.unique_pixels = torch.full((st.npix,), -1, dtype=torch.int32, device=tracks.device)
            st.wfs_buf = alloc_wfs(st.npix, pod.n_ticks, tracks.device)
        else:
            st.unique_pixels, st.wfs_buf = out
        if st.wfs_buf.dim() != 2 or st.wfs_buf.shape[0] != st.npix or st.wfs_buf.shape[1] < pod.n_ticks or st.wfs_buf.stride(1) != 1:
            raise ValueError("waveform buffer must be (npix_capacity, >= n_ticks) float32 with unit column stride")
        st.wfs_full = st.wfs_buf[:, :pod.n_ticks]
        if _deterministic if deterministic is None else deterministic:
            scratch = torch.empty(lib.larnd_deterministic_scratch_bytes(st.npix, pod.n_ticks), dtype=torch.uint8, device=tracks.device)
            _lib.check(lib.larnd_lut_accumulate_deterministic(n, C.byref(pod), lut.handle, n_events, st.npix, st.flags, _ptr(st.workspace),
                                                              ws_bytes, _ptr(st.unique_pixels), _ptr(st.wfs_buf), st.wfs_buf.stride(0),
                                                              _ptr(st.counts), _ptr(scratch), scratch.numel(), _stream()))
        else:
            if wfs_zero and out is None:
                raise ValueError("wfs_zero describes a caller-provided `out` buffer")
            _lib.check(lib.larnd_lut_accumulate(n, C.byref(pod), lut.handle, n_events, st.npix,
                                                st.flags | (_lib.FLAG_WFS_ZERO if wfs_zero else 0), _ptr(st.workspace),
                                                ws_bytes, _ptr(st.unique_pixels), _ptr(st.wfs_buf), st.wfs_buf.stride(0),
                                                _ptr(st.counts), _stream()))
    return st


def wfs_row_stride(n_ticks):
    """Row stride (floats) of the internal waveform buffer: a multiple of 4 that leaves three padding columns, so that the
    tile kernel can flush whole 4-tick groups with one 16-byte reduction (include/larnd_b200.h, larnd_lut_forward)."""
    return (int(n_ticks) + 3 + 3) // 4 * 4


def alloc_wfs(npix, n_ticks, device):
    return torch.empty((int(npix), wfs_row_stride(n_ticks)), dtype=torch.float32, device=device)


def env_flags():
    """Kernel-selection switches for tests and tuning, read HERE on the host and passed as explicit LARND_FLAG_* bits (the
    library itself reads no environment variable): LARND_ACC_IMPL = chunk | sorted, LARND_SORTED_SPLIT = 0 (one kernel variant)."""
    f = 0
    impl = os.environ.get("LARND_ACC_IMPL", "")
    if impl.startswith("c"):
        f |= _lib.FLAG_IMPL_CHUNK
    elif impl.startswith("s"):
        f |= _lib.FLAG_IMPL_SORTED
    if os.environ.get("LARND_SORTED_SPLIT", "") == "0":
        f |= _lib.FLAG_NO_SPLIT
    return f


def check_state(st):
    """Raises if the device-side checks of the forward call tripped: bit0 of counts[2] = npix_capacity < n_unique + 1,
    bit1 = an eventID outside [-1, n_events).  With an explicit npix_capacity the forward call is fully asynchronous and
    this check (one 16-byte D2H read) is the caller's; simulate_stochastic / the backward pass run it at their own sync."""
    cnt = st.counts.cpu()
    if int(cnt[2]) & 2:
        raise ValueError("eventID outside [-1, n_events) found in tracks")
    if int(cnt[2]) & 4:
        raise ValueError("a segment's longitudinal diffusion needs a template row beyond the (truncated) response_template bank")
    if int(cnt[2]) & 8:
        raise _lib.LarndError("the raw rows chop into more pieces than the batch's n_segments slots")
    if int(cnt[2]) != 0:
        raise _lib.LarndError("npix_capacity=%d is too small for %d unique pixels (+1 padding entry)" % (st.npix, int(cnt[0])))
    return st


def lut_backward(st, g_wfs, skip_garbage=False):
    """VJP of simulate_wfs: g_wfs is the gradient of the (Npix, Nticks-1) output.  Returns a float32 tensor of
    LARND_NPARAMS parameter gradients (order _lib.PARAM_ORDER)."""
    lib = _lib.get_lib()
    nt1 = st.pod.n_ticks - 1
    g = g_wfs
    if g.dtype != torch.float32 or g.dim() != 2 or g.stride(1) != 1 or g.stride(0) < nt1:
        g = g.contiguous().float()   # fee_backward's strided view of a padded full-row buffer is taken as it is
    if tuple(g.shape) != (st.npix, nt1):
        raise ValueError("gradient shape %s != %s" % (tuple(g.shape), (st.npix, nt1)))
    grad = torch.zeros(_lib.NPARAMS, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        # column c of the full row (c >= 1) is g[:, c-1]: pass g - 1 element with stride Nticks-1
        # the LutState's workspace is written by lut_forward only, so the run / tile tables it built are the ones to reuse
        bflags = (1 if skip_garbage else 0) | (st.flags & ~1) | env_flags() | _lib.FLAG_REUSE_RUNS
        impl = os.environ.get("LARND_BWD_IMPL", "")
        if impl:
            bflags = (bflags & ~(_lib.FLAG_IMPL_CHUNK | _lib.FLAG_IMPL_SORTED)) | (_lib.FLAG_IMPL_SORTED if impl.startswith("s") else _lib.FLAG_IMPL_CHUNK)
        _lib.check(lib.larnd_lut_backward(st.n, C.byref(st.pod), st.lut.handle, st.n_events, st.npix,
                                          bflags, _ptr(st.workspace), st.workspace.numel(),
                                          _ptr(st.counts), C.c_void_p(g.data_ptr() - 4), g.stride(0), _ptr(grad), _stream()))
    return grad


# Batches from this size on hand the front end's VJP to the accumulate VJP as step events (hits_backward) instead of a dense
# (Npix, Nticks) gradient when simulate_wfs -> simulate_stochastic are chained under torch.autograd; below it the chunk kernel
# on dense gradient rows is faster.
STEPS_AUTOGRAD_MIN_SEGMENTS = 200000


class _StepsLink:
    """Shared by the autograd nodes of simulate_wfs and simulate_stochastic when the second consumes the first's output
    directly.  The front end's VJP is a step function per pixel row; its node parks (FeeState, g_adc) here and returns a
    zero-stride all-zero placeholder for d loss / d wfs.  If that placeholder reaches simulate_wfs' node untouched, the
    accumulate VJP runs from the step events (no 2 GB gradient written or read); if autograd added other gradients to it
    (the waveforms had a second consumer), the sum is a fresh tensor and the dense front-end gradient is added to it."""
    __slots__ = ("st", "pending", "placeholder")

    def __init__(self):
        self.st, self.pending, self.placeholder = None, None, None


class _SimulateWfs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, params, response_template, tracks, fields, names, npix_capacity, n_events, holder=None, raw=None):
        st = lut_forward(params, response_template, tracks, fields, npix_capacity, n_events, raw=raw)
        ctx.st = st
        ctx.names = names
        ctx.mark_non_differentiable(st.unique_pixels)
        ctx.link = None
        if holder is not None:
            holder.append(st.counts)
            if st.n >= STEPS_AUTOGRAD_MIN_SEGMENTS and not os.environ.get("LARND_NO_STEPS_AUTOGRAD"):
                ctx.link = _StepsLink()
                ctx.link.st = st
            holder.append(ctx.link)
        return st.wfs_full[:, 1:], st.unique_pixels

    @staticmethod
    def backward(ctx, g_wfs, _g_pix):
        link = ctx.link
        if link is not None and link.pending is not None:
            fs, g_adc = link.pending
            ph = link.placeholder
            link.pending = link.placeholder = None
            if g_wfs.data_ptr() == ph.data_ptr() and g_wfs.stride() == ph.stride():
                grad_all = hits_backward(ctx.st, fs, g_adc)
            else:  # the waveforms had other consumers: their gradients + the dense front-end gradient
                grad_all = lut_backward(ctx.st, g_wfs + fee_backward(fs, g_adc))
        else:
            grad_all = lut_backward(ctx.st, g_wfs)
        idx = torch.tensor([_lib.PARAM_ORDER.index(n) for n in ctx.names], device=grad_all.device)
        return (grad_all[idx],) + (None,) * 9


def simulate_wfs(params, response_template, tracks, fields, npix_capacity=None, n_events=None, raw=None):
    """(wfs (Npix, Nticks-1) float32, unique_pixels (Npix,) int32 sorted, -1 padded at the front).
    Reference: sim_jax.py:689-736.  Differentiable w.r.t. the Params fields built with build_params_class.
    ``raw=(precision, n_segments)``: ``tracks`` are un-chopped rows, chopped inside the prepare kernel (see lut_forward)."""
    leaves = params.grad_leaves()
    if leaves and torch.is_grad_enabled():
        names = tuple(n for n, _ in leaves)
        theta = torch.stack([t.to(tracks.device, torch.float32) for _, t in leaves])
        holder = []
        wfs, upix = _SimulateWfs.apply(theta, params, response_template, tracks, tuple(fields), names, npix_capacity, n_events, holder, raw)
        counts = holder[0]
        if holder[1] is not None:
            wfs._larnd_steps_link = holder[1]
    else:
        st = lut_forward(params, response_template, tracks, fields, npix_capacity, n_events, raw=raw)
        wfs, upix, counts = st.wfs_full[:, 1:], st.unique_pixels, st.counts
    # With an explicit npix_capacity the call is asynchronous and a capacity overflow / bad event id is only flagged on the
    # device: the flag travels with the outputs and is checked at the consumer's own synchronisation point
    # (simulate_stochastic), so a public-API user cannot silently get all-zero waveforms.
    wfs._larnd_state = upix._larnd_state = (counts, int(upix.shape[0]))
    return wfs, upix


def check_outputs(*tensors):
    """Raises if the forward call that produced any of ``tensors`` tripped its device-side checks (see check_state)."""
    for t in tensors:
        stt = getattr(t, "_larnd_state", None)
        if stt is not None:
            cnt = stt[0].cpu()
            if int(cnt[2]) & 2:
                raise ValueError("eventID outside [-1, n_events) found in tracks")
            if int(cnt[2]) & 4:
                raise ValueError("a segment's longitudinal diffusion needs a template row beyond the (truncated) response_template bank")
            if int(cnt[2]) & 8:
                raise _lib.LarndError("the raw rows chop into more pieces than the batch's n_segments slots")
            if int(cnt[2]) != 0:
                raise _lib.LarndError("npix_capacity=%d is too small for %d unique pixels (+1 padding entry)" % (stt[1], int(cnt[0])))
            return


def simulate_signals_state(params, response_template, tracks, fields, **kw):
    """Kernel-level access for tests: returns the LutState (full waveforms incl. garbage column, workspace)."""
    return lut_forward(params, response_template, tracks, fields, **kw)


def simulate_drift_new(params, tracks, fields, response_template=None, n_events=None):
    """The ten per-segment arrays of the reference's simulate_drift_new (sim_jax.py:375-453), rebuilt from the records the
    prepare kernel writes: (main_pixels, pixels (N,5,5), nelectrons (N*25), t0_after_diff (N*25), long_diff (N*25),
    currents_idx (N*25,2), pIDs_neigh (N,P,P), currents_idx_neigh (N*P*P,2), nelectrons_neigh (N), t0_neigh (N)).
    The fused kernels never materialise these streams; this view exists for inspection and tests."""
    from .detsim import pixel2id
    _check_cuda(tracks, "tracks")
    lib = _lib.get_lib()
    tracks = tracks.contiguous()
    n = tracks.shape[0]
    cols = make_columns(fields)
    if n_events is None:
        n_events = n_events_of(tracks, fields)
    with torch.cuda.device(tracks.device):
        if response_template is not None:
            lut = get_lut(response_template, params.signal_length, params.nb_sampling_bins_per_pixel, params.number_pix_neighbors)
            pod = make_pod(params, lut.shape)
        else:
            fake = _FakeLut(params)
            lut, pod = fake, make_pod(params)
            pod.n_templates = fake.shape[0]  # only the template index record depends on it, which is not one of the outputs
        st = LutState()
        ws_bytes = lib.larnd_workspace_bytes(n, n_events, pod.n_tpc, pod.n_pixels_x, pod.n_pixels_y)
        st.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=tracks.device)
        st.counts = torch.zeros(4, dtype=torch.int32, device=tracks.device)
        st.n = n
        _lib.check(lib.larnd_lut_prepare(_ptr(tracks), n, C.byref(cols), C.byref(pod), lut.handle, n_events, _ptr(st.workspace),
                                         ws_bytes, _ptr(st.counts), _stream()))
    r = record_fields(st)
    nb, nn, ntpc = pod.nb_sampling_bins_per_pixel, pod.number_pix_neighbors, pod.n_tpc
    dev = tracks.device
    fd = lambda a, b: torch.div(a, b, rounding_mode="floor")
    ep = r["EP"]
    event, plane = fd(ep, ntpc), ep - fd(ep, ntpc) * ntpc
    k = torch.arange(-2, 3, device=dev, dtype=torch.int32)
    bxs, bys = r["BX"][:, None] + k[None, :], r["BY"][:, None] + k[None, :]
    pixels = pixel2id(params, fd(bxs, nb)[:, :, None].expand(-1, 5, 5), fd(bys, nb)[:, None, :].expand(-1, 5, 5), plane[:, None, None], event[:, None, None])
    wx = torch.stack([r["WX%d" % i] for i in range(5)], 1)
    wy = torch.stack([r["WY%d" % i] for i in range(5)], 1)
    nelectrons = (r["Q"][:, None, None] * (wx[:, :, None] * wy[:, None, :])).reshape(-1)
    t0 = r["FT"] * float(np.float32(params.t_sampling))
    cidx = lambda b: (torch.remainder(b, nb).float() - nb // 2 + 0.5).abs().to(torch.int32)
    currents_idx = torch.stack([cidx(bxs)[:, :, None].expand(-1, 5, 5), cidx(bys)[:, None, :].expand(-1, 5, 5)], -1).reshape(-1, 2)
    g = torch.arange(-nn, nn + 1, device=dev, dtype=torch.int32)
    mpx, mpy = fd(r["BX"], nb), fd(r["BY"], nb)
    pids_neigh = pixel2id(params, (mpx[:, None] + g[None, :])[:, :, None].expand(-1, 2 * nn + 1, 2 * nn + 1),
                          (mpy[:, None] + g[None, :])[:, None, :].expand(-1, 2 * nn + 1, 2 * nn + 1), plane[:, None, None], event[:, None, None]).clone()
    pids_neigh[:, nn, nn] = -999
    cn = lambda b: (torch.remainder(b, nb).float()[:, None] - nb // 2 + 0.5 - (g * nb).float()[None, :]).abs().to(torch.int32)
    cin = torch.stack([cn(r["BX"])[:, :, None].expand(-1, 2 * nn + 1, 2 * nn + 1), cn(r["BY"])[:, None, :].expand(-1, 2 * nn + 1, 2 * nn + 1)], -1).reshape(-1, 2)
    rep = lambda a: a[:, None].expand(-1, 25).reshape(-1)
    return r["MAINPIX"], pixels, nelectrons, rep(t0), rep(r["SL"]), currents_idx, pids_neigh, cin, r["Q"], t0


class _FakeLut:
    """larnd_lut_prepare reads only the LUT's shape (bins per axis for the argument check, Nt for the start tick, which
    is not one of simulate_drift_new's outputs): a zero 3-template bank of the minimal size stands in when no response
    is given."""
    _cache = {}

    def __init__(self, params, nt=1950):
        need = int(params.nb_sampling_bins_per_pixel) * int(params.number_pix_neighbors) + int(params.nb_sampling_bins_per_pixel) // 2
        need = max(need, 5)
        ntpl = 3
        key = (need, int(params.signal_length), nt, torch.cuda.current_device(), ntpl, int(params.nb_sampling_bins_per_pixel),
               int(params.number_pix_neighbors))
        h = _FakeLut._cache.get(key)
        if h is None:
            h = _LutHandle(torch.zeros((ntpl, need, need, nt), dtype=torch.float32, device="cuda"), int(params.signal_length),
                           params.nb_sampling_bins_per_pixel, params.number_pix_neighbors)
            _FakeLut._cache[key] = h
        self.handle, self.shape = h.handle, h.shape


def record_fields(st):
    """Per-segment records written by the prepare kernel as {name: tensor} (ints as int32)."""
    n = st.n
    rec = st.workspace[: len(_lib.REC_FIELDS) * n * 4].view(torch.float32).view(len(_lib.REC_FIELDS), n)
    out = {}
    for i, name in enumerate(_lib.REC_FIELDS):
        out[name] = rec[i].view(torch.int32) if name in _lib.REC_INT_FIELDS else rec[i]
    return out


# stage-by-stage operators with the reference's argument lists (sim_jax.py:142-286, 120-139, 289-335)
from .stream_ops import simulate_drift, simulate_signals, simulate_signals_new, simulate_signals_parametrized  # noqa: E402,F401


# ------------------------------------------------------------------------------------------ front end
class FeeState:
    __slots__ = ("adc", "ticks", "pixel_z", "pixel_x", "pixel_y", "event", "saved", "hits", "n_valid", "npix", "pod")


def fee_forward(params, wfs, unique_pixels, noise=None, compact=True, pod=None, clear_wfs=False):
    """Runs the fused FEE kernel on (Npix, Nticks-1) waveforms (any row stride).  ``clear_wfs``: the waveform buffer is
    scratch of a hits-only pipeline — the kernel zeroes every non-zero sample it has read (LARND_FEE_CLEAR_WFS), leaving the
    padded buffer all-zero for the next lut_forward(out=..., wfs_zero=True)."""
    _check_cuda(wfs, "wfs")
    _check_cuda(unique_pixels, "unique_pixels")
    lib = _lib.get_lib()
    if wfs.dtype != torch.float32 or wfs.dim() != 2 or wfs.stride(1) != 1:
        if clear_wfs:
            raise ValueError("clear_wfs needs the float32 waveform buffer itself (lut_forward's wfs_full[:, 1:]), not a copy")
        wfs = wfs.contiguous().float()
    unique_pixels = unique_pixels.to(torch.int32).contiguous()
    pod = make_pod(params) if pod is None else pod
    npix, ntw = wfs.shape
    if ntw != pod.n_ticks - 1:
        raise ValueError("wfs has %d ticks, params imply %d" % (ntw, pod.n_ticks - 1))
    dev = wfs.device
    fs = FeeState()
    fs.npix, fs.pod = npix, pod
    k = pod.max_adc_values
    with torch.cuda.device(dev):
        fs.adc = torch.empty((npix, k), dtype=torch.float32, device=dev)
        fs.ticks = torch.empty((npix, k), dtype=torch.float32, device=dev)
        fs.pixel_z = torch.empty((npix, k), dtype=torch.float32, device=dev)
        fs.pixel_x = torch.empty(npix, dtype=torch.float32, device=dev)
        fs.pixel_y = torch.empty(npix, dtype=torch.float32, device=dev)
        fs.event = torch.empty(npix, dtype=torch.int32, device=dev)
        fs.saved = torch.zeros((npix, 32), dtype=torch.float32, device=dev)
        fs.n_valid = torch.zeros(1, dtype=torch.int32, device=dev)
        scratch = torch.empty(lib.larnd_fee_scratch_bytes(npix), dtype=torch.uint8, device=dev)
        if compact:
            hf = torch.empty((6, npix * k), dtype=torch.float32, device=dev)
            hi = torch.empty((2, npix * k), dtype=torch.int32, device=dev)
            hp = [_ptr(hf[i]) for i in range(6)] + [_ptr(hi[0]), _ptr(hi[1])]
            fs.hits = (hf, hi)
        else:
            hp = [C.c_void_p(0)] * 8
            fs.hits = None
        if noise is not None:
            noise = noise.to(dev, torch.float32).contiguous()
            if noise.numel() != npix * (1 + 3 * k):
                raise ValueError("noise must hold npix*(1+3*MAX_ADC_VALUES) standard normals")
        _lib.check(lib.larnd_fee_forward_ex(_ptr(wfs), wfs.stride(0), _ptr(unique_pixels), npix, C.byref(pod), _ptr(noise),
                                            _ptr(fs.adc), _ptr(fs.ticks), _ptr(fs.pixel_z), _ptr(fs.pixel_x), _ptr(fs.pixel_y),
                                            _ptr(fs.event), _ptr(fs.saved), *hp, _ptr(fs.n_valid),
                                            _ptr(scratch), scratch.numel(), _lib.FEE_CLEAR_WFS if clear_wfs else 0, _stream()))
    return fs


def fee_backward(fs, g_adc, raw_charge=False):
    """VJP of the front end w.r.t. the (Npix, Nticks-1) waveforms.  The result is a view into a padded full-row buffer
    (wfs_row_stride columns: garbage column 0 and the padding are zero), i.e. exactly the layout lut_backward's tile
    kernel reads with aligned 16-byte loads — no copy between the two VJPs.  raw_charge: g_adc is the gradient w.r.t.
    get_adc_values' integrated charge (no digitiser slope)."""
    lib = _lib.get_lib()
    g = g_adc.contiguous().float()
    nt = fs.pod.n_ticks
    buf = torch.empty((fs.npix, wfs_row_stride(nt)), dtype=torch.float32, device=g.device)
    buf[:, 0] = 0      # the kernel writes columns 1 .. nt-1 of every row; the garbage column and the padding are zeroed here
    buf[:, nt:] = 0
    out = buf[:, 1:nt]
    with torch.cuda.device(g.device):
        _lib.check(lib.larnd_fee_backward(_ptr(g), _ptr(fs.ticks), _ptr(fs.saved), fs.npix, C.byref(fs.pod), _ptr(out), buf.stride(0),
                                          1 if raw_charge else 0, _stream()))
    return out


def hits_backward(st, fs, g_adc, raw_charge=False, steps=None):
    """fee_backward followed by lut_backward WITHOUT the dense (Npix, Nticks) waveform gradient: the front end's VJP is a step
    function per pixel row (<= 20 steps), handed to the accumulate VJP as a 168-byte list per row (larnd_fee_backward_steps ->
    larnd_lut_backward_steps).  ``st``: LutState of the forward, ``fs``: FeeState of fee_forward on st's waveforms, ``g_adc``:
    gradient w.r.t. the dense (Npix, 10) ADC array.  Returns the LARND_NPARAMS parameter gradients like lut_backward."""
    lib = _lib.get_lib()
    g = g_adc.contiguous().float()
    if tuple(g.shape) != (st.npix, fs.pod.max_adc_values) or fs.npix != st.npix:
        raise ValueError("g_adc must be (npix, MAX_ADC_VALUES) of the forward's pixel capacity")
    nbytes = lib.larnd_fee_steps_bytes(st.npix)
    if steps is None:
        steps = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
    grad = torch.zeros(_lib.NPARAMS, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(lib.larnd_fee_backward_steps(_ptr(g), _ptr(fs.saved), _ptr(st.unique_pixels), st.npix, C.byref(fs.pod), _ptr(steps),
                                                steps.numel(), 1 if raw_charge else 0, _stream()))
        bflags = (st.flags & ~1) | env_flags() | _lib.FLAG_REUSE_RUNS
        impl = os.environ.get("LARND_BWD_IMPL", "")
        if impl:
            bflags = (bflags & ~(_lib.FLAG_IMPL_CHUNK | _lib.FLAG_IMPL_SORTED)) | (_lib.FLAG_IMPL_SORTED if impl.startswith("s") else _lib.FLAG_IMPL_CHUNK)
        _lib.check(lib.larnd_lut_backward_steps(st.n, C.byref(st.pod), st.lut.handle, st.n_events, st.npix, bflags, _ptr(st.workspace),
                                                st.workspace.numel(), _ptr(st.counts), _ptr(steps), steps.numel(), _ptr(grad), _stream()))
    return grad


class _FeeAdc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wfs, params, unique_pixels, noise, link=None):
        fs = fee_forward(params, wfs, unique_pixels, noise, compact=False)
        ctx.fs = fs
        ctx.link = link
        ctx.mark_non_differentiable(fs.ticks, fs.pixel_x, fs.pixel_y, fs.event)
        return fs.adc, fs.ticks, fs.pixel_x, fs.pixel_y, fs.event

    @staticmethod
    def backward(ctx, g_adc, *_unused):
        link = ctx.link
        if link is not None and link.pending is None and link.st.npix == ctx.fs.npix:
            # step-event hand-over to simulate_wfs' node (see _StepsLink): nothing dense is written here
            link.pending = (ctx.fs, g_adc)
            link.placeholder = torch.zeros(1, dtype=torch.float32, device=g_adc.device).expand(ctx.fs.npix, ctx.fs.pod.n_ticks - 1)
            return link.placeholder, None, None, None, None
        return fee_backward(ctx.fs, g_adc), None, None, None, None


def make_noise(params, npix, rngseed, device):
    """Standard normals for the FEE noise terms, laid out [base | extra(10) | pass(10) | fail(10)]: the draws
    get_adc_values makes from jax.random.key(rngseed) (fee_jax.py:186,237-255,271), generated on the device with the same
    Threefry-2x32 stream (csrc/rng.cu).  ``rngseed`` may also be a key (two uint32 words, e.g. from jrandom.split)."""
    if params.RESET_NOISE_CHARGE == 0 and params.UNCORRELATED_NOISE_CHARGE == 0:
        return None
    from . import jrandom
    k = rngseed if isinstance(rngseed, (tuple, list)) else jrandom.key(int(rngseed) if rngseed is not None else 0)
    return jrandom.fee_noise(k, npix, int(params.MAX_ADC_VALUES), device)


def parse_output(params, adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, unique_pixels):
    """Stable compaction of valid hits (reference: sim_jax.py:620-647).  Returns the padded-free arrays and
    nb_valid (callers of the reference slice [:nb_valid]; here the slicing is already done)."""
    mask = (hit_prob > params.hit_prob_threshold) & (event[:, None] >= 0) & (unique_pixels[:, None] >= 0)
    k = mask.shape[1]
    # one compaction (the only host synchronisation of this function, like the reference's [:nb_valid]); the eight
    # outputs are gathers with the same index list -- boolean-mask indexing would synchronise once per output
    idx = torch.nonzero(mask.reshape(-1)).squeeze(1)
    rows = torch.div(idx, k, rounding_mode="floor")
    slot = lambda a: a.reshape(-1).index_select(0, idx)
    per_row = lambda a: a.index_select(0, rows)
    out = (slot(adcs), per_row(pixel_x), per_row(pixel_y), slot(pixel_z), slot(ticks), slot(hit_prob), per_row(event),
           per_row(unique_pixels))
    return out + (int(idx.numel()),)


def simulate_stochastic(params, wfs, unique_pixels, rngseed):
    """(adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, hit_pixels), each (nb_valid,).
    Reference: sim_jax.py:738-769.  Differentiable through adcs (w.r.t. wfs) and pixel_z (w.r.t. eField)."""
    from .detsim import get_hit_z
    noise = make_noise(params, wfs.shape[0], rngseed, wfs.device)
    need_grad = torch.is_grad_enabled() and (wfs.requires_grad or bool(params.grad_leaves()))
    if not need_grad:
        fs = fee_forward(params, wfs, unique_pixels, noise, compact=True)
        nv = int(fs.n_valid.item())
        check_outputs(wfs, unique_pixels)   # same synchronisation point: the forward call's device-side flags
        hf, hi = fs.hits
        return hf[0, :nv], hf[1, :nv], hf[2, :nv], hf[3, :nv], hf[4, :nv], hf[5, :nv], hi[0, :nv], hi[1, :nv]
    adcs, ticks, pixel_x, pixel_y, event = _FeeAdc.apply(wfs, params, unique_pixels, noise, getattr(wfs, "_larnd_steps_link", None))
    check_outputs(wfs, unique_pixels)       # parse_output below synchronises anyway
    hit_prob = torch.where(ticks < wfs.shape[1] - 3, 1.0, 0.0)
    plane = torch.div(unique_pixels, params.n_pixels_x * params.n_pixels_y, rounding_mode="floor") % np.asarray(params.tpc_borders).shape[0]
    pixel_z = get_hit_z(params, ticks.reshape(-1), plane.repeat_interleave(ticks.shape[1])).reshape(ticks.shape)
    out = parse_output(params, adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, unique_pixels.to(torch.int32))
    return out[:8]


class HitsArena:
    """Persistent scratch of the hits-only pipeline (simulate_hits): pixel list + padded waveform buffer of one
    (npix_capacity, n_ticks) shape.  The front-end kernel zeroes what it has read (fee_forward(clear_wfs=True)), so after the
    first call the 8 KB-per-row memset of the waveform buffer is never paid again; ``clean`` tracks whether that invariant
    holds (an exception between the two kernels simply costs one memset on the next call)."""
    __slots__ = ("unique_pixels", "wfs", "clean")

    def __init__(self):
        self.unique_pixels = self.wfs = None
        self.clean = False

    def fit(self, npix, n_ticks, device):
        shape = (int(npix), wfs_row_stride(n_ticks))
        if self.wfs is None or tuple(self.wfs.shape) != shape or self.wfs.device != device:
            self.wfs = torch.zeros(shape, dtype=torch.float32, device=device)
            self.unique_pixels = torch.full((int(npix),), -1, dtype=torch.int32, device=device)
            self.clean = True


_default_arenas = {}


def release_arenas():
    """Drops the per-device default arenas of simulate_hits (a spill-sized arena holds ~2 GB of waveform scratch)."""
    _default_arenas.clear()


def simulate_hits(params, response_template, tracks, fields, rngseed=0, npix_capacity=None, n_events=None, raw=None, arena=None):
    """simulate_wfs + simulate_stochastic (sim_jax.py:689-769) as ONE hits-only call: the same 8-tuple of hits, the
    waveforms stay internal scratch (no gradients, nothing to return), which lets the buffer be reused and cleaned by the
    front-end kernel instead of being allocated and memset per batch.  Needs an explicit ``npix_capacity`` (production loops
    size it once); without it, or under autograd, the two-call path runs.  ``raw``: see lut_forward."""
    need_grad = torch.is_grad_enabled() and bool(params.grad_leaves())
    if npix_capacity is None or need_grad:
        wfs, upix = simulate_wfs(params, response_template, tracks, fields, npix_capacity=npix_capacity, n_events=n_events, raw=raw)
        return simulate_stochastic(params, wfs, upix, rngseed)
    _check_cuda(tracks, "tracks")
    dev = tracks.device
    if arena is None:
        arena = _default_arenas.setdefault((dev.type, dev.index), HitsArena())
    n_ticks = int(make_pod(params).n_ticks)
    arena.fit(npix_capacity, n_ticks, dev)
    was_clean, arena.clean = arena.clean, False
    st = lut_forward(params, response_template, tracks, fields, npix_capacity=npix_capacity, n_events=n_events, raw=raw,
                     out=(arena.unique_pixels, arena.wfs), wfs_zero=was_clean)
    noise = make_noise(params, st.npix, rngseed, dev)
    fs = fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, noise, compact=True, pod=st.pod, clear_wfs=True)
    arena.clean = True
    nv = int(fs.n_valid.item())
    check_state(st)
    hf, hi = fs.hits
    return hf[0, :nv], hf[1, :nv], hf[2, :nv], hf[3, :nv], hf[4, :nv], hf[5, :nv], hi[0, :nv], hi[1, :nv]


def simulate_probabilistic(params, wfs, unique_pixels):
    """(adcs_distrib (Npix,10,Nticks-1), pixel_x, pixel_y, ticks_prob (log-probabilities), event) — reference:
    sim_jax.py:772-812.  Differentiable through adcs_distrib and ticks_prob w.r.t. the waveforms."""
    from .detsim import get_pixel_coordinates, id2pixel
    from .fee import digitize, get_adc_values_average_noise_vmap
    ticks_prob, charge_distrib = get_adc_values_average_noise_vmap(params, wfs)
    adcs_distrib = digitize(params, charge_distrib)
    px, py, plane, event = id2pixel(params, unique_pixels)
    coords = get_pixel_coordinates(params, px, py, plane)
    return adcs_distrib, coords[:, 0], coords[:, 1], ticks_prob, event


# ------------------------------------------------------------------------------------------ MC-current mode
def mc_normals(n, rngseed, device):
    """The (N,3) standard normals of generate_electrons: random.normal(rngkey1, (N,3)) with rngkey1, rngkey2 =
    random.split(random.key(rngseed)) (reference: sim_jax.py:359-360, detsim_jax.py:393), same Threefry stream."""
    from . import jrandom
    k1, _ = jrandom.split(jrandom.key(int(rngseed)), 2)
    return jrandom.normal(k1, (n, 3), device)


def mc_forward(params, tracks, fields, rnd, npix_capacity=None, n_events=None):
    """prepare (drift + electron smearing + pixel ids) -> unique -> analytic current -> scatter.  Returns a LutState
    (wfs_full includes the garbage column 0)."""
    _check_cuda(tracks, "tracks")
    tracks = tracks.contiguous()
    rnd = rnd.to(tracks.device, torch.float32).contiguous()
    lib = _lib.get_lib()
    if not hasattr(lib, "larnd_mc_forward"):
        raise _lib.LarndError("library was built without the MC-current kernels")
    with torch.cuda.device(tracks.device):
        pod = make_pod(params)
        cols = make_columns(fields)
        n = tracks.shape[0]
        if n_events is None:
            n_events = n_events_of(tracks, fields)
        ws_bytes = lib.larnd_workspace_bytes(n, n_events, pod.n_tpc, pod.n_pixels_x, pod.n_pixels_y)
        st = LutState()
        st.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=tracks.device)
        st.counts = torch.zeros(4, dtype=torch.int32, device=tracks.device)
        st.n, st.n_events, st.pod, st.lut, st.flags = n, n_events, pod, (cols, rnd), 0
        if npix_capacity is None:
            npix_capacity = _mc_count_unique(lib, tracks, n, cols, pod, rnd, n_events, st, ws_bytes)
        st.npix = int(npix_capacity)
        st.unique_pixels = torch.empty(st.npix, dtype=torch.int32, device=tracks.device)
        st.wfs_full = torch.empty((st.npix, pod.n_ticks), dtype=torch.float32, device=tracks.device)
        _lib.check(lib.larnd_mc_forward(_ptr(tracks), n, C.byref(cols), C.byref(pod), _ptr(rnd), n_events, st.npix,
                                        _ptr(st.workspace), ws_bytes, _ptr(st.unique_pixels), _ptr(st.wfs_full),
                                        _ptr(st.counts), _stream()))
    return st


def _mc_count_unique(lib, tracks, n, cols, pod, rnd, n_events, st, ws_bytes):
    """Exact number of distinct pixel ids -> the reference's pad_size(n_unique, 'unique_pixels') (sim_jax.py:363-364).
    A throw-away forward with capacity 1 fills counts[0] (the kernels bail out on the capacity flag without touching
    the 1-row dummy outputs); one 16-byte D2H read, the counterpart of jnp.unique's host sync."""
    tmp_u = torch.empty(1, dtype=torch.int32, device=tracks.device)
    tmp_w = torch.empty((1, pod.n_ticks), dtype=torch.float32, device=tracks.device)
    rc = lib.larnd_mc_forward(_ptr(tracks), n, C.byref(cols), C.byref(pod), _ptr(rnd), n_events, 1, _ptr(st.workspace), ws_bytes,
                              _ptr(tmp_u), _ptr(tmp_w), _ptr(st.counts), _stream())
    _lib.check(rc)
    cnt = st.counts.cpu()
    if int(cnt[2]) & 2:
        raise ValueError("eventID outside [-1, n_events) found in tracks")
    return pad_size(max(int(cnt[0]), 1), "unique_pixels")


def mc_backward(st, tracks, g_wfs):
    lib = _lib.get_lib()
    g = g_wfs.contiguous()
    nt1 = st.pod.n_ticks - 1
    cols, rnd = st.lut
    grad = torch.zeros(_lib.NPARAMS, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(lib.larnd_mc_backward(_ptr(tracks), st.n, C.byref(cols), C.byref(st.pod), _ptr(rnd), st.n_events, st.npix,
                                         _ptr(st.workspace), st.workspace.numel(), _ptr(st.counts),
                                         C.c_void_p(g.data_ptr() - 4), nt1, _ptr(grad), _stream()))
    return grad


class _SimulateMcWfs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, params, tracks, fields, names, rnd, npix_capacity, n_events):
        st = mc_forward(params, tracks, fields, rnd, npix_capacity, n_events)
        ctx.st, ctx.names, ctx.tracks = st, names, tracks
        ctx.mark_non_differentiable(st.unique_pixels)
        return st.wfs_full[:, 1:], st.unique_pixels

    @staticmethod
    def backward(ctx, g_wfs, _g):
        grad_all = mc_backward(ctx.st, ctx.tracks, g_wfs)
        idx = torch.tensor([_lib.PARAM_ORDER.index(n) for n in ctx.names], device=grad_all.device)
        return (grad_all[idx],) + (None,) * 7


def simulate_parametrized(params, tracks, fields, rngseed=0, rnd=None, npix_capacity=None, n_events=None):
    """MC-current simulation: (adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, unique_pixels) per valid hit.
    Reference: sim_jax.py:339-372 (requires number_pix_neighbors = 0 and mc_diff = True, the only configuration the
    reference itself supports)."""
    if not params.mc_diff:
        raise ValueError("simulate_parametrized requires mc_diff=True (the reference's non-MC branch reads the never-set "
                         "params.tran_diff_bin_edges)")
    if rnd is None:
        rnd = mc_normals(tracks.shape[0], rngseed, tracks.device)
    leaves = params.grad_leaves()
    if leaves and torch.is_grad_enabled():
        names = tuple(n for n, _ in leaves)
        theta = torch.stack([t.to(tracks.device, torch.float32) for _, t in leaves])
        wfs, upix = _SimulateMcWfs.apply(theta, params, tracks, tuple(fields), names, rnd, npix_capacity, n_events)
    else:
        st = mc_forward(params, tracks, fields, rnd, npix_capacity, n_events)
        wfs, upix = st.wfs_full[:, 1:], st.unique_pixels
    from . import jrandom
    _, rngkey2 = jrandom.split(jrandom.key(int(rngseed or 0)), 2)   # the FEE uses the second half of the split (sim_jax.py:360,368)
    return simulate_stochastic(params, wfs, upix, rngkey2)
