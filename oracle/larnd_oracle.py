"""ORACLE (test infrastructure only) — CPU restatement of the larnd-sim-jax detector
simulation hot path in numpy.

Every function cites the reference file:line it follows.  The arithmetic is float32 with
the JAX semantics catalogued in SURVEY.md §8c (float ``//`` and ``%`` expansions,
int32 ids, ``searchsorted`` side='left', truncating ``astype(int)``, weak-typed Python
constants rounded to f32 once, dropped negative scatter ids, garbage tick 0).  Passing
``dtype=np.float64`` evaluates the same algorithm in double precision (used by the
finite-difference gradient checks).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  PARITY STATUS: pinned in two ways.  (1) Against the reference-held goldens
(output/jax_ref, produced by real JAX) for geometry / id packing / coordinate and ADC maps
(tests/test_oracle_golden.py).  (2) Against outputs of the reference's own source, imported unmodified and
executed on a numpy stand-in for jax (tests/golden/jaxshim, fixtures tests/golden/refshim_*.npz,
tests/test_refshim_golden.py): drift-stage arrays, unique pixels, waveforms, self-trigger ticks, hit lists,
MC-current mode, probabilistic front end, losses and double-precision finite-difference gradients agree to
~1e-9 in double and to float32 rounding in single precision.  NOT pinned: XLA's own float32 code
generation beyond the three facts the JAX goldens expose (real JAX is not installable here).
"""
import math

import numpy as np
from scipy import special as sps

from .consts import BIRKS, BOX, ELLIPSOID, get_vdrift, linspace_jnp

# Column order of prepared_data/input_*.h5 (optimize/dataio.py:133-141)
FIELDS = ("eventID", "z_end", "trackID", "tran_diff", "z_start", "x_end", "y_end", "n_electrons",
          "pdgId", "x_start", "y_start", "t_start", "t0_start", "t0_end", "t0", "dx", "long_diff",
          "pixel_plane", "t_end", "dEdx", "dE", "t", "y", "x", "z", "n_photons")


# --------------------------------------------------------------------------- JAX arithmetic
def jnp_floor_divide_f(a, b):
    """jnp.floor_divide for floats == _float_divmod (jax/_src/numpy/ufuncs.py):
    round((a - fmod(a,b)) / b) with the sign fix-up.  Used at detsim_jax.py:482-483,507."""
    a = np.asarray(a)
    b = np.asarray(b, dtype=a.dtype)
    mod = np.fmod(a, b)
    div = (a - mod) / b
    ind = (mod != 0) & (np.sign(b) != np.sign(mod))
    div = np.where(ind, div - a.dtype.type(1), div)
    # lax.round: half away from zero
    return (np.sign(div) * np.floor(np.abs(div) + a.dtype.type(0.5))).astype(a.dtype)


def jnp_remainder_f(a, b):
    """jnp.remainder for floats: fmod with sign fix-up (sim_jax.py:406-407)."""
    a = np.asarray(a)
    b = np.asarray(b, dtype=a.dtype)
    m = np.fmod(a, b)
    plus = ((m < 0) != (b < 0)) & (m != 0)
    return np.where(plus, m + b, m).astype(a.dtype)


def _erf(x, dt):
    # XLA's f32 erf is a rational approximation good to ~1 ulp; the oracle takes the correctly
    # rounded value (double erf rounded to the working precision).
    return sps.erf(np.asarray(x, dtype=np.float64)).astype(dt)


def _erfc(x, dt):
    return sps.erfc(np.asarray(x, dtype=np.float64)).astype(dt)


def _sqrt2(dt):
    """jnp.sqrt(2): a weak-typed scalar, i.e. rounded to the working precision."""
    return dt(np.sqrt(dt(2.0)))


def _col(fields, name):
    return fields.index(name)


def _borders(params, dt):
    return np.asarray(params.tpc_borders, dtype=np.float64).astype(dt)


# --------------------------------------------------------------------------- host-side prep
def swap_xz_structured(seg):
    """x<->z swap of the loaders (optimize/dataio.py:117-128, sim_jax.py:27-38)."""
    seg = seg.copy()
    for a, b in (("x_start", "z_start"), ("x_end", "z_end"), ("x", "z")):
        tmp = seg[a].copy()
        seg[a] = seg[b]
        seg[b] = tmp
    return seg


def structured_to_f32(seg):
    """rfn.structured_to_unstructured(..., dtype=float32) (optimize/dataio.py:42-45)."""
    return np.stack([seg[n].astype(np.float32) for n in seg.dtype.names], axis=1)


def chop_tracks(tracks, fields, precision=0.001):
    """Segment subdivision (optimize/dataio.py:63-106), vectorised but arithmetically identical:
    the reference works on the float32 (N,26) array; ``steps*precision*direction`` is evaluated in
    float64 (numpy int64 * Python float * float32 scalar -> float64) and rounded on assignment."""
    tracks = np.asarray(tracks, dtype=np.float32)
    c = lambda n: _col(fields, n)
    start = np.stack([tracks[:, c("x_start")], tracks[:, c("y_start")], tracks[:, c("z_start")]], axis=1)
    end = np.stack([tracks[:, c("x_end")], tracks[:, c("y_end")], tracks[:, c("z_end")]], axis=1)
    seg = end - start
    length = np.sqrt(np.sum(seg ** 2, axis=1))  # float32
    eps = 1e-10
    direction = seg / (length[:, None] + eps)  # float32 (weak python scalar)
    nsteps = np.maximum(np.ceil(length / precision), 1).astype(int).flatten()
    out = []
    for i in range(tracks.shape[0]):
        tr = tracks[i]
        n = nsteps[i]
        ln = length[i]
        d = direction[i]
        new = np.repeat(tr.reshape(1, -1), n, axis=0)
        new[:, c("dE")] = new[:, c("dE")] * precision / (ln + 1e-10)
        steps = np.arange(0, n)
        for k, ax in enumerate("xyz"):
            new[:, c(ax + "_start")] = tr[c(ax + "_start")] + steps * precision * d[k]
            new[:, c(ax + "_end")] = tr[c(ax + "_start")] + precision * (steps + 1) * d[k]
        new[:, c("dx")] = precision
        for ax in "xyz":
            new[-1, c(ax + "_end")] = tr[c(ax + "_end")]
        new[-1, c("dE")] = tr[c("dE")] * (1 - precision * (n - 1) / (ln + 1e-10))
        new[-1, c("dx")] = ln - precision * (n - 1)
        for ax in "xyz":
            new[:, c(ax)] = 0.5 * (new[:, c(ax + "_start")] + new[:, c(ax + "_end")])
        out.append(new)
    return np.vstack(out)


def make_batches(seg_struct, max_batch_len=50.0):
    """Event-aligned batching of TracksDataset (optimize/dataio.py:186-290) for the
    production settings (no live selection, nevents=None): trajectories = unique
    (eventID, trackID); trajectories longer than ``max_batch_len`` are dropped; whole events
    are packed by floor-divide of the cumulative event length.
    Returns a list of (row_indices, sorted_global_event_ids)."""
    keys = np.ascontiguousarray(seg_struct[["eventID", "trackID"]])
    index, inverse = np.unique(keys, return_inverse=True)
    order = np.argsort(inverse, kind="stable")
    sorted_vals = inverse[order]
    _, first = np.unique(sorted_vals, return_index=True)
    last = np.append(first[1:], len(sorted_vals))
    traj_rows = [order[s:e] for s, e in zip(first, last)]
    traj_len = np.array([seg_struct[r]["dx"].sum() for r in traj_rows])
    valid = np.where(traj_len <= max_batch_len)[0]
    ev_of = np.array([seg_struct[traj_rows[v][0]]["eventID"] for v in valid])
    uev, inv = np.unique(ev_of, return_inverse=True)
    groups = [[] for _ in uev]
    ev_len = np.zeros(len(uev))
    for pos, (e, ln) in enumerate(zip(inv, traj_len[valid])):
        groups[e].append(valid[pos])
        ev_len[e] += ln
    cs = np.cumsum(ev_len)
    split = np.where(np.diff(np.floor_divide(cs, max_batch_len)) > 0)[0] + 1
    split = np.insert(np.append(split, len(cs)), 0, 0)
    batches = []
    for i in range(len(split) - 1):
        trajs = []
        for e in range(split[i], split[i + 1]):
            trajs.extend(groups[e])
        rows = np.concatenate([traj_rows[t] for t in trajs])
        gids = np.unique(seg_struct[rows]["eventID"]).astype(np.int64)
        batches.append((rows.astype(np.int64), gids))
    return batches


def batch_array(seg_struct, rows, gids, fields, chop=True, precision=0.005):
    """TracksDataset.__getitem__ (optimize/dataio.py:378-406): float32 rows, local event ids, chop."""
    sel = seg_struct[rows]
    arr = structured_to_f32(sel)
    arr[:, _col(fields, "eventID")] = np.searchsorted(gids, sel["eventID"].astype(np.int64)).astype(np.float32)
    if chop:
        arr = chop_tracks(arr, fields, precision)
    return np.asarray(arr, dtype=np.float32)


def pad_batch(arr, target_len, fields):
    """TracksDataset.pad_batch/_invalidate_rows (optimize/dataio.py:340-373)."""
    arr = np.asarray(arr, dtype=np.float32).copy()
    n = arr.shape[0]
    if target_len > n:
        arr = np.pad(arr, ((0, target_len - n), (0, 0)))
        arr[n:, _col(fields, "eventID")] = -1
        for name in ("n_electrons", "dE", "dEdx", "dx", "long_diff", "tran_diff"):
            arr[n:, _col(fields, name)] = 0
        for name in ("trackID", "pixel_plane"):
            arr[n:, _col(fields, name)] = -1
    return arr


_size_history = {}


def pad_size(cur, tag, thr=0.05, history=None):
    """Shape bucketing (sim_jax.py:61-102) for scalar sizes."""
    hist_all = _size_history if history is None else history
    hist = hist_all.setdefault(tag, [])
    if cur in hist:
        return cur
    for s in hist:
        if cur <= s <= cur * (1 + thr):
            return s
    new = int(cur * (1 + thr / 2) + 0.5)
    hist.append(new)
    hist.sort()
    return new


# --------------------------------------------------------------------------- physics stages
def shift_tracks(params, tracks, fields, dt):
    """sim_jax.py:109-119."""
    t = tracks.copy()
    for ax, s in (("x", params.shift_x), ("y", params.shift_y), ("z", params.shift_z)):
        for suf in ("_start", "_end", ""):
            t[:, _col(fields, ax + suf)] = t[:, _col(fields, ax + suf)] - dt(s)
    return t


def quench(params, tracks, fields, dt):
    """quenching_jax.py:38-75 (Birks / Box / Ellipsoid-Box)."""
    dedx = tracks[:, _col(fields, "dEdx")]
    ef_rho = dt(params.eField * params.lArDensity)
    if params.recombination_mode == BOX:
        csi = dt(params.beta) * dedx / ef_rho
        with np.errstate(divide="ignore", invalid="ignore"):
            recomb = np.maximum(dt(0), np.log(dt(params.alpha) + csi) / csi)
    elif params.recombination_mode == BIRKS:
        recomb = dt(params.Ab) / (dt(1) + dt(params.kb) * dedx / ef_rho)
    elif params.recombination_mode == ELLIPSOID:
        cosphi = np.abs(tracks[:, _col(fields, "z_end")] - tracks[:, _col(fields, "z_start")]) / (
            tracks[:, _col(fields, "dx")] + dt(1e-10))
        b_phi = dt(params.beta) / np.sqrt(dt(1) - cosphi ** 2 + dt(1.0 / params.R_param ** 2) * cosphi ** 2)
        csi = b_phi * dedx / ef_rho
        recomb = np.maximum(dt(0), np.log(dt(params.alpha) + csi) / (csi + dt(1e-10)))
    else:
        raise ValueError("Invalid recombination mode")
    out = tracks.copy()
    out[:, _col(fields, "n_electrons")] = recomb * tracks[:, _col(fields, "dE")] * dt(params.MeVToElectrons)
    return out


def drift(params, tracks, fields, dt, vdrift):
    """drifting_jax.py:19-58."""
    b = _borders(params, dt)
    m = dt(params.size_margin)
    zmin = np.minimum(b[:, 2, 1] - m, b[:, 2, 0] - m)
    zmax = np.maximum(b[:, 2, 1] + m, b[:, 2, 0] + m)
    x = tracks[:, _col(fields, "x")][:, None]
    y = tracks[:, _col(fields, "y")][:, None]
    z = tracks[:, _col(fields, "z")][:, None]
    cond = (x >= b[None, :, 0, 0] - m) & (x <= b[None, :, 0, 1] + m)
    cond &= (y >= b[None, :, 1, 0] - m) & (y <= b[None, :, 1, 1] + m)
    cond &= (z >= zmin[None, :]) & (z <= zmax[None, :])
    mask = cond.sum(axis=-1) >= 1
    plane = cond.astype(np.int32).argmax(axis=-1)
    z_anode = b[plane, 2, 0]
    zc = tracks[:, _col(fields, "z")]
    drift_distance = np.abs(zc - z_anode) + dt(1e-6)
    zs, ze = tracks[:, _col(fields, "z_start")], tracks[:, _col(fields, "z_end")]
    drift_start = np.abs(np.minimum(zs, ze) - z_anode)
    drift_end = np.abs(np.maximum(zs, ze) - z_anode)
    v = dt(vdrift)
    out = tracks.copy()
    out[:, _col(fields, "pixel_plane")] = plane.astype(dt)
    drift_time = drift_distance / v
    lifetime_red = np.exp(-drift_time / dt(params.lifetime))
    maskf = mask.astype(dt)
    out[:, _col(fields, "n_electrons")] = tracks[:, _col(fields, "n_electrons")] * lifetime_red * maskf
    out[:, _col(fields, "long_diff")] = np.sqrt(drift_time * dt(2) * dt(params.long_diff))
    out[:, _col(fields, "tran_diff")] = np.sqrt(drift_time * dt(2) * dt(params.tran_diff))
    t0c = tracks[:, _col(fields, "t0")]
    out[:, _col(fields, "t")] = tracks[:, _col(fields, "t")] + drift_time * maskf + t0c
    out[:, _col(fields, "t_start")] = tracks[:, _col(fields, "t_start")] + (np.minimum(drift_start, drift_end) / v) * maskf + t0c
    out[:, _col(fields, "t_end")] = tracks[:, _col(fields, "t_end")] + (np.maximum(drift_start, drift_end) / v) * maskf + t0c
    return out


def pixel2id(params, px, py, plane, event):
    """detsim_jax.py:232-244 with x64 disabled: int32 wrap-around arithmetic."""
    nx, ny, ntpc = int(params.n_pixels_x), int(params.n_pixels_y), int(np.asarray(params.tpc_borders).shape[0])
    px = np.asarray(px, dtype=np.int64)
    py = np.asarray(py, dtype=np.int64)
    outside = (px >= nx) | (py >= ny) | (px < 0) | (py < 0)
    pid = (np.asarray(event, dtype=np.int64) * ntpc + np.asarray(plane, dtype=np.int64)) * ny + py
    pid = pid * nx + px
    pid = ((pid + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)  # int32 wrap
    return np.where(outside, np.int32(-1), pid).astype(np.int32)


def id2pixel(params, pid):
    """detsim_jax.py:265-278 (integer floor division / modulo, numpy == jnp for ints)."""
    nx, ny, ntpc = int(params.n_pixels_x), int(params.n_pixels_y), int(np.asarray(params.tpc_borders).shape[0])
    pid = np.asarray(pid, dtype=np.int32)
    return pid % nx, (pid // nx) % ny, (pid // (nx * ny)) % ntpc, pid // (nx * ny * ntpc)


def _fma(a, b, c, dt):
    """fused multiply-add with a single rounding (exact product of two float32 fits a double)."""
    if dt is np.float64:
        return a * b + c
    return (np.asarray(a, np.float64) * np.float64(b) + np.asarray(c, np.float64)).astype(dt)


def get_pixel_coordinates(params, xpitch, ypitch, plane, dt=np.float32):
    """detsim_jax.py:296-306.  XLA:CPU contracts ``xpitch*pitch + border`` into one FMA; the goldens
    (output/jax_ref/*.h5 pix_x, pix_y) are reproduced bit for bit only with that contraction."""
    b = _borders(params, dt)[np.asarray(plane).astype(np.int32)]
    pitch = dt(params.pixel_pitch)
    half = dt(params.pixel_pitch / 2)
    px = _fma(np.asarray(xpitch).astype(dt), pitch, b[..., 0, 0], dt) + half
    py = _fma(np.asarray(ypitch).astype(dt), pitch, b[..., 1, 0], dt) + half
    return np.stack([px, py], axis=-1)


def get_hit_z(params, ticks, plane, vdrift, dt=np.float32):
    """detsim_jax.py:308-319 (fixed_v=False)."""
    b = _borders(params, dt)[np.asarray(plane).astype(np.int32)]
    z_anode, z_high = b[..., 2, 0], b[..., 2, 1]
    # XLA folds the two scalar factors: ticks * (t_sampling * v) * sign — pinned bit for bit by the goldens' pix_z
    tv = dt(dt(params.t_sampling) * dt(vdrift))
    return z_anode + np.asarray(ticks, dtype=dt) * tv * np.sign(z_high - z_anode)


def diffusion_weights_1d(bins, x0, sigma, dt):
    """gaussian_1d_integral_new (detsim_jax.py:332-341)."""
    edges = bins[None, :] - x0[:, None]
    e = np.ones_like(edges)
    e[:, 0] = -1
    with np.errstate(divide="ignore", invalid="ignore"):
        e[:, 1:-1] = _erf(edges[:, 1:-1] / (_sqrt2(dt) * sigma[:, None]), dt)
    return (dt(0.5) * (e[:, 1:] - e[:, :-1])).astype(dt)


def simulate_drift_new(params, tracks, fields, dt=np.float32, traced_efield=False):
    """sim_jax.py:375-453.  Returns a dict with the ten reference outputs plus the per-axis
    diffusion weights (handy for kernel-level parity checks)."""
    tracks = np.asarray(tracks, dtype=dt)
    v = get_vdrift(params, traced=traced_efield)
    t = shift_tracks(params, tracks, fields, dt)
    t = quench(params, t, fields, dt)
    t = drift(params, t, fields, dt, v)
    nb = int(params.nb_sampling_bins_per_pixel)
    T = int(params.nb_tran_diff_bins)
    sym = (T - 1) // 2
    n = int(params.number_pix_neighbors)
    b = _borders(params, dt)
    plane = t[:, _col(fields, "pixel_plane")].astype(np.int32)
    event = t[:, _col(fields, "eventID")].astype(np.int32)
    bd = b[plane]
    width = dt(params.pixel_pitch / nb)
    xr = t[:, _col(fields, "x")] - bd[:, 0, 0]
    yr = t[:, _col(fields, "y")] - bd[:, 1, 0]
    # get_bin_shifts, detsim_jax.py:494-512
    bins_pitches = np.stack([jnp_floor_divide_f(xr, width), jnp_floor_divide_f(yr, width)], axis=1).astype(np.int32)
    bins = linspace_jnp(dt(-sym * (params.pixel_pitch / nb)), dt((sym + 1) * (params.pixel_pitch / nb)), T + 1, dtype=dt)
    x0 = jnp_remainder_f(xr, width)
    y0 = jnp_remainder_f(yr, width)
    sigma = t[:, _col(fields, "tran_diff")]
    wx = diffusion_weights_1d(bins, x0, sigma, dt)
    wy = diffusion_weights_1d(bins, y0, sigma, dt)
    w2d = wx[:, :, None] * wy[:, None, :]
    q = t[:, _col(fields, "n_electrons")]
    nelectrons = (w2d * q[:, None, None]).reshape(-1)
    z_cath = bd[:, 2, 1]
    t0 = np.abs(t[:, _col(fields, "z")] - z_cath) / dt(v)
    long_diff = t[:, _col(fields, "long_diff")] / dt(v) / dt(params.t_sampling)
    sh = np.arange(-sym, sym + 1, dtype=np.int32)
    bx = bins_pitches[:, 0][:, None, None] + sh[None, :, None] + np.zeros((1, 1, T), np.int32)
    by = bins_pitches[:, 1][:, None, None] + sh[None, None, :] + np.zeros((1, T, 1), np.int32)
    pixels = pixel2id(params, bx // nb, by // nb, plane[:, None, None], event[:, None, None])
    main_pixels = pixels[:, sym, sym]
    cix = np.abs((bx % nb).astype(dt) - dt(nb // 2) + dt(0.5)).astype(np.int32)
    ciy = np.abs((by % nb).astype(dt) - dt(nb // 2) + dt(0.5)).astype(np.int32)
    currents_idx = np.stack([cix, ciy], axis=-1).reshape(-1, 2)
    g = np.arange(-n, n + 1, dtype=np.int32)
    P = 2 * n + 1
    bxm = (bins_pitches[:, 0] % nb).astype(dt)
    bym = (bins_pitches[:, 1] % nb).astype(dt)
    nix = np.abs(bxm[:, None] - dt(nb // 2) + dt(0.5) - (g * nb).astype(dt)[None, :]).astype(np.int32)
    niy = np.abs(bym[:, None] - dt(nb // 2) + dt(0.5) - (g * nb).astype(dt)[None, :]).astype(np.int32)
    currents_idx_neigh = np.stack([np.broadcast_to(nix[:, :, None], (len(q), P, P)),
                                   np.broadcast_to(niy[:, None, :], (len(q), P, P))], axis=-1).reshape(-1, 2)
    ppx = bins_pitches[:, 0] // nb
    ppy = bins_pitches[:, 1] // nb
    pids = pixel2id(params, ppx[:, None, None] + g[None, :, None] + np.zeros((1, 1, P), np.int32),
                    ppy[:, None, None] + g[None, None, :] + np.zeros((1, P, 1), np.int32),
                    plane[:, None, None], event[:, None, None])
    pids_neigh = pids.copy()
    pids_neigh[:, n, n] = -999
    return dict(main_pixels=main_pixels, pixels=pixels, nelectrons=nelectrons,
                t0_after_diff=np.repeat(t0, T * T), long_diff=np.repeat(long_diff, T * T),
                currents_idx=currents_idx, pIDs_neigh=pids_neigh, currents_idx_neigh=currents_idx_neigh,
                nelectrons_neigh=q, t0_neigh=t0,
                # extras (not reference outputs)
                wx=wx, wy=wy, x0=x0, y0=y0, sigma_t=sigma, long_diff_seg=long_diff, plane=plane, event=event,
                bins_pitches=bins_pitches, tracks=t, vdrift=v)


def unique_and_renumber(d, pad_to=None, history=None):
    """The unique/renumber block of simulate_wfs (sim_jax.py:717-725).
    ``pad_to``: explicit padded length (capacity mode); otherwise pad_size(...,'unique_pixels',0.2)."""
    uniq = np.unique(d["main_pixels"].ravel())
    uniq = np.append(uniq, -1)
    padded = pad_size(uniq.shape[0], "unique_pixels", 0.2, history) if pad_to is None else int(pad_to)
    if padded < uniq.shape[0]:
        raise ValueError("pad_to smaller than the number of unique pixels + 1")
    uniq = np.sort(np.pad(uniq, (0, padded - uniq.shape[0]), constant_values=-1)).astype(np.int32)
    pn = d["pIDs_neigh"].ravel()
    ren = np.searchsorted(uniq, pn)
    ok = (ren < uniq.size) & (uniq[np.minimum(ren, uniq.size - 1)] == pn)
    ren = np.where(ok, ren, 0)
    return uniq, ren.astype(np.int64)


def response_cumsum(response_template):
    """jnp.cumsum(response_template, axis=-1) (sim_jax.py:228): float32 running sum along time."""
    return np.cumsum(np.asarray(response_template), axis=-1, dtype=np.asarray(response_template).dtype)


def simulate_signals(params, unique_pixels, d, pix_renumbering_neigh, response_template, response_cum=None,
                     dt=np.float32, chunk=256, acc_dtype=np.float64):
    """sim_jax.py:142-286.  The eight index/value streams are built exactly as in the reference
    (chunked over segments only to bound memory) and summed with one scatter-add; contributions
    with flat index outside [0, Npix*Nticks) are dropped like jax.ops.segment_sum does.
    Accumulation is in ``acc_dtype`` (float64 by default: the order of XLA's f32 scatter-add is
    unspecified, so the oracle takes the exactly-rounded sum) and cast to ``dt`` at the end."""
    R = np.asarray(response_template, dtype=dt)
    C = response_cumsum(R) if response_cum is None else np.asarray(response_cum, dtype=dt)
    Npix = unique_pixels.shape[0]
    Nticks = int(params.time_interval[1] / params.t_sampling) + 1
    Ntpl, Nx, Ny, Nt = R.shape
    L = int(params.signal_length)
    T2 = int(params.nb_tran_diff_bins) ** 2
    P2 = (2 * int(params.number_pix_neighbors) + 1) ** 2
    ts = dt(params.t_sampling)
    tv = np.asarray(params.long_diff_template, dtype=dt)   # float32 values in the default mode (consts_jax.py:259)
    wfs = np.zeros(Npix * Nticks, dtype=acc_dtype)
    Nseg = d["nelectrons_neigh"].shape[0]
    ar = np.arange(L)
    local_t = np.arange(Nt - L, Nt)
    one = dt(1)

    def scatter(idx, val):
        idx = idx.ravel()
        val = val.ravel()
        keep = (idx >= 0) & (idx < Npix * Nticks)
        wfs[:] += np.bincount(idx[keep], weights=val[keep].astype(np.float64), minlength=Npix * Nticks).astype(acc_dtype)

    def tick_rule(tt):
        return np.where((tt <= 0) | (tt >= Nticks - 1), 0, tt + 1)

    def start_rule(st):
        return np.where((st <= 0) | (st >= Nticks - 1), 0, st)

    for s0 in range(0, Nseg, chunk):
        s1 = min(Nseg, s0 + chunk)
        m = slice(s0 * T2, s1 * T2)
        pixels = d["pixels"].reshape(-1)[m]
        # 1. main pixels
        ren = np.searchsorted(unique_pixels, pixels)
        ok = unique_pixels[np.minimum(ren, Npix - 1)] == pixels
        ren = np.where(ok & (ren < Npix), ren, -1).astype(np.int64)
        ft = d["t0_after_diff"][m] / ts
        ct = np.clip(np.floor(ft).astype(np.int32), 0, Nt - 1).astype(np.int64)
        frac = (ft - ct.astype(dt)).astype(dt)
        ld = d["long_diff"][m]
        idx = np.clip(np.searchsorted(tv, ld), 1, tv.shape[0] - 2)
        x0, x1, x2 = tv[idx - 1], tv[idx], tv[idx + 1]
        a = (ld - x1) * (ld - x2) / ((x0 - x1) * (x0 - x2))
        b = (ld - x0) * (ld - x2) / ((x1 - x0) * (x1 - x2))
        c = (ld - x0) * (ld - x1) / ((x2 - x0) * (x2 - x1))
        st0 = Nt - L - ct
        st1 = st0 - 1
        tt0 = tick_rule(st0[:, None] + ar)
        tt1 = tick_rule(st1[:, None] + ar)
        ci = d["currents_idx"][m]
        q = d["nelectrons"][m]
        cx, cy = ci[:, 0].astype(np.int64), ci[:, 1].astype(np.int64)
        lt = local_t[None, :]
        vals = (R[idx[:, None], cx[:, None], cy[:, None], lt] * b[:, None]
                + R[(idx - 1)[:, None], cx[:, None], cy[:, None], lt] * a[:, None]
                + R[(idx + 1)[:, None], cx[:, None], cy[:, None], lt] * c[:, None]) * q[:, None]
        scatter(ren[:, None] * Nticks + tt0, vals * (one - frac[:, None]))
        scatter(ren[:, None] * Nticks + tt1, vals * frac[:, None])
        # 3a. main boundary correction
        c0 = C[idx, cx, cy, ct]
        c1 = C[idx, cx, cy, np.clip(ct + 1, 0, Nt - 1)]
        interp = c0 * (one - frac) + c1 * frac
        dmain = (C[idx, cx, cy, Nt - L] - interp) * q
        scatter(start_rule(st0) + ren * Nticks, dmain * (one - frac))
        scatter(start_rule(st1) + ren * Nticks, dmain * frac)
        # 2. neighbours
        mn = slice(s0 * P2, s1 * P2)
        qn = np.repeat(d["nelectrons_neigh"][s0:s1], P2)
        t0n = np.repeat(d["t0_neigh"][s0:s1], P2)
        ftn = t0n / ts
        ctn = np.clip(np.floor(ftn).astype(np.int32), 0, Nt - 1).astype(np.int64)
        fn = (ftn - ctn.astype(dt)).astype(dt)
        sn0 = Nt - L - ctn
        sn1 = sn0 - 1
        rn = pix_renumbering_neigh[mn]
        cin = d["currents_idx_neigh"][mn]
        nx_, ny_ = cin[:, 0].astype(np.int64), cin[:, 1].astype(np.int64)
        inb = (nx_ < Nx) & (ny_ < Ny)  # .take(mode='fill') -> NaN out of bounds; reference configs stay in bounds
        if not inb.all():
            raise ValueError("neighbour response index outside the LUT (number_pix_neighbors too large for this LUT)")
        nv = R[0][nx_[:, None], ny_[:, None], lt] * qn[:, None]
        scatter(rn[:, None] * Nticks + tick_rule(sn0[:, None] + ar), nv * (one - fn[:, None]))
        scatter(rn[:, None] * Nticks + tick_rule(sn1[:, None] + ar), nv * fn[:, None])
        # 3b. neighbour boundary correction
        n0 = C[0][nx_, ny_, ctn]
        n1 = C[0][nx_, ny_, np.clip(ctn + 1, 0, Nt - 1)]
        interp_n = n0 * (one - fn) + n1 * fn
        dn = (C[0][nx_, ny_, Nt - L] - interp_n) * qn
        scatter(start_rule(sn0) + rn * Nticks, dn * (one - fn))
        scatter(start_rule(sn1) + rn * Nticks, dn * fn)
    return wfs.reshape(Npix, Nticks).astype(dt)


# --------------------------------------------------------------------------- legacy entry points (kept by the reference)
def _at_add(wfs_flat, idx, val):
    """``wfs.at[idx].add(val)``: negative indices wrap once, out-of-range ones are dropped; exactly rounded sums."""
    idx = np.asarray(idx, dtype=np.int64).ravel()
    val = np.asarray(val).ravel()
    size = wfs_flat.shape[0]
    idx = np.where(idx < 0, idx + size, idx)
    keep = (idx >= 0) & (idx < size)
    wfs_flat += np.bincount(idx[keep], weights=val[keep].astype(np.float64), minlength=size)


def accumulate_signals(wfs, currents_idx, charge, response, response_cum, pixID, cathode_ticks, signal_length, dt=np.float32):
    """detsim_jax.py:157-205: template-0 response rows scattered at truncated ticks + boundary correction read from the
    FLATTENED ``response_cum`` without a template offset (i.e. its template-0 block)."""
    Npix, Nticks = wfs.shape
    R = np.asarray(response, dtype=dt)
    Nx, Ny, Nt = R.shape
    L = int(signal_length)
    ct = np.asarray(cathode_ticks, dtype=np.int64)
    pid = np.asarray(pixID, dtype=np.int64)
    q = np.asarray(charge, dtype=dt)
    ci = np.asarray(currents_idx, dtype=np.int64).reshape(-1, 2)
    st = Nt - L - ct
    tt = st[:, None] + np.arange(L)
    tt = np.where((tt <= 0) | (tt >= Nticks - 1), 0, tt + 1)
    out = np.asarray(wfs, dtype=np.float64).ravel().copy()
    vals = R[ci[:, 0, None], ci[:, 1, None], np.arange(Nt - L, Nt)[None, :]] * q[:, None]
    _at_add(out, pid[:, None] * Nticks + tt, vals)
    cum_flat = np.asarray(response_cum, dtype=dt).reshape(-1)
    base = (ci[:, 0] * Ny + ci[:, 1]) * Nt
    diff = (cum_flat[base + Nt - L] - cum_flat[base + ct]) * q
    _at_add(out, np.where((st <= 0) | (st >= Nticks - 1), 0, st) + pid * Nticks, diff)
    return out.reshape(Npix, Nticks).astype(dt)


def simulate_signals_new(params, unique_pixels, d, pix_renumbering_neigh, response_template, dt=np.float32):
    """sim_jax.py:456-617 on the arrays of simulate_drift_new (``d``): truncating tick, no sub-tick split, bare
    searchsorted for the main pixels, main boundary correction from the template-0 running sum."""
    R = np.asarray(response_template, dtype=dt)
    C = response_cumsum(R)
    Npix = unique_pixels.shape[0]
    Nticks = int(params.time_interval[1] / params.t_sampling) + 1
    Ntpl, Nx, Ny, Nt = R.shape
    L = int(params.signal_length)
    ts = dt(params.t_sampling)
    tv = np.asarray(params.long_diff_template, dtype=dt)
    ren = np.searchsorted(unique_pixels, d["pixels"].reshape(-1)).astype(np.int64)
    ct = (d["t0_after_diff"] / ts).astype(np.int32).astype(np.int64)
    st = Nt - L - ct
    tt = st[:, None] + np.arange(L)
    tt = np.where((tt <= 0) | (tt >= Nticks - 1), 0, tt + 1)
    ld, q = d["long_diff"], d["nelectrons"]
    idx = np.clip(np.searchsorted(tv, ld), 1, tv.shape[0] - 2)
    x0, x1, x2 = tv[idx - 1], tv[idx], tv[idx + 1]
    a = (ld - x1) * (ld - x2) / ((x0 - x1) * (x0 - x2))
    b = (ld - x0) * (ld - x2) / ((x1 - x0) * (x1 - x2))
    c = (ld - x0) * (ld - x1) / ((x2 - x0) * (x2 - x1))
    ci = d["currents_idx"].astype(np.int64)
    lt = np.arange(Nt - L, Nt)[None, :]
    out = np.zeros(Npix * Nticks, dtype=np.float64)
    flat = ren[:, None] * Nticks + tt
    for coef, off in ((b, 0), (a, -1), (c, 1)):
        _at_add(out, flat, R[(idx + off)[:, None], ci[:, 0, None], ci[:, 1, None], lt] * q[:, None] * coef[:, None])
    cum_flat = C.reshape(-1)
    base = (ci[:, 0] * Ny + ci[:, 1]) * Nt
    diff = (cum_flat[base + Nt - L] - cum_flat[base + ct]) * q
    _at_add(out, np.where((st <= 0) | (st >= Nticks - 1), 0, st) + ren * Nticks, diff)
    wfs = out.reshape(Npix, Nticks).astype(dt)
    P2 = (2 * int(params.number_pix_neighbors) + 1) ** 2
    qn = np.repeat(d["nelectrons_neigh"], P2)
    ctn = (np.repeat(d["t0_neigh"], P2) / ts).astype(np.int32)
    return accumulate_signals(wfs, d["currents_idx_neigh"], qn, R[0], C, pix_renumbering_neigh, ctn, L, dt)


def current_lut(params, response, electrons, pixels_coord, fields, dt=np.float32):
    """detsim_jax.py:642-660."""
    e = np.asarray(electrons, dtype=dt)
    xd = np.abs(e[:, _col(fields, "x")] - pixels_coord[..., 0])
    yd = np.abs(e[:, _col(fields, "y")] - pixels_coord[..., 1])
    t0 = dt(params.response_full_drift_t) - e[:, _col(fields, "t")]
    i = np.clip((xd / dt(params.response_bin_size)).astype(np.int32), 0, response.shape[0] - 1)
    j = np.clip((yd / dt(params.response_bin_size)).astype(np.int32), 0, response.shape[1] - 1)
    return t0, np.stack([i, j], axis=-1)


def simulate_wfs(params, response_template, tracks, fields, dt=np.float32, pad_to=None, history=None,
                 response_cum=None, traced_efield=False, return_aux=False):
    """sim_jax.py:689-736: (wfs[:, 1:], unique_pixels)."""
    d = simulate_drift_new(params, tracks, fields, dt, traced_efield)
    uniq, ren = unique_and_renumber(d, pad_to, history)
    wfs = simulate_signals(params, uniq, d, ren, response_template, response_cum, dt)
    if return_aux:
        return wfs[:, 1:], uniq, d, wfs
    return wfs[:, 1:], uniq


# --------------------------------------------------------------------------- front end
def hold_interval(params):
    """round((3*CLOCK_CYCLE + ADC_HOLD_DELAY*CLOCK_CYCLE)/t_sampling) (fee_jax.py:216)."""
    return round((3 * params.CLOCK_CYCLE + params.ADC_HOLD_DELAY * params.CLOCK_CYCLE) / params.t_sampling)


def digitize(params, integral, dt=np.float32):
    """fee_jax.py:57-71 — no rounding, ADC stays float."""
    x = np.asarray(integral, dtype=dt)
    v = np.maximum(x * dt(params.GAIN) + dt(params.V_PEDESTAL) - dt(params.V_CM), dt(0))
    # (… * ADC_COUNTS / (V_REF - V_CM)): left-to-right, static Python constants
    v = v * dt(params.ADC_COUNTS) / dt(params.V_REF - params.V_CM)
    return np.minimum(v, dt(params.ADC_COUNTS)).astype(dt)


def adc2charge(dw, params, dt=np.float32):
    """losses_jax.py:380-383 (ke-)."""
    # XLA's constant folding + FMA contraction make the jitted float32 expression agree with the correctly rounded
    # result; the goldens' Q column is reproduced bit for bit by evaluating in double and rounding once.
    dw = np.asarray(dw, dtype=np.float64)
    return ((dw / params.ADC_COUNTS * (params.V_REF - params.V_CM) + params.V_CM - params.V_PEDESTAL) / params.GAIN * 1e-3).astype(dt)


def get_adc_values(params, pixels_signals, dt=np.float32, noise=None):
    """Self-trigger + ADC loop (fee_jax.py:170-279).

    ``noise``: None -> noise-free (RESET/UNCORRELATED charges multiply zero-mean normals; with both
    set to 0 the reference is deterministic).  Otherwise a dict of pre-drawn standard normals
    {'base': (Npix,), 'extra': (10,Npix), 'pass': (10,Npix), 'fail': (10,Npix)} so that a caller
    can feed the same draws to the CUDA path (JAX threefry bit-parity is out of scope).
    Returns (adc (Npix,10), ticks (Npix,10)) — ticks float (integer valued) like the reference."""
    w = np.asarray(pixels_signals, dtype=dt)
    Npix, Nt = w.shape
    thr = dt(params.DISCRIMINATION_THRESHOLD)
    interval = hold_interval(params)
    nmax = int(params.MAX_ADC_VALUES)
    q = w * dt(params.t_sampling)
    q_cumsum = np.cumsum(q, axis=-1, dtype=dt)
    reset, unc = dt(params.RESET_NOISE_CHARGE), dt(params.UNCORRELATED_NOISE_CHARGE)
    zeros = np.zeros(Npix, dtype=dt)
    base = zeros if noise is None else np.asarray(noise["base"], dt) * reset
    q_sum = base[:, None] + q_cumsum
    rows = np.arange(Npix)
    adc_out = np.zeros((Npix, nmax), dtype=dt)
    tick_out = np.zeros((Npix, nmax), dtype=dt)
    for it in range(nmax):
        cross = (q_sum[:, 1:] >= thr) & (q_sum[:, :-1] <= thr)
        has = cross.any(axis=1)
        idx_t = np.where(has, cross.argmax(axis=1), Nt - 2)
        end = idx_t + 1 + interval
        end = np.where(end >= Nt, Nt - 1, end)
        q_vals = q_sum[rows, end]
        q_nn = q_cumsum[rows, end]
        extra = zeros if noise is None else np.asarray(noise["extra"][it], dt) * unc
        adc = np.where(q_nn != 0, q_vals + extra, q_nn)
        cond = (adc < thr) | (idx_t == Nt - 2)
        adc = np.where(cond, dt(0), adc)
        ic = np.where(cond, dt(Nt - 2), idx_t.astype(dt))
        if noise is None:
            nb = zeros
        else:
            nb = np.where(cond, np.asarray(noise["fail"][it], dt) * unc, np.asarray(noise["pass"][it], dt) * reset)
        end2 = idx_t + 1 + interval + 1
        end2 = np.where(end2 >= Nt, Nt - 1, end2)
        sub = q_cumsum[rows, end2]
        q_cumsum = q_cumsum - sub[:, None]
        q_cumsum = np.where(q_cumsum < 0, dt(0), q_cumsum)
        q_sum = nb[:, None] + q_cumsum
        adc_out[:, it] = adc
        tick_out[:, it] = ic
    return adc_out, tick_out


def parse_output(params, adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, unique_pixels):
    """Stable compaction of valid hits (sim_jax.py:620-647), already sliced to nb_valid."""
    mask = (hit_prob > params.hit_prob_threshold) & (event[:, None] >= 0) & (unique_pixels[:, None] >= 0)
    fm = mask.ravel()
    k = mask.shape[1]
    return (adcs.ravel()[fm], np.repeat(pixel_x, k)[fm], np.repeat(pixel_y, k)[fm], pixel_z.ravel()[fm],
            ticks.ravel()[fm], hit_prob.ravel()[fm], np.repeat(event, k)[fm], np.repeat(unique_pixels, k)[fm])


def simulate_stochastic(params, wfs, unique_pixels, dt=np.float32, noise=None, traced_efield=False):
    """sim_jax.py:738-769 (noise-free unless ``noise`` draws are supplied)."""
    integral, ticks = get_adc_values(params, wfs, dt, noise)
    hit_prob = np.where(ticks < wfs.shape[1] - 3, dt(1), dt(0))
    adcs = digitize(params, integral, dt)
    px, py, plane, event = id2pixel(params, unique_pixels)
    coords = get_pixel_coordinates(params, px, py, plane, dt)
    v = get_vdrift(params, traced=traced_efield)
    pz = get_hit_z(params, ticks.ravel(), np.repeat(plane, ticks.shape[1]), v, dt).reshape(ticks.shape)
    return parse_output(params, adcs, coords[:, 0], coords[:, 1], pz, ticks, hit_prob, event, unique_pixels)


# --------------------------------------------------------------------------- MC-current mode
_B = (1.060, -0.909, -0.909, 5.856, 0.207, 0.207)
_C = (0.679, -1.083, -1.083, 8.772, -5.521, -5.521)
_D = (2.644, -9.174, -9.174, 13.483, 45.887, 45.887)
_T0 = (2.948, -2.705, -2.705, 4.825, 20.814, 20.814)


def _quad(p, x, y, dt):
    return dt(p[0]) + dt(p[1]) * x + dt(p[2]) * y + dt(p[3]) * x * y + dt(p[4]) * x * x + dt(p[5]) * y * y


def integrated_expon(x, loc, scale, dtk, dt):
    """detsim_jax.py:536-542."""
    z = dt(0)
    n = dt(x.shape[-1])
    half = dt(dtk / 2)
    return (np.exp(np.minimum(z, (loc - x + half) / scale)) - np.exp(np.minimum(z, (loc - x - half) / scale))
            + np.exp(np.minimum(z, (loc - half) / scale)) / n) / dt(dtk)


def emg_pdf(x, mu, sigma, lambd, dt):
    """detsim_jax.py:440-458."""
    coeff = lambd / dt(2)
    expo = coeff * (dt(2) * mu + lambd * sigma ** 2 - dt(2) * x)
    er = _erfc((mu + lambd * sigma ** 2 - x) / (_sqrt2(dt) * sigma), dt)
    return coeff * np.exp(expo) * er


def integrated_expon_diff(x, loc, scale, diff, dtk, dt):
    """detsim_jax.py:461-474."""
    lambd = dt(1) / scale
    a = x - dt(dtk / 2)
    b = x + dt(dtk / 2)
    s2 = _sqrt2(dt)
    up = dt(0.5) * _erf((b - loc) / (s2 * diff), dt) - emg_pdf(b, loc, diff, lambd, dt) / lambd
    lo = dt(0.5) * _erf((a - loc) / (s2 * diff), dt) - emg_pdf(a, loc, diff, lambd, dt) / lambd
    tv = up - lo
    tv = tv / (lo[..., 0] - up[..., -1])[..., None]
    return tv / dt(dtk)


def current_mc(params, electrons, pixels_coord, fields, dt=np.float32, vdrift=None):
    """detsim_jax.py:618-639: (t0_tick (N,), signals (N,51))."""
    nticks = int(5 / params.t_sampling) + 1
    ticks = linspace_jnp(0, 5, nticks, dtype=dt)[None, :]
    x = np.abs(electrons[:, _col(fields, "x")] - pixels_coord[..., 0])[:, None]
    y = np.abs(electrons[:, _col(fields, "y")] - pixels_coord[..., 1])[:, None]
    plane = electrons[:, _col(fields, "pixel_plane")].astype(np.int32)
    z_anode = _borders(params, dt)[plane, 2, 0]
    v = dt(get_vdrift(params) if vdrift is None else vdrift)
    t0 = np.abs(electrons[:, _col(fields, "z")] - z_anode) / v
    t0_tick = (t0 / dt(params.t_sampling) + dt(0.5)).astype(np.int32)
    t0 = (t0 - t0_tick.astype(dt) * dt(params.t_sampling))[:, None]
    dtk = 5.0 / (nticks - 1)
    a = np.minimum(_quad(_B, x, y, dt), dt(1))
    b = _quad(_C, x, y, dt)
    c = _quad(_D, x, y, dt)
    st0 = t0 + _quad(_T0, x, y, dt)
    q = electrons[:, _col(fields, "n_electrons")][:, None]
    if params.diffusion_in_current_sim:
        sig = (electrons[:, _col(fields, "long_diff")] / v)[:, None]
        cur = a * integrated_expon_diff(-ticks, -st0, b, sig, dtk, dt) + (dt(1) - a) * integrated_expon_diff(-ticks, -st0, c, sig, dtk, dt)
    else:
        cur = a * integrated_expon(-ticks, -st0, b, dtk, dt) + (dt(1) - a) * integrated_expon(-ticks, -st0, c, dtk, dt)
    return t0_tick, (cur * q).astype(dt)


def get_pixels(params, electrons, fields, dt=np.float32):
    """detsim_jax.py:477-491."""
    n = int(params.number_pix_neighbors)
    plane = electrons[:, _col(fields, "pixel_plane")].astype(np.int32)
    event = electrons[:, _col(fields, "eventID")].astype(np.int32)
    bd = _borders(params, dt)[plane]
    pitch = dt(params.pixel_pitch)
    px = jnp_floor_divide_f(electrons[:, _col(fields, "x")] - bd[:, 0, 0], pitch).astype(np.int32)
    py = jnp_floor_divide_f(electrons[:, _col(fields, "y")] - bd[:, 1, 0], pitch).astype(np.int32)
    g = np.arange(-n, n + 1, dtype=np.int32)
    X, Y = np.meshgrid(g, g, indexing="ij")
    sx, sy = X.ravel(), Y.ravel()
    return pixel2id(params, px[:, None] + sx[None, :], py[:, None] + sy[None, :], plane[:, None], event[:, None])


def simulate_drift_mc(params, tracks, fields, rnd, dt=np.float32):
    """simulate_drift with mc_diff=True (sim_jax.py:122-139; generate_electrons detsim_jax.py:376-400).
    ``rnd`` = the (N,3) standard normals the reference would draw with random.normal(key1,(N,3))."""
    tracks = np.asarray(tracks, dtype=dt)
    v = get_vdrift(params)
    t = drift(params, quench(params, shift_tracks(params, tracks, fields, dt), fields, dt), fields, dt, v)
    rnd = np.asarray(rnd, dtype=dt)
    e = t.copy()
    e[:, _col(fields, "x")] = t[:, _col(fields, "x")] + rnd[:, 0] * t[:, _col(fields, "tran_diff")]
    e[:, _col(fields, "y")] = t[:, _col(fields, "y")] + rnd[:, 1] * t[:, _col(fields, "tran_diff")]
    if not params.diffusion_in_current_sim:
        e[:, _col(fields, "z")] = t[:, _col(fields, "z")] + rnd[:, 2] * t[:, _col(fields, "long_diff")]
    return e, get_pixels(params, e, fields, dt)


def accumulate_signals_parametrized(wfs, signals, pix, start_ticks):
    """detsim_jax.py:207-228 (scatter into the flattened buffer; negative flat ids wrap like numpy/jnp .at)."""
    Npix, Nticks = wfs.shape
    tt = start_ticks[:, None] + np.arange(signals.shape[1])
    tt = np.where((tt < 0) | (tt >= Nticks - 1), 0, tt + 1)
    flat = (pix[:, None].astype(np.int64) * Nticks + tt).ravel()
    out = wfs.astype(np.float64).ravel()
    np.add.at(out, flat, signals.ravel().astype(np.float64))
    return out.reshape(Npix, Nticks).astype(wfs.dtype)


def simulate_parametrized(params, tracks, fields, rnd, dt=np.float32, pad_to=None, history=None, return_wfs=False):
    """simulate_parametrized (sim_jax.py:339-372) for number_pix_neighbors=0, mc_diff=True,
    noise-free FEE; ``rnd`` replaces the JAX draw (see simulate_drift_mc)."""
    if int(params.number_pix_neighbors) != 0:
        raise ValueError("the parametrized path only works with number_pix_neighbors=0 (SURVEY §3.3)")
    electrons, pids = simulate_drift_mc(params, tracks, fields, rnd, dt)
    pids = pids.ravel()
    uniq = np.unique(pids)
    padded = pad_size(uniq.shape[0], "unique_pixels", 0.05, history) if pad_to is None else int(pad_to)
    uniq = np.sort(np.pad(uniq, (0, padded - uniq.shape[0]), constant_values=-1)).astype(np.int32)
    px, py, plane, _ = id2pixel(params, pids)
    coords = get_pixel_coordinates(params, px, py, plane, dt)
    t0_tick, signals = current_mc(params, electrons, coords, fields, dt)
    ren = np.searchsorted(uniq, pids)
    Nticks = int(params.time_interval[1] / params.t_sampling) + 1
    wfs = accumulate_signals_parametrized(np.zeros((uniq.shape[0], Nticks), dtype=dt), signals, ren,
                                          t0_tick.astype(np.int64) - signals.shape[1])
    out = simulate_stochastic(params, wfs[:, 1:], uniq, dt)
    if return_wfs:
        return out, wfs, uniq
    return out


# --------------------------------------------------------------------------- loss (consumer; for fit-step checks)
def mse_adc(params, Q, x, y, z, hit_prob, event, ref_Q, ref_x, ref_y, ref_z, ref_hit_prob, ref_event, sigma=1.0, lambda_Q=1.0):
    """Weighted-MMD + relative charge loss (losses_jax.py:14-39,58-82), float64."""
    f = np.float64
    w, wr = np.asarray(Q, f) * np.asarray(hit_prob, f), np.asarray(ref_Q, f) * np.asarray(ref_hit_prob, f)
    a = np.stack([np.asarray(x, f) + np.asarray(event, f) * 1e5, np.asarray(y, f), np.asarray(z, f)], -1)
    b = np.stack([np.asarray(ref_x, f) + np.asarray(ref_event, f) * 1e5, np.asarray(ref_y, f), np.asarray(ref_z, f)], -1)

    def k(u, v_):
        d2 = ((u[:, None, :] - v_[None, :, :]) ** 2).sum(-1)
        return np.exp(-d2 / (2 * sigma ** 2))

    kxx = (k(a, a) * w[:, None] * w[None, :]).sum()
    kyy = (k(b, b) * wr[:, None] * wr[None, :]).sum()
    kxy = (k(a, b) * w[:, None] * wr[None, :]).sum()
    sx, sy = w.sum(), wr.sum()
    mmd = kxx / sx ** 2 + kyy / sy ** 2 - 2 * kxy / (sx * sy)
    charge = ((sx - sy) / (sy + 1e-6)) ** 2
    return mmd + lambda_Q * charge
