"""Stage-by-stage ("stream-form") operators: the reference functions whose arguments are the materialised per-segment
arrays — ``quench`` (quenching_jax.py:38-75), ``drift`` (drifting_jax.py:19-58), ``simulate_signals``
(sim_jax.py:142-286), ``simulate_drift`` (:120-139), ``simulate_signals_parametrized`` (:289-335), ``current_mc``
(detsim_jax.py:619-639), ``accumulate_signals_parametrized`` (:209-228) — with the reference's argument lists.

The fused kernels behind ``simulate_wfs`` / ``simulate_parametrized`` never build these arrays; the operators here are for
code that calls the stages one by one.  Each is one hand-written kernel behind the C ABI (csrc/stream_ops.cu,
csrc/mc_current.cu) and differentiable w.r.t. its float tensor arguments; gradients w.r.t. ``Params`` leaves flow
through the fused entry points, not through these."""
import ctypes as C

import torch

from . import _lib


def _sim():
    from . import sim
    return sim


def _track_columns(fields):
    f = tuple(fields)
    oc = _lib.TrackColumns()
    for name, _ in _lib.TrackColumns._fields_:
        if name not in f:
            raise ValueError("tracks are missing the '%s' column" % name)
        setattr(oc, name, f.index(name))
    return oc


def tracks_stage(params, tracks, fields, stages):
    """shift (1) | quench (2) | drift (4) on a (N, ncols) float32 CUDA tensor; returns the updated copy."""
    sim = _sim()
    sim._check_cuda(tracks, "tracks")
    tracks = tracks.detach().to(torch.float32).contiguous()
    out = torch.empty_like(tracks)
    pod, cols, oc = sim.make_pod(params), sim.make_columns(fields), _track_columns(fields)
    with torch.cuda.device(tracks.device):
        _lib.check(_lib.get_lib().larnd_tracks_stage(sim._ptr(tracks), tracks.shape[0], C.byref(cols), C.byref(oc), C.byref(pod),
                                                     int(stages), sim._ptr(out), sim._stream()))
    return out


def quench(params, tracks, fields):
    """Recombination: fills the ``n_electrons`` column (Birks / Box / Ellipsoid).  Reference: quenching_jax.py:38-75."""
    return tracks_stage(params, tracks, fields, 2)


def drift(params, tracks, fields):
    """TPC membership, drift time, lifetime attenuation, diffusion sigmas, arrival times.  Reference: drifting_jax.py:19-58."""
    return tracks_stage(params, tracks, fields, 4)


# ------------------------------------------------------------------------------------------ simulate_signals
class _SimulateSignals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t0_after_diff, nelectrons, long_diff, nelectrons_neigh, t0_neigh, ints, params, response_template):
        sim = _sim()
        unique_pixels, pixels, currents_idx, pix_renumbering_neigh, currents_idx_neigh = ints
        dev = unique_pixels.device
        f32 = lambda a: a.detach().to(dev, torch.float32).contiguous().reshape(-1)
        i32 = lambda a: a.detach().to(dev, torch.int32).contiguous()
        lut = sim.get_lut(response_template, params.signal_length, params.nb_sampling_bins_per_pixel, params.number_pix_neighbors)
        pod = sim.make_pod(params, lut.shape)
        args = dict(up=i32(unique_pixels), pix=i32(pixels).reshape(-1), t0=f32(t0_after_diff), q=f32(nelectrons), ld=f32(long_diff),
                    ci=i32(currents_idx).reshape(-1, 2), qn=f32(nelectrons_neigh), rn=i32(pix_renumbering_neigh).reshape(-1),
                    t0n=f32(t0_neigh), cin=i32(currents_idx_neigh).reshape(-1, 2))
        n_main, n_seg = args["pix"].numel(), args["qn"].numel()
        P2 = (2 * pod.number_pix_neighbors + 1) ** 2
        if not (args["t0"].numel() == args["q"].numel() == args["ld"].numel() == args["ci"].shape[0] == n_main):
            raise ValueError("simulate_signals: main-pixel streams have different lengths")
        if args["rn"].numel() != n_seg * P2 or args["cin"].shape[0] != n_seg * P2 or args["t0n"].numel() != n_seg:
            raise ValueError("simulate_signals: neighbour streams do not have N*(2n+1)^2 entries")
        npix = args["up"].numel()
        wfs = torch.empty((npix, pod.n_ticks), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.get_lib().larnd_signals_stream_forward(
                sim._ptr(args["up"]), npix, sim._ptr(args["pix"]), sim._ptr(args["t0"]), sim._ptr(args["q"]), sim._ptr(args["ld"]),
                sim._ptr(args["ci"]), n_main, sim._ptr(args["qn"]), sim._ptr(args["rn"]), sim._ptr(args["t0n"]), sim._ptr(args["cin"]),
                n_seg, C.byref(pod), lut.handle, sim._ptr(wfs), sim._ptr(status), sim._stream()))
        st = int(status.item())
        if st:
            raise ValueError("simulate_signals: response index outside the LUT (%s)" %
                             ("main bins must be < 5" if st & 1 else
                              "template row beyond the truncated response_template bank" if st & 4 else
                              "number_pix_neighbors too large for this LUT"))
        ctx.args, ctx.pod, ctx.lut, ctx.shapes = args, pod, lut, (t0_after_diff.shape, nelectrons.shape, long_diff.shape,
                                                                nelectrons_neigh.shape, t0_neigh.shape)
        return wfs

    @staticmethod
    def backward(ctx, g_wfs):
        sim = _sim()
        a, pod = ctx.args, ctx.pod
        g = g_wfs.contiguous()
        dev = g.device
        n_main, n_seg = a["pix"].numel(), a["qn"].numel()
        g_q, g_t0, g_ld = (torch.empty(max(n_main, 1), dtype=torch.float32, device=dev) for _ in range(3))
        g_qn, g_t0n = (torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev) for _ in range(2))
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.get_lib().larnd_signals_stream_backward(
                sim._ptr(a["up"]), a["up"].numel(), sim._ptr(a["pix"]), sim._ptr(a["t0"]), sim._ptr(a["q"]), sim._ptr(a["ld"]),
                sim._ptr(a["ci"]), n_main, sim._ptr(a["qn"]), sim._ptr(a["rn"]), sim._ptr(a["t0n"]), sim._ptr(a["cin"]), n_seg,
                C.byref(pod), ctx.lut.handle, sim._ptr(g), g.shape[1], sim._ptr(g_q), sim._ptr(g_t0), sim._ptr(g_ld), sim._ptr(g_qn),
                sim._ptr(g_t0n), sim._ptr(status), sim._stream()))
        s = ctx.shapes
        return (g_t0[:n_main].reshape(s[0]), g_q[:n_main].reshape(s[1]), g_ld[:n_main].reshape(s[2]), g_qn[:n_seg].reshape(s[3]),
                g_t0n[:n_seg].reshape(s[4]), None, None, None)


def simulate_signals(params, unique_pixels, pixels, t0_after_diff, response_template, nelectrons, long_diff, currents_idx,
                     nelectrons_neigh, pix_renumbering_neigh, t0_neigh, currents_idx_neigh):
    """(Npix, Nticks) waveforms INCLUDING the garbage column 0 from the materialised per-segment streams, argument for
    argument the reference's ``simulate_signals`` (sim_jax.py:142-286).  Differentiable w.r.t. ``t0_after_diff``,
    ``nelectrons``, ``long_diff``, ``nelectrons_neigh`` and ``t0_neigh``."""
    sim = _sim()
    sim._check_cuda(unique_pixels, "unique_pixels")
    sim._check_cuda(response_template, "response_template")
    ints = (unique_pixels, pixels, currents_idx, pix_renumbering_neigh, currents_idx_neigh)
    return _SimulateSignals.apply(t0_after_diff, nelectrons, long_diff, nelectrons_neigh, t0_neigh, ints, params, response_template)


# ------------------------------------------------------------------------------------------ legacy entry points
def _legacy_status(st, who):
    if st & 7:
        raise ValueError("%s: response index outside the LUT (%s)" % (who, "main bins must be < 5" if st & 1 else
                         "template row beyond the truncated response_template bank" if st & 4 else
                         "neighbour bin outside the response"))
    if st & 8:
        raise ValueError("%s: a drift tick lies outside the response time axis [0, Nt) — the reference would read a neighbouring "
                         "LUT row there" % who)


def simulate_signals_new(params, unique_pixels, pixels, t0_after_diff, response_template, nelectrons, long_diff, currents_idx,
                         nelectrons_neigh, pix_renumbering_neigh, t0_neigh, currents_idx_neigh):
    """The reference's earlier form of ``simulate_signals`` (sim_jax.py:456-617), kept there beside the current one and kept
    here with the same argument list: truncating tick, no sub-tick split, bare ``searchsorted`` for the main pixels,
    boundary correction of the main entries from the running sum of template 0.  Returns (Npix, Nticks) waveforms including
    the garbage column 0.  Forward only (no caller in the reference; gradients flow through ``simulate_signals``)."""
    sim = _sim()
    sim._check_cuda(unique_pixels, "unique_pixels")
    sim._check_cuda(response_template, "response_template")
    dev = unique_pixels.device
    f32 = lambda a: a.detach().to(dev, torch.float32).contiguous().reshape(-1)
    i32 = lambda a: a.detach().to(dev, torch.int32).contiguous()
    lut = sim.get_lut(response_template, params.signal_length, params.nb_sampling_bins_per_pixel, params.number_pix_neighbors)
    pod = sim.make_pod(params, lut.shape)
    up, pix, ci = i32(unique_pixels), i32(pixels).reshape(-1), i32(currents_idx).reshape(-1, 2)
    t0, q, ld = f32(t0_after_diff), f32(nelectrons), f32(long_diff)
    n_main = pix.numel()
    if not (t0.numel() == q.numel() == ld.numel() == ci.shape[0] == n_main):
        raise ValueError("simulate_signals_new: main-pixel streams have different lengths")
    rn, cin = i32(pix_renumbering_neigh).reshape(-1), i32(currents_idx_neigh).reshape(-1, 2)
    npix_n = (2 * pod.number_pix_neighbors + 1) ** 2
    n_ent = rn.numel()
    # jnp.take(nelectrons_neigh, arange(n) // npix, mode='fill', fill_value=0); cathode tick = (t0 / t_sampling).astype(int)
    elec = torch.arange(n_ent, device=dev) // npix_n
    qn_seg, t0n_seg = f32(nelectrons_neigh), f32(t0_neigh)
    ok = elec < qn_seg.numel()
    safe = elec.clamp(max=max(qn_seg.numel() - 1, 0))
    qe = torch.where(ok, qn_seg[safe], torch.zeros((), device=dev)) if n_ent else qn_seg[:0]
    t0e = torch.where(ok, t0n_seg[safe], torch.zeros((), device=dev)) if n_ent else t0n_seg[:0]
    cte = torch.div(t0e, torch.tensor(params.t_sampling, dtype=torch.float32, device=dev)).to(torch.int32)
    npix = up.numel()
    wfs = torch.zeros((npix, pod.n_ticks), dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.get_lib().larnd_signals_legacy_forward(
            sim._ptr(up), npix, sim._ptr(pix), sim._ptr(t0), sim._ptr(q), sim._ptr(ld), sim._ptr(ci), n_main, sim._ptr(qe.contiguous()),
            sim._ptr(rn), sim._ptr(cte.contiguous()), sim._ptr(cin), n_ent, C.byref(pod), lut.handle, sim._ptr(wfs), sim._ptr(status),
            sim._stream()))
    _legacy_status(int(status.item()), "simulate_signals_new")
    return wfs


def accumulate_signals(wfs, currents_idx, charge, response, response_cum, pixID, cathode_ticks, signal_length):
    """``wfs`` + the template-0 response of every (currents_idx, charge, pixID, cathode_ticks) entry, with the boundary
    correction from ``response_cum`` — the reference's ``accumulate_signals`` (detsim_jax.py:157-205), same argument list.
    ``response`` is the (Nx, Ny, Nt) template-0 response; ``response_cum`` is accepted for signature compatibility (its
    template-0 block is the running sum the kernel's tables hold; it is recomputed from ``response``)."""
    sim = _sim()
    sim._check_cuda(wfs, "wfs")
    sim._check_cuda(response, "response")
    if response.dim() != 3:
        raise ValueError("accumulate_signals: response must be (Nx, Ny, Nt)")
    dev = wfs.device
    bank = response.detach().to(torch.float32)[None].expand(3, -1, -1, -1).contiguous()   # the table builder wants >= 3 templates
    with torch.cuda.device(dev):
        lut = sim._LutHandle(bank, int(signal_length), 10, 0)
    pod = _lib.ParamsPOD()
    pod.n_ticks, pod.signal_length, pod.n_templates, pod.t_sampling = int(wfs.shape[1]), int(signal_length), 3, 1.0
    i32 = lambda a: a.detach().to(dev, torch.int32).contiguous()
    ci, pid, ct = i32(currents_idx).reshape(-1, 2), i32(pixID).reshape(-1), i32(cathode_ticks).reshape(-1)
    q = charge.detach().to(dev, torch.float32).contiguous().reshape(-1)
    n = pid.numel()
    if not (ci.shape[0] == ct.numel() == q.numel() == n):
        raise ValueError("accumulate_signals: entry arrays have different lengths")
    out = wfs.detach().to(torch.float32).clone().contiguous()
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    null = C.c_void_p(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.get_lib().larnd_signals_legacy_forward(
            null, int(out.shape[0]), null, null, null, null, null, 0, sim._ptr(q), sim._ptr(pid), sim._ptr(ct), sim._ptr(ci), n,
            C.byref(pod), lut.handle, sim._ptr(out), sim._ptr(status), sim._stream()))
    _legacy_status(int(status.item()), "accumulate_signals")
    return out


def current_lut(params, response, electrons, pixels_coord, fields):
    """(t0, currents_idx): drift time left on the response axis and the response bin of |electron - pixel centre| —
    the reference's ``current_lut`` (detsim_jax.py:642-660)."""
    sim = _sim()
    sim._check_cuda(electrons, "electrons")
    f = tuple(fields)
    el = electrons.detach().to(torch.float32).contiguous()
    pc = pixels_coord.detach().to(el.device, torch.float32).contiguous().reshape(-1, 2)
    n = el.shape[0]
    if pc.shape[0] != n:
        raise ValueError("current_lut: one pixel centre per electron expected")
    t0 = torch.empty(n, dtype=torch.float32, device=el.device)
    idx = torch.empty((n, 2), dtype=torch.int32, device=el.device)
    with torch.cuda.device(el.device):
        _lib.check(_lib.get_lib().larnd_current_lut(sim._ptr(el), n, el.shape[1], f.index("x"), f.index("y"), f.index("t"), sim._ptr(pc),
                                                    float(params.response_full_drift_t), float(params.response_bin_size),
                                                    int(response.shape[0]), int(response.shape[1]), sim._ptr(t0), sim._ptr(idx),
                                                    sim._stream()))
    return t0, idx


# ------------------------------------------------------------------------------------------ MC-current stages
def _current_columns(fields):
    f = tuple(fields)
    c = _lib.CurrentColumns()
    c.ncols = len(f)
    for name in ("x", "y", "z", "long_diff", "n_electrons", "pixel_plane"):
        if name not in f:
            raise ValueError("electrons are missing the '%s' column" % name)
        setattr(c, name, f.index(name))
    return c


class _CurrentMc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, electrons, pixels_coord, params, fields):
        sim = _sim()
        el = electrons.detach().to(torch.float32).contiguous()
        pc = pixels_coord.detach().to(el.device, torch.float32).contiguous().reshape(-1, 2)
        n = el.shape[0]
        if pc.shape[0] != n:
            raise ValueError("current_mc: one pixel centre per electron expected (number_pix_neighbors = 0), got %d for %d"
                             % (pc.shape[0], n))
        pod, cols = sim.make_pod(params), _current_columns(fields)
        t0_tick = torch.empty(n, dtype=torch.int32, device=el.device)
        signals = torch.empty((n, 51), dtype=torch.float32, device=el.device)
        with torch.cuda.device(el.device):
            _lib.check(_lib.get_lib().larnd_current_mc(sim._ptr(el), n, C.byref(cols), sim._ptr(pc), C.byref(pod), sim._ptr(t0_tick),
                                                       sim._ptr(signals), sim._stream()))
        ctx.saved, ctx.pc_shape = (el, pc, pod, cols), pixels_coord.shape
        ctx.mark_non_differentiable(t0_tick)
        return t0_tick, signals

    @staticmethod
    def backward(ctx, _g_tick, g_signals):
        sim = _sim()
        el, pc, pod, cols = ctx.saved
        g = g_signals.contiguous()
        g_el, g_pc = torch.empty_like(el), torch.empty_like(pc)
        with torch.cuda.device(el.device):
            _lib.check(_lib.get_lib().larnd_current_mc_backward(sim._ptr(el), el.shape[0], C.byref(cols), sim._ptr(pc), C.byref(pod),
                                                                sim._ptr(g), sim._ptr(g_el), sim._ptr(g_pc), sim._stream()))
        return g_el, g_pc.reshape(ctx.pc_shape), None, None


def current_mc(params, electrons, pixels_coord, fields):
    """(t0_tick (N,) int32, signals (N, 51)): analytic induced current of every electron cloud on its pixel, 51 ticks of
    0.1 us.  Reference: detsim_jax.py:619-639 (current_model[_diff] :546-615)."""
    _sim()._check_cuda(electrons, "electrons")
    return _CurrentMc.apply(electrons, pixels_coord, params, tuple(fields))


class _AccumulateParametrized(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wfs, signals, pixID, start_ticks):
        sim = _sim()
        out = wfs.detach().to(torch.float32).clone().contiguous()
        sig = signals.detach().to(out.device, torch.float32).contiguous()
        pix = pixID.detach().to(out.device, torch.int32).contiguous().reshape(-1)
        st = start_ticks.detach().to(out.device, torch.int32).contiguous().reshape(-1)
        n = sig.shape[0]
        if pix.numel() != n or st.numel() != n:
            raise ValueError("accumulate_signals_parametrized: signals, pixID and start_ticks disagree on the number of rows")
        with torch.cuda.device(out.device):
            _lib.check(_lib.get_lib().larnd_accumulate_parametrized(sim._ptr(out), out.shape[0], out.shape[1], sim._ptr(sig), sig.shape[1],
                                                                    sim._ptr(pix), sim._ptr(st), n, sim._stream()))
        ctx.saved = (pix, st, sig.shape, out.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        sim = _sim()
        pix, st, sshape, oshape = ctx.saved
        g = g.contiguous()
        g_sig = torch.empty(sshape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.get_lib().larnd_accumulate_parametrized_backward(sim._ptr(g), oshape[0], oshape[1], sim._ptr(g_sig), sshape[1],
                                                                             sim._ptr(pix), sim._ptr(st), sshape[0], sim._stream()))
        return g, g_sig, None, None


def accumulate_signals_parametrized(wfs, signals, pixID, start_ticks):
    """wfs + scatter of the (N, 51) currents at (pixID, start_ticks + k), out-of-range ticks in column 0.
    Reference: detsim_jax.py:209-228."""
    _sim()._check_cuda(wfs, "wfs")
    return _AccumulateParametrized.apply(wfs, signals, pixID, start_ticks)


def simulate_drift(params, tracks, fields, rngkey):
    """(electrons (N, ncols), pIDs (N, P, P)): shift -> quench -> drift -> Gaussian smearing of the electron positions ->
    pixel ids.  Reference: sim_jax.py:120-139 (mc_diff branch; the other branch is unusable in the reference too)."""
    from . import detsim
    new_tracks = tracks_stage(params, tracks, fields, 1 | 2 | 4)
    if params.mc_diff:
        electrons = detsim.generate_electrons(new_tracks, fields, rngkey, not params.diffusion_in_current_sim)
    else:
        electrons = detsim.apply_tran_diff(params, new_tracks, fields)
    return electrons, detsim.get_pixels(params, electrons, fields)


def simulate_signals_parametrized(params, electrons, pIDs, unique_pixels, rngkey, fields):
    """(adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event), dense (Npix, 10) / (Npix,) arrays before ``parse_output``.
    Reference: sim_jax.py:289-335."""
    from . import detsim
    sim = _sim()
    pIDs = pIDs.reshape(-1)
    xp, yp, plane, _ = detsim.id2pixel(params, pIDs)
    pixels_coord = detsim.get_pixel_coordinates(params, xp, yp, plane)
    t0, signals = current_mc(params, electrons, pixels_coord, fields)
    pix_renumbering = torch.searchsorted(unique_pixels.to(torch.int64), pIDs.to(torch.int64))
    nticks_wf = int(params.time_interval[1] / params.t_sampling) + 1
    wfs = torch.zeros((unique_pixels.shape[0], nticks_wf), dtype=torch.float32, device=electrons.device)
    wfs = accumulate_signals_parametrized(wfs, signals, pix_renumbering, t0 - signals.shape[1])
    # get_adc_values + digitize + id2pixel + get_pixel_coordinates are one fused kernel (differentiable w.r.t. the waveforms)
    noise = sim.make_noise(params, unique_pixels.shape[0], rngkey, electrons.device)
    adcs, ticks, pixel_x, pixel_y, event = sim._FeeAdc.apply(wfs[:, 1:], params, unique_pixels.to(torch.int32), noise)
    pixel_plane = detsim.id2pixel(params, unique_pixels)[2]
    pixel_z = detsim.get_hit_z(params, ticks.flatten(), torch.repeat_interleave(pixel_plane, ticks.shape[1]))
    hit_prob = torch.where(ticks < wfs.shape[1] - 3, 1.0, 0.0)
    return adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event
