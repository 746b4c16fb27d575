"""Front-end electronics: host-side mirror of the reference's ``larndsim.fee_jax`` (digitize :57,
get_adc_values :170).  get_adc_values runs the fused sm_100a FEE kernel."""
import torch

from . import sim as _sim


def digitize(params, integral_list):
    """ADC counts from integrated charge — no rounding, ADC stays float (reference: fee_jax.py:57-71)."""
    x = integral_list if torch.is_tensor(integral_list) else torch.as_tensor(integral_list, dtype=torch.float32)
    v = torch.clamp(x * params.GAIN + params.V_PEDESTAL - params.V_CM, min=0)
    # a true float32 division like XLA's: torch turns `tensor / python_scalar` into a multiplication by the reciprocal on
    # CUDA, which is off by an ulp now and then — divide by a 0-d DEVICE tensor instead
    den = torch.tensor(params.V_REF - params.V_CM, dtype=v.dtype, device=v.device)
    return torch.clamp(v * params.ADC_COUNTS / den, max=params.ADC_COUNTS)


class _AdcValues(torch.autograd.Function):
    """get_adc_values as one kernel launch: the FEE kernel's integrated charge (before the digitiser) and its VJP."""

    @staticmethod
    def forward(ctx, wfs, params, noise):
        upix = torch.zeros(wfs.shape[0], dtype=torch.int32, device=wfs.device)
        fs = _sim.fee_forward(params, wfs, upix, noise, compact=False)
        ctx.fs = fs
        ctx.mark_non_differentiable(fs.ticks)
        return fs.saved[:, 13:13 + fs.pod.max_adc_values].contiguous(), fs.ticks

    @staticmethod
    def backward(ctx, g_q, _g_ticks):
        return _sim.fee_backward(ctx.fs, g_q, raw_charge=True), None, None


def get_adc_values(params, pixels_signals, noise_rng_key=None):
    """(adc (Npix,10) integrated charge in the reference's units, ticks (Npix,10) float, integer valued) — reference:
    fee_jax.py:170-279.  The values are the kernel's own pre-digitiser integrals (bit-identical to what it digitises, also
    for saturated hits) and differentiable w.r.t. pixels_signals like the reference's; digitize() maps them to ADC."""
    _sim._check_cuda(pixels_signals, "pixels_signals")
    noise = _sim.make_noise(params, pixels_signals.shape[0], noise_rng_key, pixels_signals.device)
    return _AdcValues.apply(pixels_signals, params, noise)


def _prob_forward(params, w, stop_threshold, want_state):
    import ctypes as C
    from . import _lib
    lib = _lib.get_lib()
    pod = _sim.make_pod(params)
    npix, nt = w.shape
    k, npaths = pod.max_adc_values, int(params.fee_paths_scaling)
    dev = w.device
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    with torch.cuda.device(dev):
        lp = torch.empty((npix, k, nt - 1), dtype=torch.float32, device=dev)
        qd = torch.empty((npix, k, nt - 1), dtype=torch.float32, device=dev)
        top = torch.empty((npix, k, npaths), dtype=torch.int32, device=dev) if want_state else None
        state = torch.empty((npix, k, 2 * npaths), dtype=torch.float32, device=dev) if want_state else None
        flags = torch.zeros(k + 1, dtype=torch.int32, device=dev) if want_state else None
        scratch = torch.empty(lib.larnd_prob_fee_scratch_bytes(npix, nt, npaths, k), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.larnd_prob_fee_forward(vp(w), w.stride(0), npix, nt, C.byref(pod), npaths, float(stop_threshold), vp(lp), vp(qd),
                                              vp(top), vp(state), vp(flags), vp(scratch), scratch.numel(), st))
    return lp, qd, top, state, flags, pod


class _ProbFee(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wfs, params, stop_threshold):
        w = wfs.detach()
        if w.dtype != torch.float32 or w.stride(1) != 1:
            w = w.contiguous().float()
        lp, qd, top, state, flags, pod = _prob_forward(params, w, stop_threshold, True)
        ctx.save_for_backward(w, lp, top, state, flags)
        ctx.pod, ctx.npaths = pod, int(params.fee_paths_scaling)
        return lp, qd

    @staticmethod
    def backward(ctx, g_lp, g_q):
        import ctypes as C
        from . import _lib
        w, lp, top, state, flags = ctx.saved_tensors
        lib = _lib.get_lib()
        npix, nt = w.shape
        g_lp = (torch.zeros_like(lp) if g_lp is None else g_lp).contiguous().float()
        g_q = (torch.zeros_like(lp) if g_q is None else g_q).contiguous().float()
        vp = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(w.device):
            g = torch.empty((npix, nt), dtype=torch.float32, device=w.device)
            scratch = torch.empty(lib.larnd_prob_fee_bwd_scratch_bytes(npix, nt, ctx.npaths), dtype=torch.uint8, device=w.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.larnd_prob_fee_backward(vp(w), w.stride(0), npix, nt, C.byref(ctx.pod), ctx.npaths, vp(g_lp), vp(g_q), vp(lp),
                                                   vp(state), vp(top), vp(flags), vp(g), nt, vp(scratch), scratch.numel(), st))
        return g, None, None


def get_adc_values_average_noise_vmap(params, wfs, stop_threshold=1e-9, return_top_ticks=False):
    """(log_prob_distrib (Npix, MAX_ADC_VALUES, Nticks-1), charge_distrib (same shape)) — the noise-averaged beam-search
    front end of the reference (fee_jax.py:390-461), run by the k_prob_* kernels (csrc/prob_fee.cu).  Differentiable
    w.r.t. the waveforms (the beam ticks are the forward's, like lax.stop_gradient in the reference)."""
    _sim._check_cuda(wfs, "wfs")
    if wfs.dim() != 2:
        raise ValueError("wfs must be (Npix, Nticks)")
    if torch.is_grad_enabled() and wfs.requires_grad and not return_top_ticks:
        return _ProbFee.apply(wfs, params, stop_threshold)
    w = wfs.detach()
    if w.dtype != torch.float32 or w.stride(1) != 1:
        w = w.contiguous().float()
    lp, qd, top, _, _, _ = _prob_forward(params, w, stop_threshold, return_top_ticks)
    return (lp, qd, top) if return_top_ticks else (lp, qd)


def get_average_hit_values(ticks_prob, adcs_distrib):
    """Expected tick, expected ADC and lambda = sum_t P(t) per (pixel, hit index) (reference: fee_jax.py:463-481)."""
    lam = ticks_prob.sum(dim=2)
    den = torch.clamp(lam, min=1e-10)
    t = torch.arange(ticks_prob.shape[2], device=ticks_prob.device, dtype=ticks_prob.dtype)
    return (t[None, None, :] * ticks_prob).sum(dim=2) / den, (adcs_distrib * ticks_prob).sum(dim=2) / den, lam
