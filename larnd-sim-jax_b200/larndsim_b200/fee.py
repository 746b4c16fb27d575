"""Front-end electronics: host-side mirror of the reference's ``larndsim.fee_jax`` (digitize :57,
get_adc_values :170).  get_adc_values runs the fused sm_100a FEE kernel."""
import torch

from . import sim as _sim


def digitize(params, integral_list):
    """ADC counts from integrated charge — no rounding, ADC stays float (reference: fee_jax.py:57-71)."""
    x = integral_list if torch.is_tensor(integral_list) else torch.as_tensor(integral_list, dtype=torch.float32)
    v = torch.clamp(x * params.GAIN + params.V_PEDESTAL - params.V_CM, min=0)
    return torch.clamp(v * params.ADC_COUNTS / (params.V_REF - params.V_CM), max=params.ADC_COUNTS)


def undigitize(params, adcs):
    return (adcs * (params.V_REF - params.V_CM) / params.ADC_COUNTS + params.V_CM - params.V_PEDESTAL) / params.GAIN


def get_adc_values(params, pixels_signals, noise_rng_key=None):
    """(adc (Npix,10) integrated charge, ticks (Npix,10)) like the reference (fee_jax.py:170-279).  The kernel
    returns digitised ADC; the integral is recovered where the digitiser is invertible (unclipped hits)."""
    noise = _sim.make_noise(params, pixels_signals.shape[0], noise_rng_key, pixels_signals.device)
    upix = torch.zeros(pixels_signals.shape[0], dtype=torch.int32, device=pixels_signals.device)
    fs = _sim.fee_forward(params, pixels_signals, upix, noise, compact=False)
    integral = torch.where(fs.ticks < pixels_signals.shape[1] - 2, undigitize(params, fs.adc), torch.zeros_like(fs.adc))
    return integral, fs.ticks


def get_adc_values_average_noise_vmap(params, wfs, stop_threshold=1e-9, return_top_ticks=False):
    """(log_prob_distrib (Npix, MAX_ADC_VALUES, Nticks-1), charge_distrib (same shape)) — the noise-averaged beam-search
    front end of the reference (fee_jax.py:390-461), run by the k_prob_* kernels (csrc/prob_fee.cu).  Forward only in this
    round: the VJP w.r.t. the waveforms is the next row of the build plan."""
    import ctypes as C
    from . import _lib
    _sim._check_cuda(wfs, "wfs")
    if torch.is_grad_enabled() and wfs.requires_grad:
        raise NotImplementedError("the probabilistic front end is forward-only in this build (no VJP yet)")
    w = wfs.detach()
    if w.dtype != torch.float32 or w.dim() != 2 or w.stride(1) != 1:
        w = w.contiguous().float()
    lib = _lib.get_lib()
    pod = _sim.make_pod(params)
    npix, nt = w.shape
    k, npaths = pod.max_adc_values, int(params.fee_paths_scaling)
    dev = w.device
    with torch.cuda.device(dev):
        lp = torch.empty((npix, k, nt - 1), dtype=torch.float32, device=dev)
        qd = torch.empty((npix, k, nt - 1), dtype=torch.float32, device=dev)
        top = torch.empty((npix, k, npaths), dtype=torch.int32, device=dev) if return_top_ticks else None
        scratch = torch.empty(lib.larnd_prob_fee_scratch_bytes(npix, nt, npaths, k), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.larnd_prob_fee_forward(C.c_void_p(w.data_ptr()), w.stride(0), npix, nt, C.byref(pod), npaths,
                                              float(stop_threshold), C.c_void_p(lp.data_ptr()), C.c_void_p(qd.data_ptr()),
                                              C.c_void_p(top.data_ptr()) if top is not None else C.c_void_p(0),
                                              C.c_void_p(scratch.data_ptr()), scratch.numel(), st))
    return (lp, qd, top) if return_top_ticks else (lp, qd)


def get_average_hit_values(ticks_prob, adcs_distrib):
    """Expected tick, expected ADC and lambda = sum_t P(t) per (pixel, hit index) (reference: fee_jax.py:463-481)."""
    lam = ticks_prob.sum(dim=2)
    den = torch.clamp(lam, min=1e-10)
    t = torch.arange(ticks_prob.shape[2], device=ticks_prob.device, dtype=ticks_prob.dtype)
    return (t[None, None, :] * ticks_prob).sum(dim=2) / den, (adcs_distrib * ticks_prob).sum(dim=2) / den, lam
