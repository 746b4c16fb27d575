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
