"""TEST INFRASTRUCTURE (oracle): numpy restatement of the noise-averaged ("probabilistic") front end of the reference —
fee_jax.py:12-53 (_soft_max, _soft_where, log_diff_ndtr), :334-388 (_find_one_hit_step), :390-461
(get_adc_values_average_noise_vmap), :463-481 (get_average_hit_values), sim_jax.py:772-812 (simulate_probabilistic).

jax.scipy.special.log_ndtr, jax.nn.softplus / sigmoid / logsumexp, lax.cummax and lax.top_k are third-party (jax,
unpinned in the reference, absent here); they are restated from their published definitions: log_ndtr is the
TensorFlow-Probability three-segment formula jax uses (float32: asymptotic series of order 3 below -10, log(ndtr) up to
5, -ndtr(-x) above), softplus = logaddexp(x, 0), logsumexp is max-shifted, top_k returns the k largest in descending
order with the lower index first among equals.  Parity of this part is UNPINNED (no golden output of the reference
exists for it, and jax cannot run here); the CUDA kernels are checked against this restatement, in float32 for values and
in float64 by central differences for gradients."""
import math

import numpy as np
from scipy import special as sps

from . import larnd_oracle as lo


def softplus(x):
    return np.logaddexp(x, x.dtype.type(0))


def sigmoid(x):
    return (x.dtype.type(1) / (x.dtype.type(1) + np.exp(-x))).astype(x.dtype)


def soft_max(x, lo_, sharpness):
    dt = x.dtype.type
    return (dt(lo_) + softplus(((x - dt(lo_)) * dt(sharpness)).astype(x.dtype)) / dt(sharpness)).astype(x.dtype)


def soft_where(cond, tv, fv, sharpness):
    dt = cond.dtype.type
    w = sigmoid((cond * dt(sharpness)).astype(cond.dtype))
    return (w * tv + (dt(1) - w) * dt(fv)).astype(cond.dtype)


def _ndtr(x):
    dt = x.dtype.type
    hs2 = dt(0.5) * np.sqrt(dt(2.0))
    w = (x * hs2).astype(x.dtype)
    z = np.abs(w)
    y = np.where(z < hs2, dt(1) + sps.erf(w).astype(x.dtype), np.where(w > 0, dt(2) - sps.erfc(z).astype(x.dtype), sps.erfc(z).astype(x.dtype)))
    return (dt(0.5) * y).astype(x.dtype)


def log_ndtr(x):
    """jax.scipy.special.log_ndtr (series_order = 3)."""
    x = np.asarray(x)
    dt = x.dtype.type
    lower, upper = (dt(-20), dt(8)) if x.dtype == np.float64 else (dt(-10), dt(5))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        xl = np.minimum(x, lower)
        x2 = xl * xl
        log_scale = -dt(0.5) * x2 - np.log(-xl) - dt(0.5 * math.log(2.0 * math.pi))
        even = dt(3) / (x2 * x2)                 # n = 2: 3!! / x^4
        odd = dt(1) / x2 + dt(15) / (x2 * x2 * x2)  # n = 1: 1!! / x^2, n = 3: 5!! / x^6
        low = log_scale + np.log(dt(1) + even - odd)
        mid = np.log(_ndtr(np.maximum(x, lower)))
        high = -_ndtr(-x)
    return np.where(x > upper, high, np.where(x > lower, mid, low)).astype(x.dtype)


def log_diff_ndtr(a, b):
    dt = a.dtype.type
    la, lb = log_ndtr(a), log_ndtr(b)
    safe_diff = np.where(a > b, lb - la, dt(-1.0))
    neg_expm1 = -np.expm1(safe_diff)
    with np.errstate(over="ignore"):
        neg_safe = soft_max(neg_expm1.astype(a.dtype), 1e-30, 1e10)
    log_term = np.log(neg_safe)
    log_term_safe = soft_max(log_term.astype(a.dtype), -100.0, 10.0)
    log_prob = (la + log_term_safe).astype(a.dtype)
    return soft_where((a - b).astype(a.dtype), log_prob, -1000.0, 1000.0)


def logsumexp(a, axis):
    amax = np.max(a, axis=axis, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, a.dtype.type(0))
    out = np.log(np.sum(np.exp(a - amax), axis=axis, keepdims=True)) + amax
    return np.squeeze(out, axis=axis).astype(a.dtype)


def top_k_indices(v, k):
    """lax.top_k indices: k largest, descending, lower index first among equals."""
    return np.argsort(-v, kind="stable")[:k]


def find_one_hit_step(q_sum, prev_charges, prev_log_prob, sigma, threshold, interval, nvalues, force_top=None):
    """One beam-search step for ONE pixel (fee_jax.py:334-388).  q_sum (Nt,), prev_* (nvalues,)."""
    dt = q_sum.dtype.type
    nt = q_sum.shape[0]
    z = dt(1.0) / dt(sigma)
    thr = dt(threshold)
    shifted = np.clip(np.arange(nt - 1) + interval + 1, 0, nt - 1)
    loc = (q_sum[None, :] - prev_charges[:, None]).astype(q_sum.dtype)
    qmf = np.maximum.accumulate(loc, axis=1)
    log_guess = log_diff_ndtr(((qmf[:, 1:] - thr) * z).astype(q_sum.dtype), ((qmf[:, :-1] - thr) * z).astype(q_sum.dtype))
    lpe = log_diff_ndtr(((loc[:, shifted] - thr) * z).astype(q_sum.dtype), ((loc[:, :-1] - thr) * z).astype(q_sum.dtype))
    lpe = np.minimum(lpe, log_guess)
    lpe = soft_max(lpe, -1000.0, 1.0)
    esp = (q_sum[shifted] + thr - dt(0.5) * (q_sum[1:] + q_sum[:-1])).astype(q_sum.dtype)
    lpe = soft_where(np.broadcast_to((esp - thr).astype(q_sum.dtype), lpe.shape).copy(), lpe, -1000.0, 10.0)
    lpd = (lpe + prev_log_prob[:, None]).astype(q_sum.dtype)
    log_hit = logsumexp(lpd, 0)
    log_tot = logsumexp((log_guess + prev_log_prob[:, None]).astype(q_sum.dtype), 0)
    next_q = loc[:, np.clip(shifted + 1, 0, nt - 1)]
    mfs = np.maximum.accumulate(loc[:, ::-1], axis=1)[:, ::-1]
    fend = np.clip(shifted + interval + 1, 0, nt - 1)
    lf = log_ndtr(((mfs[:, fend] - next_q - thr) * z).astype(q_sum.dtype))
    lsel = logsumexp((log_guess + lf + prev_log_prob[:, None]).astype(q_sum.dtype), 0)
    top = top_k_indices(lsel, nvalues) if force_top is None else np.asarray(force_top)
    new_lp = log_tot[top]
    best_next = np.clip(shifted[top] + 1, 0, nt - 1)
    return (q_sum[best_next], new_lp), (log_hit, esp), top


def get_adc_values_average_noise(params, wfs, stop_threshold=1e-9, dt=np.float32, return_state=False, force_tops=None,
                                 force_active=None):
    """(log_prob_distrib (Npix, MAX_ADC, Nt-1), charge_distrib (Npix, MAX_ADC, Nt-1)) — fee_jax.py:390-461.
    ``force_tops`` / ``force_active`` pin the discrete choices (beam ticks, global stop flags per step) so that finite
    differences of a float64 evaluation differentiate the same smooth branch as a float32 run."""
    wfs = np.asarray(wfs, dtype=dt)
    npix, nt = wfs.shape
    nv = int(params.fee_paths_scaling)
    interval = lo.hold_interval(params)
    nsteps = int(params.MAX_ADC_VALUES)
    if dt == np.float32:
        q_sum = np.empty_like(wfs)
        acc = np.zeros(npix, dt)
        q = (wfs * dt(params.t_sampling)).astype(dt)
        for t in range(nt):          # strictly left-to-right float32 running sum (like the stochastic FEE kernel)
            acc = (acc + q[:, t]).astype(dt)
            q_sum[:, t] = acc
    else:
        q_sum = np.cumsum(wfs * dt(params.t_sampling), axis=1)
    charges = np.zeros((npix, nv), dt)
    lps = np.full((npix, nv), -1000.0, dt)
    lps[:, 0] = 0
    active = True
    out_lp = np.empty((npix, nsteps, nt - 1), dt)
    out_q = np.empty((npix, nsteps, nt - 1), dt)
    tops = np.zeros((npix, nsteps, nv), np.int64)
    for s in range(nsteps):
        if active:
            tot = np.empty(npix, dt)
            for p in range(npix):
                (c_new, lp_new), (lh, esp), top = find_one_hit_step(q_sum[p], charges[p], lps[p], params.RESET_NOISE_CHARGE,
                                                                   params.DISCRIMINATION_THRESHOLD, interval, nv,
                                                                   None if force_tops is None else force_tops[p, s])
                charges[p], lps[p] = c_new, lp_new
                out_lp[p, s], out_q[p, s] = lh, esp
                tops[p, s] = top
                tot[p] = logsumexp(lp_new, 0)
            active = bool(np.any(tot > dt(math.log(stop_threshold)))) if force_active is None else bool(force_active[s + 1])
        else:
            out_lp[:, s] = -1000.0
            out_q[:, s] = 0.0
    if return_state:
        return out_lp, out_q, tops
    return out_lp, out_q


def get_average_hit_values(ticks_prob, adcs_distrib):
    """fee_jax.py:463-481: per (pixel, hit index) expected tick, expected ADC and lambda = sum of probabilities."""
    dt = ticks_prob.dtype.type
    lam = ticks_prob.sum(axis=2)
    den = np.maximum(lam, dt(1e-10))
    t = np.arange(ticks_prob.shape[2]).astype(ticks_prob.dtype)
    return (t[None, None, :] * ticks_prob).sum(axis=2) / den, (adcs_distrib * ticks_prob).sum(axis=2) / den, lam


def simulate_probabilistic(params, wfs, unique_pixels, dt=np.float32):
    """sim_jax.py:772-812: (adcs_distrib, pixel_x, pixel_y, ticks_prob (log), event)."""
    lp, qd = get_adc_values_average_noise(params, wfs, dt=dt)
    adcs = lo.digitize(params, qd, dt)
    px, py, plane, event = lo.id2pixel(params, unique_pixels)
    coords = lo.get_pixel_coordinates(params, px, py, plane, dt)
    return adcs, coords[:, 0], coords[:, 1], lp, event
