"""Consumer-side helpers kept for drop-in use of the simulated hits: adc2charge and the weighted-MMD
``mse_adc`` loss (reference: larndsim.losses_jax :380-383, :14-39, :58-82), as differentiable torch ops.
The losses are O(hits^2) on <= 1e3 hits and are not part of the accelerated path."""
import torch

from . import sim as _sim


def adc2charge(dw, params):
    """ke- from ADC counts; evaluated in double and rounded once, which is what reproduces the goldens' Q column."""
    d = dw.double()
    return ((d / params.ADC_COUNTS * (params.V_REF - params.V_CM) + params.V_CM - params.V_PEDESTAL) / params.GAIN * 1e-3).to(dw.dtype)


def rbf_kernel(x, y, sigma):
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    return torch.exp(-d2 / (2 * sigma ** 2))


def _kernel_sum(x, y, px, py, sigma, block=4096):
    """sum_ij K(x_i, y_j) px_i py_j, blocked over rows so that large hit lists do not materialise an N x M matrix."""
    tot = x.new_zeros(())
    for i in range(0, x.shape[0], block):
        tot = tot + (rbf_kernel(x[i:i + block], y, sigma) * px[i:i + block, None] * py[None, :]).sum()
    return tot


def rbf_field(targets, sources, weights, sigma):
    """(N, 4) field {sum_j w_j K(t, z_j), sum_j w_j K(t, z_j) (z_j - t)} of the weighted sources at the targets, computed by
    the tiled k_rbf_field kernel (csrc/losses.cu); pairs beyond 15 sigma are skipped (exactly zero in float32)."""
    import ctypes as C
    from . import _lib
    lib = _lib.get_lib()
    t = targets.detach().contiguous().float()
    z = sources.detach().contiguous().float()
    w = weights.detach().contiguous().float()
    out = torch.empty((t.shape[0], 4), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        scratch = torch.empty(max(lib.larnd_rbf_field_scratch_bytes(t.shape[0], z.shape[0]), 4), dtype=torch.uint8, device=t.device)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        vp = lambda a: C.c_void_p(a.data_ptr())
        _lib.check(lib.larnd_rbf_field(vp(t), t.shape[0], vp(z), vp(w), z.shape[0], float(sigma), vp(out), vp(scratch), scratch.numel(), st))
    return out


class _KernelSums(torch.autograd.Function):
    """(K_xx, K_yy, K_xy) of the weighted MMD with gradients w.r.t. x and px (the reference side y, py is data)."""

    @staticmethod
    def forward(ctx, x, px, y, py, sigma):
        fxx = rbf_field(x, x, px, sigma)
        fxy = rbf_field(x, y, py, sigma)
        fyy = rbf_field(y, y, py, sigma)
        ctx.save_for_backward(fxx, fxy, px.detach())
        ctx.sigma = float(sigma)
        pxd, pyd = px.detach().float(), py.detach().float()
        return torch.stack([(pxd * fxx[:, 0]).sum(), (pyd * fyy[:, 0]).sum(), (pxd * fxy[:, 0]).sum()])

    @staticmethod
    def backward(ctx, g):
        fxx, fxy, px = ctx.saved_tensors
        inv = 1.0 / (ctx.sigma * ctx.sigma)
        g_x = (2.0 * g[0] * px[:, None] * fxx[:, 1:] + g[2] * px[:, None] * fxy[:, 1:]) * inv
        g_px = 2.0 * g[0] * fxx[:, 0] + g[2] * fxy[:, 0]
        return g_x, g_px, None, None, None


def mmd(x, y, px, py, sigma, reduce=None):
    """Weighted MMD^2.  ``reduce``: optional differentiable sum over ranks applied to the five sums (events never
    interact across ranks: the event*1e5 offset puts them > 1e5 apart), see parallel.allreduce_sum_differentiable."""
    if x.is_cuda and x.shape[-1] == 3:   # tiled CUDA kernel (csrc/losses.cu); the torch expression below serves CPU tensors
        ks = _KernelSums.apply(x, px, y, py, sigma)
        sums = torch.cat([ks, torch.stack([px.sum(), py.sum()])])
    else:
        sums = torch.stack([_kernel_sum(x, x, px, px, sigma), _kernel_sum(y, y, py, py, sigma), _kernel_sum(x, y, px, py, sigma),
                            px.sum(), py.sum()])
    if reduce is not None:
        sums = reduce(sums)
    kxx, kyy, kxy, sx, sy = sums
    return kxx / sx ** 2 + kyy / sy ** 2 - 2 * kxy / (sx * sy)


def mse_adc(params, Q, x, y, z, ticks, hit_prob, event, ref_Q, ref_x, ref_y, ref_z, ref_ticks, ref_hit_prob, ref_event,
            sigma=1, lambda_Q=1, reduce=None):
    w_ref, w = ref_Q * ref_hit_prob, Q * hit_prob
    ref_st = torch.stack((ref_x + ref_event * 1e5, ref_y, ref_z), dim=-1)
    st = torch.stack((x + event * 1e5, y, z), dim=-1)
    mmd_term = mmd(st, ref_st, w, w_ref, sigma, reduce)
    tot_ref, tot = w_ref.sum(), w.sum()
    if reduce is not None:
        tot_ref, tot = reduce(torch.stack([tot_ref, tot]))
    charge_loss = ((tot - tot_ref) / (tot_ref + 1e-6)) ** 2
    aux = {"charge_loss": charge_loss, "mmd_loss_term": mmd_term, "Q": Q, "ref_Q": ref_Q, "ref_hit_prob": ref_hit_prob,
           "hit_prob": hit_prob}
    return mmd_term + lambda_Q * charge_loss, aux


def params_loss(params, response, ref_adcs, ref_x, ref_y, ref_z, ref_ticks, ref_hit_prob, ref_event, tracks, fields,
                rngkey=None, loss_fn=mse_adc, **loss_kwargs):
    """loss(params) through simulate_wfs + simulate_stochastic (reference: losses_jax.py:385-398)."""
    wfs, unique_pixels = _sim.simulate_wfs(params, response, tracks, fields)
    adcs, x, y, z, ticks, hit_prob, event, _ = _sim.simulate_stochastic(params, wfs, unique_pixels, rngseed=rngkey)
    Q, ref_Q = adc2charge(adcs, params), adc2charge(ref_adcs, params)
    return loss_fn(params, Q, x, y, z, ticks, hit_prob, event.float(), ref_Q, ref_x, ref_y, ref_z, ref_ticks, ref_hit_prob,
                   ref_event.float(), **loss_kwargs)
