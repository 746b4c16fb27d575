"""Pixel id packing and geometry helpers: host-side mirror of the reference's ``larndsim.detsim_jax``
(pixel2id :232, id2pixel :265, get_pixel_coordinates :297, get_hit_z :309).  Small elementwise torch ops on
CUDA tensors; the per-segment versions of these run inside the prepare / FEE kernels."""
import numpy as np
import torch

from .consts import get_vdrift


_borders_cache = {}


def _borders(params, device):
    """tpc_borders as a float32 device tensor; uploaded once per (values, device) -- a fit step asks for it every
    iteration and a pageable host-to-device copy synchronises."""
    b = np.asarray(params.tpc_borders, dtype=np.float64)
    key = (b.tobytes(), b.shape, str(device))
    t = _borders_cache.get(key)
    if t is None:
        if len(_borders_cache) > 16:
            _borders_cache.clear()
        t = _borders_cache[key] = torch.as_tensor(b, dtype=torch.float32, device=device)
    return t


def pixel2id(params, pixel_x, pixel_y, pixel_plane, eventID):
    """int32 id = ((event*n_tpc + plane)*ny + y)*nx + x, -1 outside the plane (x64 is never enabled in the
    reference, so its int64 casts are int32)."""
    nx, ny = int(params.n_pixels_x), int(params.n_pixels_y)
    ntpc = int(np.asarray(params.tpc_borders).shape[0])
    outside = (pixel_x >= nx) | (pixel_y >= ny) | (pixel_x < 0) | (pixel_y < 0)
    pid = (eventID.to(torch.int32) * ntpc + pixel_plane.to(torch.int32)) * ny + pixel_y.to(torch.int32)
    pid = pid * nx + pixel_x.to(torch.int32)
    return torch.where(outside, torch.full_like(pid, -1), pid)


def id2pixel(params, pid):
    nx, ny = int(params.n_pixels_x), int(params.n_pixels_y)
    ntpc = int(np.asarray(params.tpc_borders).shape[0])
    fd = lambda a, b: torch.div(a, b, rounding_mode="floor")
    return pid % nx, fd(pid, nx) % ny, fd(pid, nx * ny) % ntpc, fd(pid, nx * ny * ntpc)


def get_pixel_coordinates(params, xpitch, ypitch, plane):
    b = _borders(params, xpitch.device)[plane.long()]
    pitch = float(params.pixel_pitch)
    # fma(index, pitch, border) + pitch/2 (XLA contracts the multiply-add; pinned by the reference goldens)
    px = torch.addcmul(b[..., 0, 0].double(), xpitch.double(), torch.tensor(float(np.float32(pitch)), dtype=torch.float64, device=b.device)).float() + pitch / 2
    py = torch.addcmul(b[..., 1, 0].double(), ypitch.double(), torch.tensor(float(np.float32(pitch)), dtype=torch.float64, device=b.device)).float() + pitch / 2
    return torch.stack([px, py], dim=-1)


def get_hit_z(params, ticks, plane, fixed_v=False):
    """z of a hit from its tick: z_anode + tick * t_sampling * v * sign(z_cathode - z_anode).  Differentiable
    w.r.t. eField through get_vdrift when eField is a Params leaf."""
    b = _borders(params, ticks.device)[plane.long()]
    z_anode, z_high = b[..., 2, 0], b[..., 2, 1]
    v = params.vdrift_static if fixed_v else get_vdrift(params)
    if torch.is_tensor(v):
        v = v.to(ticks.device)
    tv = float(np.float32(params.t_sampling)) * v      # XLA folds the two scalars first (pinned by the goldens' pix_z)
    if not torch.is_tensor(tv):
        tv = float(np.float32(tv))
    return z_anode + ticks * tv * torch.sign(z_high - z_anode)


# ------------------------------------------------------------------------------------------ id packing limits / validators
# Host-side guards the reference runs on every batch before simulating (detsim_jax.py:26-152, called from
# optimize/simulate.py:111-113 and dataio): same names, same exceptions (ValueError / OverflowError).
_I32_MAX = 2 ** 31 - 1
_I64_MAX = np.iinfo(np.int64).max


def _ntpc(params):
    return int(np.asarray(params.tpc_borders).shape[0])


def _pixel_id_stride(params):
    return int(params.n_pixels_x) * int(params.n_pixels_y) * _ntpc(params)


def _bin_id_stride(params):
    nb = int(params.nb_sampling_bins_per_pixel)
    return int(params.n_pixels_x) * nb * int(params.n_pixels_y) * nb * _ntpc(params)


def max_safe_event_id_for_pixel_packing(params):
    s = _pixel_id_stride(params)
    return (_I64_MAX - (s - 1)) // s


def max_safe_event_id_for_bin_packing(params):
    s = _bin_id_stride(params)
    return (_I64_MAX - (s - 1)) // s


def max_possible_pixel_id(params, max_event_id):
    return (int(max_event_id) + 1) * _pixel_id_stride(params) - 1


def max_possible_bin_id(params, max_event_id):
    return (int(max_event_id) + 1) * _bin_id_stride(params) - 1


def _as_i64(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.int64)


def _reject_below_minus_one(ids, what, context):
    bad = ids < -1
    if bad.any():
        raise ValueError("%s found %s below -1: %s" % (context, what, np.unique(ids[bad])[:16].tolist()))


def validate_event_ids_for_packing(params, event_ids, kind="pixel", context=""):
    ids = _as_i64(event_ids)
    if ids.size == 0:
        return
    _reject_below_minus_one(ids, "eventID values", context)
    if not (ids >= 0).any():
        return
    if kind not in ("pixel", "bin"):
        raise ValueError("Unknown packing kind '%s'" % kind)
    top = int(ids.max())
    limit = max_safe_event_id_for_pixel_packing(params) if kind == "pixel" else max_safe_event_id_for_bin_packing(params)
    biggest = max_possible_pixel_id(params, top) if kind == "pixel" else max_possible_bin_id(params, top)
    if top > limit:
        raise OverflowError("%s eventID %d exceeds the int64-safe limit %d for %s packing" % (context, top, limit, kind))
    if biggest > _I64_MAX:
        raise OverflowError("%s maximum packed %s id %d exceeds int64 max" % (context, kind, biggest))
    if kind == "pixel" and biggest > _I32_MAX:
        # the kernels pack pixel ids in int32 — and so does the reference as it is run (jax_enable_x64 is never set, so its
        # astype(jnp.int64) yields int32 and larger ids wrap silently, detsim_jax.py:241-244); wrapping is refused here
        raise OverflowError("%s maximum packed pixel id %d exceeds the int32 range ids are packed in (eventID must stay <= %d)"
                            % (context, biggest, (_I32_MAX + 1) // _pixel_id_stride(params) - 1))


def validate_packed_ids_for_decoding(params, packed_ids, kind="pixel", context=""):
    ids = _as_i64(packed_ids)
    if ids.size == 0:
        return
    _reject_below_minus_one(ids, "packed ids", context)
    ok = ids >= 0
    if not ok.any():
        return
    if kind not in ("pixel", "bin"):
        raise ValueError("Unknown decoding kind '%s'" % kind)
    stride = _pixel_id_stride(params) if kind == "pixel" else _bin_id_stride(params)
    limit = max_safe_event_id_for_pixel_packing(params) if kind == "pixel" else max_safe_event_id_for_bin_packing(params)
    top = int((ids[ok] // stride).max())
    if top > limit:
        raise OverflowError("%s decoded eventID %d exceeds the int64-safe limit %d for %s packing" % (context, top, limit, kind))


def validate_local_event_ids(event_ids, context=""):
    """Event ids of a batch must be exactly {-1} U {0..N-1} (batch-local namespace, dataio.py:47-61)."""
    ids = _as_i64(event_ids)
    if ids.size == 0:
        return
    _reject_below_minus_one(ids, "eventID values (non-local namespace)", context)
    valid = np.unique(ids[ids >= 0])
    if valid.size and not np.array_equal(valid, np.arange(valid.size)):
        raise ValueError("%s eventID values do not form contiguous local namespace [0, %d]. Found unique IDs: %s"
                         % (context, valid.size - 1, valid.tolist()))


# ------------------------------------------------------------------------------------------ bins, diffusion weights, electrons
def bin2id(params, bin_x, bin_y, pixel_plane, eventID):
    """Packed id of a sub-pixel bin, -1 outside the plane (reference: detsim_jax.py:247-262)."""
    nb = int(params.nb_sampling_bins_per_pixel)
    nbx, nby = int(params.n_pixels_x) * nb, int(params.n_pixels_y) * nb
    outside = (bin_x >= nbx) | (bin_y >= nby) | (bin_x < 0) | (bin_y < 0)
    bid = bin_x + nbx * (bin_y + nby * (pixel_plane + _ntpc(params) * eventID))
    return torch.where(outside, torch.full_like(bid, -1), bid)


def id2bin(params, bin_id):
    nb = int(params.nb_sampling_bins_per_pixel)
    nbx, nby = int(params.n_pixels_x) * nb, int(params.n_pixels_y) * nb
    fd = lambda a, b: torch.div(a, b, rounding_mode="floor")
    return bin_id % nbx, fd(bin_id, nbx) % nby, fd(bin_id, nbx * nby) % _ntpc(params), fd(bin_id, nbx * nby * _ntpc(params))


def _plane_borders(params, electrons, fields):
    plane = electrons[:, tuple(fields).index("pixel_plane")].long()
    return _borders(params, electrons.device)[plane], plane


def get_bin_shifts(params, electrons, fields):
    """(N, 2) int32 sub-pixel bin indices ((x - x0_tpc) // (pitch / bins_per_pixel)), reference: detsim_jax.py:494-512."""
    f = tuple(fields)
    b, _ = _plane_borders(params, electrons, fields)
    w = float(np.float32(params.pixel_pitch / params.nb_sampling_bins_per_pixel))
    bx = torch.floor_divide(electrons[:, f.index("x")] - b[:, 0, 0], w)
    by = torch.floor_divide(electrons[:, f.index("y")] - b[:, 1, 0], w)
    return torch.stack([bx, by], dim=1).to(torch.int32)


def get_pixels(params, electrons, fields):
    """Packed pixel ids of the electrons' own pixel and its (2n+1)^2 neighbourhood (reference: detsim_jax.py:478-491)."""
    f = tuple(fields)
    n = int(params.number_pix_neighbors)
    b, plane = _plane_borders(params, electrons, fields)
    pitch = float(np.float32(params.pixel_pitch))
    px = torch.floor_divide(electrons[:, f.index("x")] - b[:, 0, 0], pitch).to(torch.int32)
    py = torch.floor_divide(electrons[:, f.index("y")] - b[:, 1, 0], pitch).to(torch.int32)
    g = torch.arange(-n, n + 1, device=electrons.device, dtype=torch.int32)
    ev = electrons[:, f.index("eventID")].to(torch.int32)
    return pixel2id(params, px[:, None, None] + g[None, :, None], py[:, None, None] + g[None, None, :], plane[:, None, None].to(torch.int32),
                    ev[:, None, None])


def density_2d(bins, x0, y0, sigma):
    """(N, 5, 5) transverse-diffusion weights: per axis the Gaussian integral over the bins with the outer edges forced
    to -1/+1 (tails folded in), outer product (reference: detsim_jax.py:332-354)."""
    def axis(edges, c):
        e = torch.erf((edges[None, :] - c[:, None]) / (np.float32(np.sqrt(2)) * sigma[:, None]))
        e = torch.cat([torch.full_like(e[:, :1], -1.0), e[:, 1:-1], torch.full_like(e[:, :1], 1.0)], dim=1)
        return 0.5 * (e[:, 1:] - e[:, :-1])
    return axis(bins, x0)[:, :, None] * axis(bins, y0)[:, None, :]


def generate_electrons(tracks, fields, rngkey, apply_long_diffusion=True):
    """Gaussian smearing of (x, y[, z]) by (tran_diff, tran_diff, long_diff) with random.normal(rngkey, (N, 3))
    (reference: detsim_jax.py:377-400); ``rngkey`` = two uint32 words (jrandom.key / jrandom.split)."""
    from . import jrandom
    f = tuple(fields)
    rnd = jrandom.normal(rngkey, (tracks.shape[0], 3), tracks.device)
    out = tracks.clone()
    out[:, f.index("x")] += rnd[:, 0] * tracks[:, f.index("tran_diff")]
    out[:, f.index("y")] += rnd[:, 1] * tracks[:, f.index("tran_diff")]
    if apply_long_diffusion:
        out[:, f.index("z")] += rnd[:, 2] * tracks[:, f.index("long_diff")]
    return out


def apply_tran_diff(params, electrons, fields):
    """The reference's deterministic (non-MC) transverse split reads params.tran_diff_bin_edges, which no loader ever sets
    (consts_jax.py:160,244-298): it fails there with a TypeError on None; the same condition is reported here."""
    if getattr(params, "tran_diff_bin_edges", None) is None:
        raise TypeError("params.tran_diff_bin_edges is None: the non-MC transverse split is unusable in the reference as well "
                        "(every parametrized script passes --mc_diff)")
    raise NotImplementedError("apply_tran_diff with explicit bin edges is not part of the accelerated path")


# stage-by-stage MC-current operators with the reference's argument lists (detsim_jax.py:619-639, 209-228)
from .stream_ops import accumulate_signals, accumulate_signals_parametrized, current_lut, current_mc  # noqa: E402,F401
