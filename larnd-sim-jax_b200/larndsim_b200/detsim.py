"""Pixel id packing and geometry helpers: host-side mirror of the reference's ``larndsim.detsim_jax``
(pixel2id :232, id2pixel :265, get_pixel_coordinates :297, get_hit_z :309).  Small elementwise torch ops on
CUDA tensors; the per-segment versions of these run inside the prepare / FEE kernels."""
import numpy as np
import torch

from .consts import get_vdrift


def _borders(params, device):
    return torch.as_tensor(np.asarray(params.tpc_borders, dtype=np.float64), dtype=torch.float32, device=device)


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
