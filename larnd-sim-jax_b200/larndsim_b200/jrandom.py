"""jax.random-compatible keys and normals on the device (mirror of the jax.random calls on the reference's hot path:
``key``, ``split``, ``normal``).  Threefry-2x32 on the GPU (csrc/rng.cu): for a given seed the random bits are the ones
jax draws, in either counter layout (``partitionable=True`` = jax_threefry_partitionable, JAX's default since 0.5.0)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

PARTITIONABLE = True  # module default, matches current JAX; set False to reproduce JAX < 0.5 streams


def _arr(k):
    return (C.c_uint32 * 2)(int(k[0]), int(k[1]))


def key(seed):
    """jax.random.key(seed): the two uint32 words (seed >> 32, seed & 0xffffffff)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (seed >> 32, seed & 0xFFFFFFFF)


def split(k, num=2, partitionable=None):
    part = PARTITIONABLE if partitionable is None else partitionable
    out = (C.c_uint32 * (2 * num))()
    _lib.check(_lib.get_lib().larnd_rng_split(_arr(k), num, int(bool(part)), out))
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(num)]


def normal(k, shape, device="cuda", partitionable=None):
    """jax.random.normal(key, shape, float32) as a CUDA tensor."""
    part = PARTITIONABLE if partitionable is None else partitionable
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.LarndError("jrandom.normal needs a CUDA device (larndsim_b200 has no CPU path)")
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.get_lib().larnd_rng_normal(_arr(k), n, int(bool(part)), C.c_void_p(out.data_ptr()), st))
    return out.reshape(shape)


def fee_noise(k, npix, n_adc=10, device="cuda", partitionable=None):
    """Every standard normal get_adc_values(params, wfs, key) draws (fee_jax.py:186,237-255,271) in the FEE kernel's layout."""
    part = PARTITIONABLE if partitionable is None else partitionable
    dev = torch.device(device)
    out = torch.empty(npix * (1 + 3 * n_adc), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.get_lib().larnd_rng_fee_noise(_arr(k), npix, n_adc, int(bool(part)), C.c_void_p(out.data_ptr()), st))
    return out
