"""Parameter container, detector/pixel YAML loader, drift-velocity model and LUT loader.

Host-side mirror of the reference's ``larndsim.consts_jax`` (same names and call signatures):
``RecombinationMode`` (consts_jax.py:20), ``build_params_class`` (:162), ``get_vdrift`` (:193),
``load_detector_properties`` (:220), ``load_lut`` (:387).  The reference's container is a flax struct whose
``params_with_grad`` fields are traced JAX scalars; here those fields hold 0-d ``torch`` tensors with
``requires_grad=True`` and every other field is a plain Python value.  ``Params.replace(**kw)`` is the only
mutation API, as in the reference.
"""
import json
import math
from enum import Enum

import numpy as np
import torch

from . import _lib


class RecombinationMode(Enum):
    BOX = 1
    BIRKS = 2
    ELLIPSOID = 3


_DEFAULTS = dict(
    eField=0.50, Ab=0.8, kb=0.0486, vdrift=0.1648, vdrift_static=0.159645, lifetime=2.2e3,
    long_diff=4.0e-6, tran_diff=8.8e-6, shift_x=0.0, shift_y=0.0, shift_z=0.0,
    recombination_mode=RecombinationMode.BIRKS, lArDensity=1.38, alpha=0.93, beta=0.212, R_param=1.25,
    MeVToElectrons=4.237e4, temperature=87.17, max_active_pixels=0, max_radius=0, min_step_size=0.001,
    time_max=0, time_window=189.1, e_charge=1.602e-19, t_sampling=0.1, time_padding=190,
    time_interval=(0, 200), drift_length=0, response_bin_size=0.04434, number_pix_neighbors=1,
    electron_sampling_resolution=0.001, signal_length=150, MAX_ADC_VALUES=10, DISCRIMINATION_THRESHOLD=7e3,
    ADC_HOLD_DELAY=15, CLOCK_CYCLE=0.1, GAIN=4e-3, V_CM=288, V_REF=1300, V_PEDESTAL=580, ADC_COUNTS=2 ** 8,
    RESET_NOISE_CHARGE=900, UNCORRELATED_NOISE_CHARGE=500,
    ELECTRON_MOBILITY_PARAMS=(551.6, 7158.3, 4440.43, 4.29, 43.63, 0.2053), size_margin=2e-2,
    diffusion_in_current_sim=True, mc_diff=False, response_full_drift_t=190.61638, nb_tran_diff_bins=5,
    nb_sampling_bins_per_pixel=10, long_diff_template=None, long_diff_extent=20, roi_threshold=0.01,
    roi_split_length=400, fee_paths_scaling=20, hit_prob_threshold=1e-5, tran_diff_bin_edges=None,
    tpc_borders=None, pixel_pitch=0.4434, n_pixels_x=0, n_pixels_y=0,
)


def linspace_f32(start, stop, num):
    """jnp.linspace in float32: start*(1-s) + stop*s, s = i/(num-1), exact end point."""
    f = np.float32
    div = num - 1
    step = (np.arange(div, dtype=np.float32) / f(div)).astype(np.float32)
    body = f(start) * (f(1) - step) + f(stop) * step
    return np.concatenate([body, np.array([stop], dtype=np.float32)]).astype(np.float32)


class _ParamsBase:
    """Attribute container; subclasses are created by build_params_class."""
    _grad_fields = ()
    _tensorize = True   # False: the fields of _grad_fields stay plain floats (host-side value carriers of the fused fit step)

    def __init__(self, **kw):
        unknown = set(kw) - set(_DEFAULTS)
        if unknown:
            raise TypeError("unknown parameter(s): %s" % sorted(unknown))
        vals = dict(_DEFAULTS)
        vals.update(kw)
        if vals["long_diff_template"] is None:
            vals["long_diff_template"] = linspace_f32(0.001, 10, 100)
        for k, v in vals.items():
            if k in self._grad_fields and self._tensorize and not torch.is_tensor(v):
                v = torch.tensor(float(v), dtype=torch.float32, requires_grad=True)
            object.__setattr__(self, k, v)

    def __setattr__(self, k, v):
        raise AttributeError("Params is immutable; use .replace(**kw)")

    def replace(self, **kw):
        d = {k: getattr(self, k) for k in _DEFAULTS}
        d.update(kw)
        return type(self)(**d)

    def value(self, name):
        """Python float value of a (possibly differentiable) scalar field.  The tensor-valued fields of a Params object
        are fetched together, one device-to-host copy (one synchronisation) per object, and the snapshot is kept while
        none of the tensors is modified in place (their version counters are part of the key) -- a parameter block is
        built several times per fit step (waveforms, front end, loss) and would otherwise synchronise per field."""
        v = getattr(self, name)
        if not torch.is_tensor(v):
            return float(v)
        tens = [(k, getattr(self, k)) for k in _DEFAULTS if torch.is_tensor(getattr(self, k))]
        key = tuple((id(t), t._version) for _, t in tens)
        cache = self.__dict__.get("_host_snapshot")
        if cache is None or cache[0] != key:
            host = {}
            by_dev = {}
            for k, t in tens:
                if t.numel() != 1:
                    continue  # array-valued tensor fields are not scalars of the parameter block
                by_dev.setdefault(t.device, []).append((k, t))
            for dev, items in by_dev.items():
                vals = torch.stack([t.detach().reshape(()).to(torch.float64) for _, t in items]).cpu().tolist()
                host.update({k: x for (k, _), x in zip(items, vals)})
            cache = (key, host)
            object.__setattr__(self, "_host_snapshot", cache)
        return cache[1][name] if name in cache[1] else float(v.detach())

    def grad_leaves(self):
        """[(name, tensor)] of differentiable fields, in the C ABI's LARND_P_* order."""
        return [(n, getattr(self, n)) for n in _lib.PARAM_ORDER
                if torch.is_tensor(getattr(self, n)) and getattr(self, n).requires_grad]

    def __repr__(self):
        return "Params(grad=%s)" % (list(self._grad_fields),)


def build_params_class(params_with_grad):
    """Returns a Params class whose ``params_with_grad`` fields are differentiable leaves
    (reference: consts_jax.py:162-190)."""
    for n in params_with_grad:
        if n not in _DEFAULTS:
            raise ValueError("unknown parameter '%s'" % n)
    return type("Params", (_ParamsBase,), {"_grad_fields": tuple(params_with_grad)})


_float_classes = {}


def float_params_class(cls):
    """The Params class ``cls`` with the same fitted-field list (what decides e.g. the float32 evaluation of the drift
    velocity) whose fitted fields are NOT turned into torch leaves: a plain-float value carrier for host-side loops that own
    the parameter values themselves (fit.FusedFitStep builds one per step; tensor leaves there cost ~0.1 ms of host time)."""
    c = _float_classes.get(cls)
    if c is None:
        c = _float_classes[cls] = type(cls.__name__ + "Floats", (cls,), {"_tensorize": False})
    return c


def _mobility(e, temperature, mob):
    a0, a1, a2, a3, a4, a5 = mob
    num = a0 + a1 * e + a2 * e ** 1.5 + a3 * e ** 2.5
    den = 1 + (a1 / a0) * e + a4 * e ** 2 + a5 * e ** 3
    return num / den * (temperature / 89) ** -1.5 / 1000


def get_vdrift(params):
    """Drift velocity in cm/us from eField and temperature (reference: consts_jax.py:193-216).
    Returns a Python float for a static eField and a differentiable torch scalar otherwise."""
    e = params.eField
    return _mobility(e, params.temperature, params.ELECTRON_MOBILITY_PARAMS) * e


def _vdrift_f32(e, temperature, mob):
    """get_vdrift evaluated op by op in float32, the way the reference's jitted function runs when eField is a traced
    float32 leaf (consts_jax.py:208-216; Python-float constants are weak-typed and rounded to float32)."""
    f = np.float32
    a0, a1, a2, a3, a4, a5 = (f(a) for a in mob)
    e = f(e)
    num = a0 + a1 * e + a2 * np.power(e, f(1.5)) + a3 * np.power(e, f(2.5))
    den = f(1) + f(mob[1] / mob[0]) * e + a4 * np.power(e, f(2)) + a5 * np.power(e, f(3))
    mu = num / den * f((temperature / 89) ** -1.5) / f(1000)
    return float(f(mu * e))


def vdrift_and_derivative(params):
    """(v, dv/dE) for the kernels' parameter block.  v follows the reference's arithmetic: Python doubles rounded once when
    eField is static, float32 op by op when it is a fitted leaf (a traced float32 scalar in the reference) — the tick,
    fraction and template-index boundaries depend on the last bit of v.  dv/dE is the closed form, in double."""
    e = params.value("eField")
    a0, a1, a2, a3, a4, a5 = params.ELECTRON_MOBILITY_PARAMS
    tc = (params.temperature / 89) ** -1.5 / 1000
    num = a0 + a1 * e + a2 * e ** 1.5 + a3 * e ** 2.5
    den = 1 + (a1 / a0) * e + a4 * e ** 2 + a5 * e ** 3
    dnum = a1 + 1.5 * a2 * e ** 0.5 + 2.5 * a3 * e ** 1.5
    dden = a1 / a0 + 2 * a4 * e + 3 * a5 * e ** 2
    mu, dmu = num / den * tc, (dnum * den - num * dden) / den ** 2 * tc
    v = mu * e
    if "eField" in getattr(params, "_grad_fields", ()):
        v = _vdrift_f32(e, params.temperature, params.ELECTRON_MOBILITY_PARAMS)
    return v, mu + e * dmu


def _geometry_from_yaml(detprop_file, pixel_file):
    import yaml

    vals = {}
    with open(detprop_file) as fh:
        det = yaml.safe_load(fh)
    centers = np.zeros((2, 3))
    for key, value in det.items():
        if key == "tpc_centers":
            centers = np.array(value, dtype=np.float64)
        elif key in _DEFAULTS:
            vals[key] = tuple(value) if isinstance(value, list) else value
        else:
            raise ValueError("Key '%s' in detector properties file is not recognized." % key)
    centers = centers[:, [2, 1, 0]]  # the loaders swap x and z
    with open(pixel_file) as fh:
        layout = yaml.safe_load(fh)
    cm = 0.1
    pitch = layout["pixel_pitch"] * cm
    chan = np.array(list(layout["chip_channel_to_position"].values()))
    xs, ys = chan[:, 0] * pitch, chan[:, 1] * pitch
    half_x, half_y = (xs.max() + pitch) / 2, (ys.max() + pitch) / 2
    tile_tpc = {t: idx[0] for t, idx in layout["tile_indeces"].items()}
    tpcs = sorted(set(tile_tpc.values()))
    if "drift_length" in layout:
        drift_length = layout["drift_length"] * cm
    else:
        anodes = np.array(list(layout["tile_positions"].values()))[:, 0] * cm
        drift_length = 0.5 * (anodes.max() - anodes.min()) * cm
    borders = np.zeros((len(tpcs), 3, 2))
    for i, tpc in enumerate(tpcs):
        tiles = [t for t, k in tile_tpc.items() if k == tpc]
        direction = {layout["tile_orientations"][t][0] for t in tiles}
        if len(direction) != 1:
            raise ValueError("Tiles in same anode plane have different drift directions.")
        direction = direction.pop()
        if direction not in (1, -1):
            raise ValueError("Cathode direction should be either 1 or -1.")
        pos = np.array([layout["tile_positions"][t] for t in tiles], dtype=np.float64) * cm
        borders[i, 0] = pos[:, 2].min() - half_x + centers[i][0], pos[:, 2].max() + half_x + centers[i][0]
        borders[i, 1] = pos[:, 1].min() - half_y + centers[i][1], pos[:, 1].max() + half_y + centers[i][1]
        borders[i, 2] = pos[:, 0].min() + centers[i][2], pos[:, 0].max() + drift_length * direction + centers[i][2]
    vals.update(pixel_pitch=pitch, drift_length=drift_length, tpc_borders=borders,
                n_pixels_x=len(np.unique(xs)) * 2, n_pixels_y=len(np.unique(ys)) * 4)
    return vals


def load_detector_properties(params_cls, detprop_file, pixel_file):
    """Params from the detector-properties and pixel-layout YAML files (reference: consts_jax.py:220-385)."""
    return params_cls(**_geometry_from_yaml(detprop_file, pixel_file))


def load_geometry_json(params_cls, path):
    """Params from a JSON dump of the derived geometry (tests/golden/module0_geometry.json)."""
    with open(path) as fh:
        d = json.load(fh)
    d["tpc_borders"] = np.array(d["tpc_borders"], dtype=np.float64)
    d["time_interval"] = tuple(d["time_interval"])
    return params_cls(**d)


def gaussian_bank_kernels(params):
    ext = int(params.long_diff_extent)
    x = np.arange(-ext, ext + 1, dtype=np.float32)
    sig = np.asarray(params.long_diff_template, dtype=np.float32)[:, None]
    # jax.scipy.stats.norm.pdf = exp(logpdf), logpdf = (log(2 pi scale^2) + x^2 / scale^2) / -2, float32 operation by operation
    s2 = (sig * sig).astype(np.float32)
    g = np.exp(((np.log(np.float32(2 * math.pi) * s2) + (x[None, :] * x[None, :]) / s2) / np.float32(-2)).astype(np.float32)).astype(np.float32)
    return (g / g.sum(axis=1, keepdims=True)).astype(np.float32)


def build_response_template(response, params, device="cuda"):
    """Template bank (n_templates, Nx, Ny, Nt): row t = response convolved ('same') with the normalised
    Gaussian of width long_diff_template[t] ticks; row 0 = the raw response (reference: consts_jax.py:427-447).
    Built on the device by k_build_bank (csrc/lut_tables.cu)."""
    import ctypes as C
    from . import _lib
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.LarndError("build_response_template needs a CUDA device (larndsim_b200 has no CPU path)")
    resp = torch.as_tensor(np.ascontiguousarray(response, dtype=np.float32), device=dev)
    if resp.dim() != 3:
        raise ValueError("response must have shape (Nx, Ny, Nt)")
    nx, ny, nt = resp.shape
    g = torch.as_tensor(gaussian_bank_kernels(params), device=dev).contiguous()
    ntpl, taps = g.shape
    bank = torch.empty((ntpl, nx, ny, nt), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.get_lib().larnd_build_bank(C.c_void_p(resp.data_ptr()), nx, ny, nt, C.c_void_p(g.data_ptr()), ntpl, taps,
                                                   C.c_void_p(bank.data_ptr()), st))
    return bank


def _bank_cache_key(resp, params):
    import hashlib
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(resp, dtype=np.float32).tobytes())
    h.update(np.asarray(params.long_diff_template, dtype=np.float32).tobytes())
    h.update(("extent=%d;v=2" % int(params.long_diff_extent)).encode())
    return h.hexdigest()


def load_lut(lut_file, params, device="cuda", cache_dir=None):
    """(response_template, params') from a .npy/.npz response file (reference: consts_jax.py:387-449).

    ``cache_dir``: on-disk cache of the convolved bank, keyed by a hash of (response samples, template grid, extent).  The
    reference rebuilds the bank on every start (100 x 2025 ``jnp.convolve`` calls); here the build is one ~2 ms kernel, so
    the cache only saves the host-side read of the response — it exists for drivers that share one bank between many
    processes (e.g. one per GPU) and is validated by shape on load (a stale or truncated file is rebuilt)."""
    resp = np.load(lut_file)
    new_params = params.replace()
    if isinstance(resp, np.lib.npyio.NpzFile):
        if "response" not in resp:
            raise ValueError("No 'response' key found in the npz file.")
        data = resp["response"]
        if "drift_length" in resp:
            new_params = new_params.replace(response_full_drift_t=float(resp["drift_length"]) / params.vdrift_static)
        resp = data
    elif not isinstance(resp, np.ndarray):
        raise ValueError("Unsupported response format. Expected npz or numpy array.")
    if cache_dir is None:
        return build_response_template(resp, params, device), new_params
    import os
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, "bank_%s.npy" % _bank_cache_key(resp, params))
    want = (len(params.long_diff_template),) + tuple(resp.shape)
    if os.path.exists(path):
        try:
            cached = np.load(path, mmap_mode="r")
            if tuple(cached.shape) == want and cached.dtype == np.float32:
                return torch.from_numpy(np.array(cached)).to(device), new_params
        except (ValueError, OSError):
            pass
    bank = build_response_template(resp, params, device)
    tmp = path + ".tmp.%d" % os.getpid()
    np.save(tmp, bank.cpu().numpy())
    os.replace(tmp + ".npy" if not tmp.endswith(".npy") else tmp, path)      # atomic: concurrent ranks see a complete file or none
    return bank, new_params
