"""ORACLE (test infrastructure only) — parameter container, geometry loader, drift
velocity, LUT bank builder and the synthetic LUT generator.

CPU restatement of reference ``src/larndsim/consts_jax.py``:
  * parameter defaults ........... consts_jax.py:244-298
  * geometry from the YAMLs ...... consts_jax.py:300-385
  * get_vdrift ................... consts_jax.py:193-216
  * load_lut (template bank) ..... consts_jax.py:387-449
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; the product path never does.

PARITY STATUS: the reference is pure JAX and JAX is not installable here, so this
restatement is pinned only against what the reference's goldens can pin without the
(missing) response LUT blob: geometry, pixel-id packing, pixel coordinates, hit z and
the ADC<->charge maps (tests/test_oracle_golden.py).  Waveform arithmetic is
"parity unpinned" (see DESIGN.md).
"""
import json
import math
from types import SimpleNamespace

import numpy as np

BOX, BIRKS, ELLIPSOID = 1, 2, 3  # RecombinationMode values, consts_jax.py:20-23


class Params(SimpleNamespace):
    """Plain attribute bag standing in for the flax struct (consts_jax.py:26-160)."""

    def replace(self, **kw):
        d = dict(self.__dict__)
        d.update(kw)
        return Params(**d)


def linspace_jnp(start, stop, num, dtype=np.float32):
    """jnp.linspace semantics: start*(1-s)+stop*s with s=i/div, exact endpoint appended
    (jax/_src/numpy/lax_numpy.py::_linspace; used at consts_jax.py:295, sim_jax.py:401,
    detsim_jax.py:621)."""
    dt = np.dtype(dtype).type
    div = num - 1
    step = (np.arange(div, dtype=dtype) / dt(div)).astype(dtype)
    out = dt(start) * (dt(1) - step) + dt(stop) * step
    return np.concatenate([out, np.array([stop], dtype=dtype)]).astype(dtype)


def default_params_dict():
    """Defaults of load_detector_properties (consts_jax.py:244-298)."""
    return dict(
        eField=0.50, Ab=0.8, kb=0.0486, vdrift=0.1648, vdrift_static=0.159645,
        lifetime=2.2e3, long_diff=4.0e-6, tran_diff=8.8e-6,
        shift_x=0.0, shift_y=0.0, shift_z=0.0,
        recombination_mode=BIRKS, lArDensity=1.38, alpha=0.93, beta=0.212, R_param=1.25,
        MeVToElectrons=4.237e4, temperature=87.17, max_active_pixels=0, max_radius=0,
        min_step_size=0.001, time_max=0, time_window=189.1, e_charge=1.602e-19,
        t_sampling=0.1, time_padding=190, time_interval=[0, 200], drift_length=0,
        response_bin_size=0.04434, number_pix_neighbors=1,
        electron_sampling_resolution=0.001, signal_length=150, MAX_ADC_VALUES=10,
        DISCRIMINATION_THRESHOLD=7e3, ADC_HOLD_DELAY=15, CLOCK_CYCLE=0.1, GAIN=4e-3,
        V_CM=288, V_REF=1300, V_PEDESTAL=580, ADC_COUNTS=2 ** 8,
        RESET_NOISE_CHARGE=900, UNCORRELATED_NOISE_CHARGE=500,
        ELECTRON_MOBILITY_PARAMS=(551.6, 7158.3, 4440.43, 4.29, 43.63, 0.2053),
        size_margin=2e-2, diffusion_in_current_sim=True, mc_diff=False,
        tpc_centers=np.zeros((2, 3)), response_full_drift_t=190.61638,
        nb_tran_diff_bins=5, nb_sampling_bins_per_pixel=10,
        long_diff_template=linspace_jnp(0.001, 10, 100), long_diff_extent=20,
        roi_threshold=0.01, roi_split_length=400, fee_paths_scaling=20,
        hit_prob_threshold=1e-5, tran_diff_bin_edges=None,
    )


def load_detector_properties(detprop_file, pixel_file):
    """Geometry + constants from the two YAML files (consts_jax.py:220-385)."""
    import yaml

    p = default_params_dict()
    with open(detprop_file) as fh:
        det = yaml.safe_load(fh)
    for k, v in det.items():
        if k not in p:
            raise ValueError("Key '%s' in detector properties file is not recognized." % k)
        p[k] = np.array(v) if isinstance(v, list) else v
    centers = np.array(p["tpc_centers"], dtype=np.float64)
    centers[:, [2, 0]] = centers[:, [0, 2]]  # consts_jax.py:315
    with open(pixel_file) as fh:
        lay = yaml.safe_load(fh)
    mm2cm = 0.1
    pitch = lay["pixel_pitch"] * mm2cm
    pos = np.array(list(lay["chip_channel_to_position"].values()))
    xs = pos[:, 0] * pitch
    ys = pos[:, 1] * pitch
    tile_bx = (-(xs.max() + pitch) / 2, (xs.max() + pitch) / 2)
    tile_by = (-(ys.max() + pitch) / 2, (ys.max() + pitch) / 2)
    tile_idx = lay["tile_indeces"]
    tpc_ids = np.unique(np.array(list(tile_idx.values()))[:, 0])
    try:
        drift_length = lay["drift_length"] * mm2cm
    except KeyError:
        tp = np.array(list(lay["tile_positions"].values())) * mm2cm
        drift_length = 0.5 * (tp[:, 0].max() - tp[:, 0].min()) * mm2cm
    borders = np.zeros((len(tpc_ids), 3, 2))
    for it, tid in enumerate(tpc_ids):
        tiles = [t for t in tile_idx if tile_idx[t][0] == tid]
        dirs = {lay["tile_orientations"][t][0] for t in tiles}
        if len(dirs) != 1:
            raise ValueError("Tiles in same anode plane have different drift directions.")
        d = dirs.pop()
        if d not in (1, -1):
            raise ValueError("Cathode direction should be either 1 or -1.")
        tp = np.vstack([lay["tile_positions"][t] for t in tiles]) * mm2cm
        borders[it, 0] = (tp[:, 2].min() + tile_bx[0] + centers[it][0], tp[:, 2].max() + tile_bx[1] + centers[it][0])
        borders[it, 1] = (tp[:, 1].min() + tile_by[0] + centers[it][1], tp[:, 1].max() + tile_by[1] + centers[it][1])
        borders[it, 2] = (tp[:, 0].min() + centers[it][2], tp[:, 0].max() + drift_length * d + centers[it][2])
    p["pixel_pitch"] = pitch
    p["drift_length"] = drift_length
    p["n_pixels_x"] = len(np.unique(xs)) * 2
    p["n_pixels_y"] = len(np.unique(ys)) * 4
    p["tpc_borders"] = borders  # float64 here; the simulation casts to float32 (x64 disabled in the reference)
    p.pop("tpc_centers")
    return Params(**p)


_GEOM_KEYS = ("pixel_pitch", "drift_length", "n_pixels_x", "n_pixels_y", "tpc_borders",
              "DISCRIMINATION_THRESHOLD", "V_CM", "V_REF", "eField", "lifetime", "vdrift",
              "MeVToElectrons", "Ab", "kb", "long_diff", "tran_diff", "time_interval")


def save_geometry_json(params, path):
    d = {}
    for k in _GEOM_KEYS:
        v = getattr(params, k)
        d[k] = np.asarray(v).tolist() if isinstance(v, (np.ndarray, list, tuple)) else v
    with open(path, "w") as fh:
        json.dump(d, fh, indent=1)


def params_from_geometry_json(path):
    """Module-0 parameters from the committed fixture (tests/golden/module0_geometry.json),
    produced by tests/golden/make_fixtures.py from the reference YAMLs."""
    with open(path) as fh:
        d = json.load(fh)
    p = default_params_dict()
    p.pop("tpc_centers")
    for k, v in d.items():
        p[k] = np.array(v, dtype=np.float64) if k == "tpc_borders" else v
    return Params(**p)


def get_vdrift(params, traced=False):
    """Walkowiak/BNL mobility parametrisation (consts_jax.py:193-216).
    ``traced=False``: eField is a static Python float, the whole expression is evaluated in
    Python double precision and enters the kernels as one f32 constant.  ``traced=True``:
    eField is a differentiable leaf -> evaluated op by op in float32."""
    a0, a1, a2, a3, a4, a5 = params.ELECTRON_MOBILITY_PARAMS
    if not traced:
        e = float(params.eField)
        num = a0 + a1 * e + a2 * pow(e, 1.5) + a3 * pow(e, 2.5)
        den = 1 + (a1 / a0) * e + a4 * pow(e, 2) + a5 * pow(e, 3)
        tc = pow(params.temperature / 89, -1.5)
        return num / den * tc / 1000 * e
    f = np.float32
    e = f(params.eField)
    num = f(a0) + f(a1) * e + f(a2) * np.power(e, f(1.5)) + f(a3) * np.power(e, f(2.5))
    den = f(1) + f(a1 / a0) * e + f(a4) * np.power(e, f(2)) + f(a5) * np.power(e, f(3))
    tc = f(pow(params.temperature / 89, -1.5))
    return num / den * tc / f(1000) * e


def gaussian_taps(params, dtype=np.float32):
    """Normalised Gaussian kernels, one per long_diff_template entry (consts_jax.py:427-428)."""
    dt = np.dtype(dtype).type
    ext = int(params.long_diff_extent)
    x = np.arange(-ext, ext + 1, 1).astype(dtype)
    sig = np.asarray(params.long_diff_template, dtype=dtype)[:, None]
    # jax.scipy.stats.norm.pdf = exp(logpdf), logpdf = (log(2 pi scale^2) + (x - loc)^2 / scale^2) / -2
    # (jax/_src/scipy/stats/norm.py), evaluated operation by operation in the working precision
    s2 = (sig * sig).astype(dtype)
    g = np.exp(((np.log((dt(2 * math.pi) * s2).astype(dtype)).astype(dtype) + ((x[None, :] * x[None, :]).astype(dtype) / s2).astype(dtype)) / dt(-2)).astype(dtype))
    g = g.astype(dtype)
    return (g / g.sum(axis=1, keepdims=True)).astype(dtype)


def build_response_template(response, params, n_templates=None, dtype=np.float32):
    """Template bank of load_lut (consts_jax.py:427-447): row t = response (*) Gaussian_t
    with numpy 'same' convolution along time, row 0 overwritten by the raw response.
    ``n_templates`` truncates the bank (tests use < 100 rows to bound memory; valid as long
    as every used template index + 1 stays below it)."""
    from scipy.ndimage import convolve1d

    response = np.asarray(response, dtype=dtype)
    g = gaussian_taps(params, dtype)
    nt = g.shape[0] if n_templates is None else int(n_templates)
    bank = np.empty((nt,) + response.shape, dtype=dtype)
    for t in range(nt):
        # symmetric odd-length kernel + zero padding == np.convolve(x, k, mode='same')
        bank[t] = convolve1d(response, g[t], axis=-1, mode="constant", cval=0.0)
    bank[0] = response
    return bank


def load_lut(lut_file, params, n_templates=None):
    """load_lut (consts_jax.py:387-449) for .npy / .npz response files."""
    resp = np.load(lut_file)
    if isinstance(resp, np.lib.npyio.NpzFile):
        if "response" not in resp:
            raise ValueError("No 'response' key found in the npz file.")
        resp = resp["response"]
    return build_response_template(resp, params, n_templates), params.replace()


def synthetic_response(nx=45, ny=45, nt=1950, seed=7, t_sampling=0.1):
    """Synthetic stand-in for the missing ``response_44.npy``.

    Collecting bins (i,j < 5): unipolar pulse rising with tau ~ 12 ticks to a peak ~70 ticks
    before the end of the axis, normalised so that sum_t R * t_sampling = 1 (one electron
    collected).  Other bins: zero-net bipolar induction pulse whose amplitude decays with
    max(i,j).  Smooth, deterministic given ``seed``."""
    rng = np.random.default_rng(seed)
    t = np.arange(nt, dtype=np.float64)
    peak = nt - 70.0
    resp = np.zeros((nx, ny, nt), dtype=np.float64)
    jitter = rng.uniform(-1.0, 1.0, size=(nx, ny))
    for i in range(nx):
        for j in range(ny):
            r = max(i, j)
            d = math.hypot(i, j)
            pk = peak + 0.6 * jitter[i, j] - 0.15 * d
            if i < 5 and j < 5:
                tau = 12.0 + 0.8 * d
                rise = np.where(t <= pk, np.exp(np.minimum(t - pk, 0.0) / tau), np.exp(-np.maximum(t - pk, 0.0) / 1.5))
                rise /= rise.sum() * t_sampling
                resp[i, j] = rise
            else:
                amp = 0.08 * math.exp(-(r - 4) / 6.0) * (1.0 + 0.1 * jitter[i, j])
                w = 14.0 + 0.5 * r
                gp = np.exp(-0.5 * ((t - (pk - 1.2 * w)) / w) ** 2)
                gm = np.exp(-0.5 * ((t - (pk + 0.2 * w)) / (0.5 * w)) ** 2)
                gp /= gp.sum()
                gm /= gm.sum()
                resp[i, j] = amp * (gp - gm) / t_sampling * 0.1
    return resp.astype(np.float32)
