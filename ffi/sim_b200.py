"""jax.ffi + jax.custom_vjp binding of liblarnd_b200.so behind the reference's own entry points.

Drop this file next to the reference's ``src/larndsim/sim_jax.py`` (or put ``ffi/`` on the path) and switch two imports:

    from sim_b200 import simulate_wfs, simulate_stochastic, simulate_parametrized      # was: from larndsim.sim_jax import ...

``params`` is the reference's flax ``Params`` (``consts_jax.build_params_class``), ``tracks`` / ``response_template`` are
``jax.Array``s on a CUDA device, results are ``jax.Array``s with the reference's shapes and dtypes, and
``jax.value_and_grad(params_loss)`` works unchanged (``optimize/fit_params.py:704-735``): the three entry points are
``jax.custom_vjp`` functions whose forward / backward rules are the XLA custom calls of ``ffi/larnd_ffi.cc``.
Like the reference's ``simulate_wfs`` (``jnp.unique``) they are not traceable under an outer ``jax.jit`` — the pixel
capacity is read back once per call, exactly where the reference synchronises.  Forward-mode AD (the Hessian / Taylor tools'
``jacfwd``) is not supported by a VJP boundary.

STATUS: jax is installed neither in the build image nor on the GPU box of this project (profiles/r2_probe_jax.txt), so this
module has never been executed; ``tests/test_ffi_shim.py`` checks what can be checked without jax (the C++ handlers compile
against a mock of the XLA FFI header, every custom-call target named here exists there, the parameter-block layout used
here is the ctypes mirror of include/larnd_b200.h) and runs the real thing wherever ``import jax`` succeeds.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

import jax
import jax.numpy as jnp
from jax import lax

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# the ctypes mirror of include/larnd_b200.h (pure ctypes, no torch): struct layouts, parameter order, library path
_lib = _load("larnd_b200_lib", os.path.join(PKG, "_lib.py"))
PARAM_ORDER = _lib.PARAM_ORDER
NPARAMS = _lib.NPARAMS

_core = None   # liblarnd_b200.so through ctypes (host-side helpers: LUT handle, workspace size)
_ffi = None    # liblarnd_ffi.so (the XLA handlers)

_TARGETS = ("lut_prepare", "lut_accumulate", "lut_backward", "fee_forward", "fee_backward", "mc_forward", "mc_backward")


def _libs():
    global _core, _ffi
    if _core is None:
        _core = _lib.get_lib()
        path = os.path.join(HERE, "liblarnd_ffi.so")
        if not os.path.exists(path):
            raise _lib.LarndError("ffi/liblarnd_ffi.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "in an environment where `import jax` works")
        _ffi = C.CDLL(path)
        for t in _TARGETS:
            jax.ffi.register_ffi_target("larnd_" + t, jax.ffi.pycapsule(getattr(_ffi, "larnd_ffi_" + t)), platform="CUDA")
    return _core


# ------------------------------------------------------------------------------------------ parameter block
_DEPENDENT = ("Ab", "kb", "alpha", "beta", "inv_R2", "efield_rho", "MeVToElectrons", "vdrift", "dvdrift_dEfield", "lifetime",
              "long_diff", "tran_diff", "shift_x", "shift_y", "shift_z", "eField", "lArDensity", "R_param", "ts_vdrift")


def _word(field):
    off = getattr(_lib.ParamsPOD, field).offset
    assert off % 4 == 0
    return off // 4


def _static_pod(params, n_templates):
    """Host image of larnd_params_t for the STATIC fields (same rules as larndsim_b200.sim.make_pod: Python-float constant
    expressions in double, rounded to float32 once).  Fields that depend on fittable parameters are filled by params_blob."""
    P = _lib.ParamsPOD()
    mode = params.recombination_mode
    P.recombination_mode = int(getattr(mode, "value", mode))
    P.size_margin = float(params.size_margin)
    borders = np.asarray(params.tpc_borders, dtype=np.float64)
    P.n_tpc = borders.shape[0]
    for i in range(borders.shape[0]):
        for j in range(3):
            for k in range(2):
                P.tpc_borders[i][j][k] = borders[i, j, k]
    nb = int(params.nb_sampling_bins_per_pixel)
    P.pixel_pitch, P.bin_width, P.half_pitch = params.pixel_pitch, params.pixel_pitch / nb, params.pixel_pitch / 2
    P.nb_sampling_bins_per_pixel, P.n_pixels_x, P.n_pixels_y = nb, int(params.n_pixels_x), int(params.n_pixels_y)
    P.number_pix_neighbors = int(params.number_pix_neighbors)
    sym = (_lib.NB_TRAN_BINS - 1) // 2
    w = np.float32(params.pixel_pitch / nb)
    for i, e in enumerate(np.asarray(jnp.linspace(-sym * w, (sym + 1) * w, _lib.NB_TRAN_BINS + 1), dtype=np.float32)):
        P.tran_bin_edges[i] = e
    P.t_sampling = params.t_sampling
    P.n_ticks = int(params.time_interval[1] / params.t_sampling) + 1
    P.signal_length = int(params.signal_length)
    tpl = np.asarray(params.long_diff_template, dtype=np.float32)
    P.n_templates = tpl.shape[0]
    for i, t in enumerate(tpl):
        P.long_diff_template[i] = t
    P.discrimination_threshold = params.DISCRIMINATION_THRESHOLD
    P.reset_noise_charge, P.uncorrelated_noise_charge = params.RESET_NOISE_CHARGE, params.UNCORRELATED_NOISE_CHARGE
    P.gain, P.v_cm, P.v_pedestal = params.GAIN, params.V_CM, params.V_PEDESTAL
    P.v_ref_minus_cm = params.V_REF - params.V_CM
    P.adc_counts = params.ADC_COUNTS
    P.hit_prob_threshold = params.hit_prob_threshold
    P.hold_interval = round((3 * params.CLOCK_CYCLE + params.ADC_HOLD_DELAY * params.CLOCK_CYCLE) / params.t_sampling)
    P.max_adc_values = int(params.MAX_ADC_VALUES)
    P.diffusion_in_current_sim = int(bool(params.diffusion_in_current_sim))
    return np.frombuffer(bytes(P), dtype=np.uint32).copy()


def params_blob(params, static_words, get_vdrift):
    """uint8 device array holding larnd_params_t: the static words are constants, every field that depends on a fittable
    parameter is a jnp expression of the (possibly traced) Params leaves — evaluated by XLA with the reference's own
    rounding (``get_vdrift`` is the reference's function; its derivative comes from ``jax.grad``)."""
    f32 = lambda x: jnp.asarray(x, dtype=jnp.float32)
    e = f32(params.eField)
    v = f32(get_vdrift(params))
    dv = f32(jax.grad(lambda x: get_vdrift(params.replace(eField=x)))(e))
    vals = dict(Ab=params.Ab, kb=params.kb, alpha=params.alpha, beta=params.beta, inv_R2=1.0 / f32(params.R_param) ** 2,
                efield_rho=e * f32(params.lArDensity), MeVToElectrons=params.MeVToElectrons, vdrift=v, dvdrift_dEfield=dv,
                lifetime=params.lifetime, long_diff=params.long_diff, tran_diff=params.tran_diff, shift_x=params.shift_x,
                shift_y=params.shift_y, shift_z=params.shift_z, eField=e, lArDensity=params.lArDensity, R_param=params.R_param,
                ts_vdrift=f32(params.t_sampling) * v)
    words = jnp.asarray(static_words)
    idx = jnp.asarray([_word(k) for k in _DEPENDENT], dtype=jnp.int32)
    new = jnp.stack([lax.bitcast_convert_type(f32(lax.stop_gradient(vals[k])), jnp.uint32) for k in _DEPENDENT])
    words = words.at[idx].set(new)
    return lax.bitcast_convert_type(words, jnp.uint8).reshape(-1)


def columns_blob(fields):
    c = _lib.Columns()
    f = tuple(fields)
    c.ncols = len(f)
    for name in ("eventID", "x", "y", "z", "z_start", "z_end", "dx", "dEdx", "dE", "t0"):
        setattr(c, name, f.index(name))
    return jnp.asarray(np.frombuffer(bytes(c), dtype=np.uint8))


# ------------------------------------------------------------------------------------------ static per-call context
_lut_cache = {}


class _Ctx:
    """Everything that is not a traced value: LUT handle, shapes, the static half of the parameter block."""

    def __init__(self, params, response_template, tracks, fields):
        from larndsim.consts_jax import get_vdrift          # the reference's own function
        lib = _libs()
        self.get_vdrift = get_vdrift
        self.fields = tuple(fields)
        self.leaf_names = tuple(n for n in PARAM_ORDER if hasattr(params, n) and isinstance(getattr(params, n), jax.Array))
        self.n = int(tracks.shape[0])
        ev = np.asarray(tracks[:, self.fields.index("eventID")])
        self.n_events = max(int(ev.max()) + 1, 0) if ev.size else 0
        self.static_words = _static_pod(params, None)
        self.n_ticks = int(params.time_interval[1] / params.t_sampling) + 1
        self.stride = (self.n_ticks + 3 + 3) // 4 * 4          # padded waveform rows: 16-byte vector reductions (larnd_b200.h)
        self.n_tpc = int(np.asarray(params.tpc_borders).shape[0])
        self.ws_bytes = int(lib.larnd_workspace_bytes(self.n, self.n_events, self.n_tpc, int(params.n_pixels_x), int(params.n_pixels_y)))
        self.lut = 0
        if response_template is not None:
            key = (response_template.unsafe_buffer_pointer(), tuple(response_template.shape), int(params.signal_length),
                   int(params.nb_sampling_bins_per_pixel), int(params.number_pix_neighbors))
            if key not in _lut_cache:
                h = C.c_void_p()
                ntpl, nx, ny, nt = response_template.shape
                jax.block_until_ready(response_template)
                _lib.check(lib.larnd_lut_create(C.c_void_p(key[0]), ntpl, nx, ny, nt, int(params.signal_length), C.c_void_p(0), C.byref(h)))
                _lib.check(lib.larnd_lut_prepare_neighbours(h, int(params.nb_sampling_bins_per_pixel), int(params.number_pix_neighbors), C.c_void_p(0)))
                C.cdll.LoadLibrary("libcudart.so").cudaDeviceSynchronize()
                _lut_cache[key] = (h, response_template)        # keeps the bank alive as long as the handle
            self.lut = int(_lut_cache[key][0].value)


def _pad_size(cur, tag, thr):
    from larndsim.sim_jax import pad_size                       # shares the reference's size history
    return pad_size(cur, tag, thr)


def _shape(shape, dtype):
    return jax.ShapeDtypeStruct(tuple(int(s) for s in shape), dtype)


# ------------------------------------------------------------------------------------------ simulate_wfs
def _wfs_forward(ctx, pod, tracks):
    ws, counts = jax.ffi.ffi_call("larnd_lut_prepare", (_shape((ctx.ws_bytes,), jnp.uint8), _shape((4,), jnp.int32)))(
        tracks, pod, columns_blob(ctx.fields), lut=np.int64(ctx.lut), n_events=np.int32(ctx.n_events))
    cnt = np.asarray(counts)                                     # the reference synchronises here too (jnp.unique)
    if cnt[2] & 2:
        raise ValueError("eventID outside [-1, n_events) found in tracks")
    npix = _pad_size(int(cnt[0]) + 1, "unique_pixels", 0.2)
    wfs, upix, ws, counts = jax.ffi.ffi_call(
        "larnd_lut_accumulate",
        (_shape((npix, ctx.stride), jnp.float32), _shape((npix,), jnp.int32), _shape((ctx.ws_bytes,), jnp.uint8), _shape((4,), jnp.int32)),
        input_output_aliases={0: 2, 1: 3})(ws, counts, pod, lut=np.int64(ctx.lut), n_events=np.int32(ctx.n_events),
                                            n_segments=np.int64(ctx.n), flags=np.int32(0))
    return wfs, upix, ws, counts


def _make_wfs(ctx, params):
    """custom_vjp over theta = the traced Params leaves, closed over everything static."""

    def with_theta(theta):
        return params.replace(**{n: theta[i] for i, n in enumerate(ctx.leaf_names)}) if ctx.leaf_names else params

    @jax.custom_vjp
    def wfs_fn(theta, tracks):
        wfs, upix, _, _ = _wfs_forward(ctx, params_blob(with_theta(theta), ctx.static_words, ctx.get_vdrift), tracks)
        return wfs, upix

    def fwd(theta, tracks):
        pod = params_blob(with_theta(theta), ctx.static_words, ctx.get_vdrift)
        wfs, upix, ws, counts = _wfs_forward(ctx, pod, tracks)
        return (wfs, upix), (pod, ws, counts, tracks.shape)

    def bwd(res, g):
        pod, ws, counts, tshape = res
        g_wfs = g[0]
        grad, _ = jax.ffi.ffi_call("larnd_lut_backward", (_shape((NPARAMS,), jnp.float32), _shape((ctx.ws_bytes,), jnp.uint8)),
                                   input_output_aliases={1: 1})(g_wfs, ws, counts, pod, lut=np.int64(ctx.lut),
                                                                n_events=np.int32(ctx.n_events), n_segments=np.int64(ctx.n),
                                                                flags=np.int32(0))
        idx = jnp.asarray([PARAM_ORDER.index(n) for n in ctx.leaf_names], dtype=jnp.int32)
        return grad[idx], jnp.zeros(tshape, jnp.float32)

    wfs_fn.defvjp(fwd, bwd)
    return wfs_fn


def simulate_wfs(params, response_template, tracks, fields):
    """(wfs (Npix, Nticks-1) f32, unique_pixels (Npix,) int32) — reference: sim_jax.py:689-736."""
    ctx = _Ctx(params, response_template, tracks, fields)
    theta = jnp.stack([jnp.asarray(getattr(params, n), jnp.float32) for n in ctx.leaf_names]) if ctx.leaf_names else jnp.zeros((0,), jnp.float32)
    wfs, upix = _make_wfs(ctx, params)(theta, jnp.asarray(tracks, jnp.float32))
    # the padded buffer never leaves this module: column 0 is the garbage tick, columns >= n_ticks the alignment padding
    return wfs[:, 1:ctx.n_ticks], upix


# ------------------------------------------------------------------------------------------ simulate_stochastic
def _fee_noise(params, npix, rngseed):
    """The standard normals get_adc_values draws from jax.random.key(rngseed) (fee_jax.py:186,237-255,271), in the FEE
    kernel's layout [base | extra(10) | pass(10) | fail(10)] — drawn with jax.random itself here."""
    if params.RESET_NOISE_CHARGE == 0 and params.UNCORRELATED_NOISE_CHARGE == 0:
        return jnp.zeros((0,), jnp.float32)
    key0 = rngseed if not isinstance(rngseed, (int, np.integer)) else jax.random.key(int(rngseed))
    k = int(params.MAX_ADC_VALUES)
    base = jax.random.normal(key0, (npix,))                      # fee_jax.py:186
    key = jax.random.split(key0, 1)[0]                           # :271
    extra, passed, failed = [], [], []
    for _ in range(k):                                           # :237-255, one split before every draw
        key, = jax.random.split(key, 1)
        extra.append(jax.random.normal(key, (npix,)))
        key, = jax.random.split(key, 1)
        passed.append(jax.random.normal(key, (npix,)))
        key, = jax.random.split(key, 1)
        failed.append(jax.random.normal(key, (npix,)))
    return jnp.concatenate([base] + extra + passed + failed).astype(jnp.float32)


def _make_fee(pod, npix, n_ticks, stride, k, scratch_bytes):
    @jax.custom_vjp
    def fee(wfs_full, upix, noise):
        return _fee_call(wfs_full, upix, noise)[:6]

    def _fee_call(wfs_full, upix, noise):
        shapes = (_shape((npix, k), jnp.float32),) * 3 + (_shape((npix,), jnp.float32),) * 2 + \
                 (_shape((npix,), jnp.int32), _shape((npix, 32), jnp.float32), _shape((1,), jnp.int32), _shape((scratch_bytes,), jnp.uint8))
        adc, ticks, pz, px, py, ev, saved, nv, _ = jax.ffi.ffi_call("larnd_fee_forward", shapes)(wfs_full, upix, pod, noise)
        return adc, ticks, pz, px, py, ev, saved

    def fwd(wfs_full, upix, noise):
        out = _fee_call(wfs_full, upix, noise)
        return out[:6], (out[1], out[6], upix.shape, noise.shape)

    def bwd(res, g):
        ticks, saved, ushape, nshape = res
        g_wfs = jax.ffi.ffi_call("larnd_fee_backward", _shape((npix, stride), jnp.float32))(g[0], ticks, saved, pod,
                                                                                            raw_charge=np.int32(0))
        return g_wfs, np.zeros(ushape, jax.dtypes.float0), jnp.zeros(nshape, jnp.float32)

    fee.defvjp(fwd, bwd)
    return fee


def simulate_stochastic(params, wfs, unique_pixels, rngseed):
    """(adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, hit_pixels), each (nb_valid,) — reference: sim_jax.py:738-769.
    get_adc_values + digitize + id2pixel + get_pixel_coordinates run in the fused kernel; get_hit_z and parse_output stay
    the reference's jnp code, so eField reaches pixel_z through ordinary JAX autodiff."""
    from larndsim.consts_jax import get_vdrift
    from larndsim.detsim_jax import get_hit_z, id2pixel
    from larndsim.sim_jax import parse_output
    lib = _libs()
    npix, ntw = int(wfs.shape[0]), int(wfs.shape[1])
    n_ticks = ntw + 1
    stride = (n_ticks + 3 + 3) // 4 * 4
    k = int(params.MAX_ADC_VALUES)
    pod = params_blob(params, _static_pod(params, None), get_vdrift)
    wfs_full = jnp.pad(wfs, ((0, 0), (1, stride - n_ticks)))      # back to the kernels' padded row layout
    fee = _make_fee(pod, npix, n_ticks, stride, k, int(lib.larnd_fee_scratch_bytes(npix)))
    adcs, ticks, _pz, pixel_x, pixel_y, event = fee(wfs_full, unique_pixels.astype(jnp.int32), _fee_noise(params, npix, rngseed))
    hit_prob = jnp.where(ticks < ntw - 3, 1., 0.)
    _, _, pixel_plane, _ = id2pixel(params, unique_pixels)
    pixel_z = get_hit_z(params, ticks.flatten(), jnp.repeat(pixel_plane, k))
    adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, hit_pixels, nb_valid = parse_output(
        params, adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, unique_pixels)
    return (adcs[:nb_valid], pixel_x[:nb_valid], pixel_y[:nb_valid], pixel_z[:nb_valid], ticks[:nb_valid], hit_prob[:nb_valid],
            event[:nb_valid], hit_pixels[:nb_valid])


# ------------------------------------------------------------------------------------------ simulate_parametrized
def simulate_parametrized(params, tracks, fields, rngseed=0):
    """MC-current mode (n = 0, mc_diff) — reference: sim_jax.py:339-372.  The (N, 3) normals of generate_electrons are drawn
    with jax.random from the first half of split(key(rngseed)), the front end uses the second half, like the reference."""
    ctx = _Ctx(params, None, tracks, fields)
    master = jax.random.key(rngseed)
    k1, k2 = jax.random.split(master)
    rnd = jax.random.normal(k1, (ctx.n, 3)).astype(jnp.float32)
    tracks = jnp.asarray(tracks, jnp.float32)

    def with_theta(theta):
        return params.replace(**{n: theta[i] for i, n in enumerate(ctx.leaf_names)}) if ctx.leaf_names else params

    def call(pod, npix):
        return jax.ffi.ffi_call("larnd_mc_forward", (_shape((npix, ctx.n_ticks), jnp.float32), _shape((npix,), jnp.int32),
                                                     _shape((4,), jnp.int32), _shape((ctx.ws_bytes,), jnp.uint8)))(
            tracks, rnd, pod, columns_blob(ctx.fields), n_events=np.int32(ctx.n_events))

    pod0 = params_blob(params, ctx.static_words, ctx.get_vdrift)
    cnt = np.asarray(call(lax.stop_gradient(pod0), 1)[2])        # capacity 1: only counts the distinct pixels (one sync, like jnp.unique)
    npix = _pad_size(max(int(cnt[0]), 1), "unique_pixels", 0.05)

    @jax.custom_vjp
    def mc(theta):
        wfs, upix, _, _ = call(params_blob(with_theta(theta), ctx.static_words, ctx.get_vdrift), npix)
        return wfs, upix

    def fwd(theta):
        pod = params_blob(with_theta(theta), ctx.static_words, ctx.get_vdrift)
        wfs, upix, counts, ws = call(pod, npix)
        return (wfs, upix), (pod, ws, counts)

    def bwd(res, g):
        pod, ws, counts = res
        grad, _ = jax.ffi.ffi_call("larnd_mc_backward", (_shape((NPARAMS,), jnp.float32), _shape((ctx.ws_bytes,), jnp.uint8)),
                                   input_output_aliases={3: 1})(g[0], tracks, rnd, ws, counts, pod, columns_blob(ctx.fields),
                                                                n_events=np.int32(ctx.n_events))
        return (grad[jnp.asarray([PARAM_ORDER.index(n) for n in ctx.leaf_names], dtype=jnp.int32)],)

    mc.defvjp(fwd, bwd)
    theta = jnp.stack([jnp.asarray(getattr(params, n), jnp.float32) for n in ctx.leaf_names]) if ctx.leaf_names else jnp.zeros((0,), jnp.float32)
    wfs, upix = mc(theta)
    return simulate_stochastic(params, wfs[:, 1:], upix, k2)
