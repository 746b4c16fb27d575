"""ctypes binding of liblarnd_b200.so (the C ABI declared in include/larnd_b200.h).

The shared library is built in-tree by :func:`build_library` (nvcc, sm_100a) and loaded lazily.  There is
no CPU fallback: if the library is missing or cannot be loaded every kernel entry point raises.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(HERE, "liblarnd_b200.so")
SOURCES = ["api.cu", "prepare.cu", "lut_tables.cu", "accumulate.cu", "accumulate_sorted.cu", "accumulate_bwd.cu", "accumulate_bwd_sorted.cu", "fee.cu", "mc_current.cu", "chop.cu", "rng.cu", "prob_fee.cu", "losses.cu", "stream_ops.cu"]

MAX_TPC = 8
MAX_TEMPLATES = 128
NB_TRAN_BINS = 5
MAX_ADC = 10
NPARAMS = 15
PARAM_ORDER = ("Ab", "kb", "eField", "lifetime", "long_diff", "tran_diff", "shift_x", "shift_y", "shift_z",
               "alpha", "beta", "R_param", "lArDensity", "MeVToElectrons", "vdrift")

# flags of larnd_lut_forward / _accumulate / _backward (include/larnd_b200.h)
FLAG_SKIP_GARBAGE, FLAG_IMPL_CHUNK, FLAG_IMPL_SORTED, FLAG_NO_SPLIT, FLAG_REUSE_RUNS, FLAG_WFS_ZERO = 1, 2, 4, 8, 16, 32
FEE_CLEAR_WFS = 1

# record fields inside the workspace (enum in larnd_b200.h)
REC_FIELDS = ("Q", "FRAC", "SL", "A", "B", "C", "WX0", "WX1", "WX2", "WX3", "WX4", "WY0", "WY1", "WY2", "WY3", "WY4",
              "TD", "X0", "Y0", "ST", "REC", "FT", "XI", "COS2", "T0", "IDX", "BX", "BY", "EP", "FLAGS", "MAINPIX")
REC_INT_FIELDS = ("T0", "IDX", "BX", "BY", "EP", "FLAGS", "MAINPIX")


class Columns(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ncols", "eventID", "x", "y", "z", "z_start", "z_end", "dx", "dEdx", "dE", "t0")]


class PadColumns(C.Structure):
    _fields_ = [("eventID", C.c_int32), ("trackID", C.c_int32), ("pixel_plane", C.c_int32)]


class ChopColumns(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ncols", "x", "y", "z", "x_start", "y_start", "z_start", "x_end", "y_end", "z_end",
                                         "dx", "dE")]


class TrackColumns(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("x_start", "x_end", "y_start", "y_end", "n_electrons", "long_diff", "tran_diff",
                                         "pixel_plane", "t", "t_start", "t_end")]


class CurrentColumns(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ncols", "x", "y", "z", "long_diff", "n_electrons", "pixel_plane")]


class ParamsPOD(C.Structure):
    _fields_ = [
        ("recombination_mode", C.c_int32),
        ("Ab", C.c_float), ("kb", C.c_float), ("alpha", C.c_float), ("beta", C.c_float), ("inv_R2", C.c_float),
        ("efield_rho", C.c_float), ("MeVToElectrons", C.c_float),
        ("vdrift", C.c_float), ("lifetime", C.c_float), ("long_diff", C.c_float), ("tran_diff", C.c_float),
        ("size_margin", C.c_float),
        ("shift_x", C.c_float), ("shift_y", C.c_float), ("shift_z", C.c_float),
        ("n_tpc", C.c_int32),
        ("tpc_borders", C.c_float * 2 * 3 * MAX_TPC),
        ("pixel_pitch", C.c_float), ("bin_width", C.c_float), ("half_pitch", C.c_float),
        ("nb_sampling_bins_per_pixel", C.c_int32), ("n_pixels_x", C.c_int32), ("n_pixels_y", C.c_int32),
        ("number_pix_neighbors", C.c_int32),
        ("tran_bin_edges", C.c_float * (NB_TRAN_BINS + 1)),
        ("t_sampling", C.c_float),
        ("n_ticks", C.c_int32), ("signal_length", C.c_int32), ("n_templates", C.c_int32),
        ("long_diff_template", C.c_float * MAX_TEMPLATES),
        ("discrimination_threshold", C.c_float), ("reset_noise_charge", C.c_float),
        ("uncorrelated_noise_charge", C.c_float),
        ("gain", C.c_float), ("v_cm", C.c_float), ("v_ref_minus_cm", C.c_float), ("v_pedestal", C.c_float),
        ("adc_counts", C.c_float), ("hit_prob_threshold", C.c_float),
        ("hold_interval", C.c_int32), ("max_adc_values", C.c_int32),
        ("diffusion_in_current_sim", C.c_int32),
        ("dvdrift_dEfield", C.c_float),
        ("eField", C.c_float), ("lArDensity", C.c_float), ("R_param", C.c_float),
        ("ts_vdrift", C.c_float),
    ]


class LarndError(RuntimeError):
    pass


_lib = None


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
BUILD_DIR = os.path.join(os.path.dirname(HERE), "build")
HEADERS = ["larnd_common.cuh", "sorted_runs.cuh", "bwd_chain.cuh", "segment_physics.cuh", "chop_math.cuh"]


def nvcc_command(out=LIB_PATH, extra=()):
    """The one-shot form of the build (all sources in one nvcc call); build_library compiles per file in parallel."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return ["nvcc"] + NVCC_FLAGS + ["-shared", "-I" + INCLUDE, "-o", out] + list(extra) + srcs


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a (one nvcc per translation unit, in parallel, objects cached under
    larnd-sim-jax_b200/build/) and link liblarnd_b200.so in-tree, next to this file.  Safe to call from several processes
    at once (torchrun ranks): the build runs under an exclusive file lock and the link writes to a temporary name that is
    renamed into place."""
    import fcntl
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, "larnd_b200.h")]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)

    def fresh():
        return os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in srcs + hdrs)

    if not force and fresh():
        return LIB_PATH
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():  # another process built it while we waited
                return LIB_PATH
            os.makedirs(BUILD_DIR, exist_ok=True)
            jobs, objs = [], []
            for src in srcs:
                obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
                objs.append(obj)
                if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
                    continue
                cmd = ["nvcc"] + NVCC_FLAGS + ["-I" + INCLUDE, "-c", src, "-o", obj]
                jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            failed = []
            for cmd, proc in jobs:
                out, _ = proc.communicate()
                if verbose or proc.returncode != 0:
                    print(" ".join(cmd))
                    print(out)
                if proc.returncode != 0:
                    failed.append(out)
            if failed:
                raise LarndError("nvcc failed:\n" + "\n".join(failed))
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                print(" ".join(cmd))
                print(res.stdout + res.stderr)
            if res.returncode != 0:
                raise LarndError("link failed:\n" + res.stderr)
            os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _declare(lib):
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    PP, PC = C.POINTER(ParamsPOD), C.POINTER(Columns)
    lib.larnd_last_error.restype = C.c_char_p
    lib.larnd_abi_version.restype = C.c_int
    lib.larnd_launch_count.restype = C.c_uint64
    lib.larnd_lut_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.larnd_lut_destroy.argtypes = [vp]
    lib.larnd_lut_destroy.restype = None
    lib.larnd_workspace_bytes.argtypes = [i64, i32, i32, i32, i32]
    lib.larnd_workspace_bytes.restype = sz
    lib.larnd_lut_prepare_neighbours.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.larnd_lut_forward.argtypes = [vp, i64, PC, PP, vp, i32, i32, i32, vp, sz, vp, vp, i64, vp, vp]
    lib.larnd_lut_prepare.argtypes = [vp, i64, PC, PP, vp, i32, vp, sz, vp, vp]
    lib.larnd_lut_accumulate.argtypes = [i64, PP, vp, i32, i32, i32, vp, sz, vp, vp, i64, vp, vp]
    lib.larnd_lut_backward.argtypes = [i64, PP, vp, i32, i32, i32, vp, sz, vp, vp, i64, vp, vp]
    lib.larnd_fee_forward.argtypes = [vp, i64, vp, i32, PP, vp] + [vp] * 16 + [vp, sz, vp]
    lib.larnd_fee_forward_ex.argtypes = [vp, i64, vp, i32, PP, vp] + [vp] * 16 + [vp, sz, i32, vp]
    lib.larnd_fee_forward_ex.restype = C.c_int
    lib.larnd_profile_enable.argtypes = [C.c_int]
    lib.larnd_profile_enable.restype = C.c_int
    lib.larnd_profile_read.argtypes = [C.POINTER(C.c_float)]
    lib.larnd_profile_read.restype = C.c_int
    lib.larnd_fee_scratch_bytes.argtypes = [i32]
    lib.larnd_fee_scratch_bytes.restype = sz
    lib.larnd_fee_backward.argtypes = [vp, vp, vp, i32, PP, vp, i64, i32, vp]
    lib.larnd_build_bank.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    lib.larnd_build_bank.restype = C.c_int
    U2 = C.POINTER(C.c_uint32)
    lib.larnd_rng_split.argtypes = [U2, C.c_int, C.c_int, U2]
    lib.larnd_rng_normal.argtypes = [U2, i64, C.c_int, vp, vp]
    lib.larnd_rng_fee_noise.argtypes = [U2, i32, i32, C.c_int, vp, vp]
    for name in ("larnd_rng_split", "larnd_rng_normal", "larnd_rng_fee_noise"):
        getattr(lib, name).restype = C.c_int
    lib.larnd_prob_fee_scratch_bytes.argtypes = [i32, i32, i32, i32]
    lib.larnd_prob_fee_scratch_bytes.restype = sz
    lib.larnd_prob_fee_forward.argtypes = [vp, i64, i32, i32, PP, i32, C.c_float, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.larnd_prob_fee_forward.restype = C.c_int
    lib.larnd_prob_fee_bwd_scratch_bytes.argtypes = [i32, i32, i32]
    lib.larnd_prob_fee_bwd_scratch_bytes.restype = sz
    lib.larnd_prob_fee_backward.argtypes = [vp, i64, i32, i32, PP, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp, sz, vp]
    lib.larnd_prob_fee_backward.restype = C.c_int
    lib.larnd_rbf_field_scratch_bytes.argtypes = [i32, i32]
    lib.larnd_rbf_field_scratch_bytes.restype = sz
    lib.larnd_rbf_field.argtypes = [vp, i32, vp, vp, i32, C.c_float, vp, vp, sz, vp]
    lib.larnd_rbf_field.restype = C.c_int
    lib.larnd_mse_adc_scratch_bytes.argtypes = [i32, i32, i32]
    lib.larnd_mse_adc_scratch_bytes.restype = sz
    lib.larnd_mse_adc_sums.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, PP, vp, vp, i32, C.c_float, vp, vp, sz, vp]
    lib.larnd_mse_adc_sums.restype = C.c_int
    lib.larnd_mse_adc_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, PP, i32, C.c_float, C.c_float, vp, vp, vp, vp, sz, vp]
    lib.larnd_mse_adc_backward.restype = C.c_int
    PCC = C.POINTER(ChopColumns)
    lib.larnd_chop_count.argtypes = [vp, i64, PCC, C.c_double, vp, vp]
    lib.larnd_chop_tracks.argtypes = [vp, i64, PCC, C.c_double, vp, vp, i64, vp]
    lib.larnd_chop_count.restype = C.c_int
    lib.larnd_lut_prepare_raw.argtypes = [vp, i64, PCC, PC, C.c_double, vp, i64, PP, vp, i32, vp, sz, vp, vp]
    lib.larnd_lut_prepare_raw.restype = C.c_int
    lib.larnd_chop_tracks.restype = C.c_int
    lib.larnd_fee_steps_bytes.argtypes = [i32]
    lib.larnd_fee_steps_bytes.restype = sz
    lib.larnd_fee_backward_steps.argtypes = [vp, vp, vp, i32, PP, vp, sz, i32, vp]
    lib.larnd_fee_backward_steps.restype = C.c_int
    lib.larnd_lut_backward_steps.argtypes = [i64, PP, vp, i32, i32, i32, vp, sz, vp, vp, sz, vp, vp]
    lib.larnd_lut_backward_steps.restype = C.c_int
    lib.larnd_deterministic_scratch_bytes.argtypes = [i32, i32]
    lib.larnd_deterministic_scratch_bytes.restype = sz
    lib.larnd_lut_accumulate_deterministic.argtypes = [i64, PP, vp, i32, i32, i32, vp, sz, vp, vp, i64, vp, vp, sz, vp]
    lib.larnd_lut_accumulate_deterministic.restype = C.c_int
    lib.larnd_lut_forward_deterministic.argtypes = [vp, i64, PC, PP, vp, i32, i32, i32, vp, sz, vp, vp, i64, vp, vp, sz, vp]
    lib.larnd_lut_forward_deterministic.restype = C.c_int
    lib.larnd_batch_gather.argtypes = [vp, i32, vp, vp, i64, i32, vp, vp]
    lib.larnd_batch_gather.restype = C.c_int
    lib.larnd_pad_rows.argtypes = [vp, i32, vp, i64, C.POINTER(PadColumns), vp]
    lib.larnd_pad_rows.restype = C.c_int
    lib.larnd_tracks_stage.argtypes = [vp, i64, PC, C.POINTER(TrackColumns), PP, i32, vp, vp]
    lib.larnd_signals_stream_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, PP, vp, vp, vp, vp]
    lib.larnd_signals_stream_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, PP, vp, vp, i64,
                                                  vp, vp, vp, vp, vp, vp, vp]
    lib.larnd_signals_legacy_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, PP, vp, vp, vp, vp]
    lib.larnd_signals_legacy_forward.restype = C.c_int
    lib.larnd_current_lut.argtypes = [vp, i64, i32, i32, i32, i32, vp, C.c_float, C.c_float, i32, i32, vp, vp, vp]
    lib.larnd_current_lut.restype = C.c_int
    PCU = C.POINTER(CurrentColumns)
    lib.larnd_current_mc.argtypes = [vp, i64, PCU, vp, PP, vp, vp, vp]
    lib.larnd_current_mc_backward.argtypes = [vp, i64, PCU, vp, PP, vp, vp, vp, vp]
    lib.larnd_accumulate_parametrized.argtypes = [vp, i32, i32, vp, i32, vp, vp, i64, vp]
    lib.larnd_accumulate_parametrized_backward.argtypes = [vp, i32, i32, vp, i32, vp, vp, i64, vp]
    for name in ("larnd_tracks_stage", "larnd_signals_stream_forward", "larnd_signals_stream_backward", "larnd_current_mc",
                 "larnd_current_mc_backward", "larnd_accumulate_parametrized", "larnd_accumulate_parametrized_backward"):
        getattr(lib, name).restype = C.c_int
    if hasattr(lib, "larnd_mc_forward"):
        lib.larnd_mc_forward.argtypes = [vp, i64, PC, PP, vp, i32, i32, vp, sz, vp, vp, vp, vp]
        lib.larnd_mc_backward.argtypes = [vp, i64, PC, PP, vp, i32, i32, vp, sz, vp, vp, i64, vp, vp]
    for name in ("larnd_lut_create", "larnd_lut_forward", "larnd_lut_prepare", "larnd_lut_accumulate", "larnd_lut_backward",
                 "larnd_fee_forward", "larnd_fee_backward", "larnd_mc_forward", "larnd_mc_backward", "larnd_lut_prepare_neighbours"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_int


def get_lib():
    """Load the CUDA library; raises LarndError (never falls back to a CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LarndError("liblarnd_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(the CUDA extension is mandatory, there is no CPU fallback)")
        lib = C.CDLL(os.environ.get("LARND_B200_LIB", LIB_PATH))  # override: A/B timing of alternative builds (scripts/)
        _declare(lib)
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise LarndError("larnd_b200 error %d: %s" % (rc, get_lib().larnd_last_error().decode()))
