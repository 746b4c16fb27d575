#!/usr/bin/env python
"""Pins the oracle and the CUDA path against the LIVE reference — wherever ``import jax`` works.

    python scripts/pin_against_reference.py [--reference /root/reference] [--files 0,1,2,3,4] [--platform cpu]

Runs the UNMODIFIED reference (pgranger23/larnd-sim-jax: ``simulate_wfs`` sim_jax.py:689, ``simulate_stochastic`` :738,
``simulate_parametrized`` :339, ``jax.grad(params_loss)`` losses_jax.py:385) on the committed fixture batches
(tests/golden/segments_input_*.npz, batched and chopped exactly as the tests do) with the synthetic response the whole test
suite uses (the real response_44.npy is missing from the reference checkout) and writes

    tests/golden/jaxref_lut_<file>.npz      per batch: unique_pixels, waveform rows of the real pixels, the eight hit arrays
    tests/golden/jaxref_grad.npz            loss and d loss / d (Ab, kb, eField, lifetime, tran_diff, long_diff) of mse_adc
    tests/golden/jaxref_mc.npz              simulate_parametrized hits (mc_diff, n = 0) for seed 0
    tests/golden/jaxref_meta.json           jax / jaxlib versions, platform, x64 flag

tests/test_jaxref_golden.py consumes these files when they exist (oracle on the CPU, CUDA kernels on the GPU) and is skipped
with the reason recorded otherwise.  STATUS: jax is absent from the build image and from the GPU box of this project
(profiles/r2_probe_jax.txt) — this script has not been run; the waveform arithmetic of the oracle stays "parity unpinned"
until it is.  It deliberately uses nothing but the reference's public functions, numpy and the oracle's fixture readers.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRAD_NAMES = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")


def load_reference(ref_root, platform):
    os.environ.setdefault("JAX_PLATFORMS", platform)        # optimize/simulate.py:12-13 defaults to the CPU backend as well
    import jax  # noqa: F401  (ImportError here = nothing to pin against)
    for p in (os.path.join(ref_root, "src"), ref_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from larndsim import consts_jax, losses_jax, sim_jax
    return jax, consts_jax, sim_jax, losses_jax


def reference_params(consts_jax, ref_root, names, **overrides):
    Params = consts_jax.build_params_class(list(names))
    p = consts_jax.load_detector_properties(
        Params, os.path.join(ref_root, "src/larndsim/detector_properties/module0.yaml"),
        os.path.join(ref_root, "src/larndsim/pixel_layouts/multi_tile_layout-2.4.16_v4.yaml"))
    return p.replace(**overrides)


def synthetic_lut_file(nx=45, ny=45, nt=1950):
    """The synthetic response of the test suite (oracle/consts.py::synthetic_response) as a .npy file load_lut can read."""
    from oracle import consts as oc
    path = os.path.join(tempfile.mkdtemp(prefix="larnd_pin_"), "response_synthetic.npy")
    np.save(path, oc.synthetic_response(nx, ny, nt))
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--files", default="0,1,2,3,4")
    ap.add_argument("--platform", default="cpu")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    try:
        jax, consts_jax, sim_jax, losses_jax = load_reference(a.reference, a.platform)
    except ImportError as e:
        print("cannot pin: %s (jax / flax are not importable here)" % e)
        return 2
    import jax.numpy as jnp
    import common as cm
    from oracle import larnd_oracle as lo  # noqa: F401  (fixture readers only)

    base = dict(number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.005, RESET_NOISE_CHARGE=0,
                UNCORRELATED_NOISE_CHARGE=0, time_window=100)
    params = reference_params(consts_jax, a.reference, [], **base)
    response, params = consts_jax.load_lut(synthetic_lut_file(), params)
    fields = cm.FIELDS
    for ifile in [int(x) for x in a.files.split(",")]:
        out = {}
        for ib, (arr, gids) in enumerate(cm.fixture_batches(ifile, 0.005)):
            wfs, upix = sim_jax.simulate_wfs(params, response, jnp.asarray(arr), fields)
            hits = sim_jax.simulate_stochastic(params, wfs, upix, 0)
            upix_np = np.asarray(upix)
            out["b%d/unique_pixels" % ib] = upix_np
            out["b%d/wfs_real" % ib] = np.asarray(wfs)[upix_np >= 0].astype(np.float32)
            for name, h in zip(("adc", "x", "y", "z", "ticks", "hit_prob", "event", "pixel"), hits):
                out["b%d/%s" % (ib, name)] = np.asarray(h)
        np.savez_compressed(os.path.join(a.out, "jaxref_lut_%d.npz" % ifile), **out)
        print("file %d: %d batches" % (ifile, len(cm.fixture_batches(ifile, 0.005))))

    # gradients of the fit loss (BASELINE config 4 settings) on one small batch
    base4 = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0,
                 UNCORRELATED_NOISE_CHARGE=0)
    p_fit = reference_params(consts_jax, a.reference, GRAD_NAMES, **base4)
    resp25, p_fit = consts_jax.load_lut(synthetic_lut_file(25, 25, 1950), p_fit)
    tr = cm.small_batch(2000, ibatch=0, pad=0, precision=0.01)
    p_tgt = p_fit.replace(Ab=0.83, kb=0.055, eField=0.52, lifetime=1.8e3, long_diff=5.0e-6, tran_diff=10e-6)
    w, u = sim_jax.simulate_wfs(p_tgt, resp25, jnp.asarray(tr), fields)
    ref = sim_jax.simulate_stochastic(p_tgt, w, u, 0)
    loss_fn = lambda p: losses_jax.params_loss(p, resp25, ref[0], ref[1], ref[2], ref[3], ref[4], ref[5], ref[6], jnp.asarray(tr), fields,
                                               rngkey=0, loss_fn=losses_jax.mse_adc)
    (loss, _aux), grads = jax.value_and_grad(loss_fn, has_aux=True)(p_fit)
    np.savez(os.path.join(a.out, "jaxref_grad.npz"), loss=float(loss), names=np.array(GRAD_NAMES),
             grads=np.array([float(getattr(grads, n)) for n in GRAD_NAMES]), **{"ref_%d" % i: np.asarray(r) for i, r in enumerate(ref)})

    # MC-current mode
    p_mc = reference_params(consts_jax, a.reference, [], number_pix_neighbors=0, signal_length=150, electron_sampling_resolution=0.01,
                            RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0, mc_diff=True, diffusion_in_current_sim=True)
    mc = sim_jax.simulate_parametrized(p_mc, jnp.asarray(cm.small_batch(500, ibatch=1, pad=8, precision=0.01)), fields, rngseed=0)
    np.savez(os.path.join(a.out, "jaxref_mc.npz"), **{n: np.asarray(h) for n, h in zip(("adc", "x", "y", "z", "ticks", "hit_prob", "event", "pixel"), mc)})

    import jaxlib
    with open(os.path.join(a.out, "jaxref_meta.json"), "w") as fh:
        json.dump({"jax": jax.__version__, "jaxlib": jaxlib.__version__, "platform": a.platform,
                   "x64": bool(jax.config.jax_enable_x64), "threefry_partitionable": bool(jax.config.jax_threefry_partitionable)}, fh)
    print("golden vectors written to", a.out)
    return 0


def reference_rate(ref_root, tracks, fields, steps=3, platform="cpu"):
    """segments/s of the UNMODIFIED reference (simulate_wfs + simulate_stochastic, JAX on the host cores) on one batch — the
    `kind: "reference"` arm of bench.py wherever jax is importable; the first (compiling) call is discarded."""
    import time
    jax, consts_jax, sim_jax, _ = load_reference(ref_root, platform)
    import jax.numpy as jnp
    params = reference_params(consts_jax, ref_root, [], number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.01,
                              RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0, time_window=100)
    response, params = consts_jax.load_lut(synthetic_lut_file(), params)
    tr = jnp.asarray(tracks)
    rates = []
    for i in range(steps + 1):
        t0 = time.time()
        wfs, upix = sim_jax.simulate_wfs(params, response, tr, fields)
        out = sim_jax.simulate_stochastic(params, wfs, upix, 0)
        jax.block_until_ready(out)
        if i:
            rates.append(tracks.shape[0] / (time.time() - t0))
    return float(np.mean(rates)), jax.__version__


if __name__ == "__main__":
    sys.exit(main())
