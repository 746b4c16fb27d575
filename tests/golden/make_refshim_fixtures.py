"""Runs the UNMODIFIED reference source (/root/reference/src/larndsim/*.py, optimize/dataio.py) on the numpy stand-in for
jax (tests/golden/jaxshim/, see its README) and commits what it produced as fixtures.

    python tests/golden/make_refshim_fixtures.py            # both precisions (spawns itself with JAXSHIM_X64=0 / 1)

Outputs (tests/golden/):
  refshim_lut_{f32,f64}.npz   sim_jax.simulate_drift_new / simulate_wfs / simulate_stochastic (sim_jax.py:375,689,738) on
                              four fixture batches x configurations (neighbours 0..4, L = 100 / 150, Birks / Box /
                              Ellipsoid, shifts, one case with front-end noise); legacy sim_jax.simulate_signals_new (:456),
                              detsim_jax.accumulate_signals (:157) and current_lut (:642) on the same drift arrays
  refshim_mc_{f32,f64}.npz    sim_jax.simulate_parametrized (:339) for both diffusion variants + the electrons / pixel ids
                              of simulate_drift (:120)
  refshim_prob_{f32,f64}.npz  fee_jax.get_adc_values_average_noise_vmap (:387) + sim_jax.simulate_probabilistic (:772)
  refshim_grad_f64.npz        central finite differences (double) of sum(G * simulate_wfs) w.r.t. every fitted Params leaf in
                              the three recombination models, of the same for simulate_parametrized, and of
                              losses_jax.params_loss (mse_adc, :385) w.r.t. the six leaves of optimize/fit_test.sh
  refshim_misc.npz            consts_jax.load_lut bank rows, get_vdrift, digitize, adc2charge, mse_adc / mmd values,
                              pad_size sequence, dataio.chop_tracks / pad_batch outputs
tests/test_refshim_golden.py holds the oracle (CPU) and the CUDA kernels (GPU) to these files.
/root/reference does not exist on the GPU box: only the .npz files travel.
"""
import os
import subprocess
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("LARND_REFERENCE", "/root/reference")

# name -> (input file, batch, segments, padding rows, chop precision, bank (templates, nx, ny), Params overrides)
LUT_CASES = {
    "n2_L100_birks": (0, 1, 600, 8, 0.01, (32, 25, 25), dict(number_pix_neighbors=2, signal_length=100)),
    "n4_L100_birks": (1, 0, 500, 6, 0.01, (32, 45, 45), dict(number_pix_neighbors=4, signal_length=100)),
    "n1_L150_box_shift": (2, 2, 400, 4, 0.01, (32, 15, 15), dict(number_pix_neighbors=1, signal_length=150, recombination_mode=1,
                                                                 shift_x=0.013, shift_y=-0.021, shift_z=0.017)),
    "n0_L100_ellipsoid": (3, 1, 400, 0, 0.01, (32, 5, 5), dict(number_pix_neighbors=0, signal_length=100, recombination_mode=3)),
    "n2_L100_noise": (0, 2, 500, 8, 0.01, (32, 25, 25), dict(number_pix_neighbors=2, signal_length=100, RESET_NOISE_CHARGE=900,
                                                              UNCORRELATED_NOISE_CHARGE=500)),
}
BASE = dict(electron_sampling_resolution=0.005, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0, time_window=100)
GRAD_STEPS = dict(eField=1e-7, lifetime=1e-2, long_diff=1e-11, tran_diff=1e-11, shift_x=1e-6, shift_y=1e-6, shift_z=1e-6,
                  MeVToElectrons=1e-1, lArDensity=1e-6, Ab=1e-6, kb=1e-7, alpha=1e-6, beta=1e-6, R_param=1e-5)
FIT_NAMES = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")


def weight_field(npix, nticks, seed):
    """The smooth positive cotangent the gradient tests contract the waveforms with."""
    rng = np.random.default_rng(seed)
    tt = np.arange(nticks)
    return rng.uniform(0.5, 1.5, (npix, 1)) * (1 + 0.5 * np.sin(tt[None, :] / 41.0 + rng.uniform(0, 6, (npix, 1))))


def worker():
    sys.path[:0] = [os.path.join(HERE, "jaxshim"), os.path.join(REF, "src"), REF, os.path.join(ROOT, "tests"), ROOT]
    warnings.filterwarnings("ignore")
    import jax
    import jax.numpy as jnp
    from larndsim import consts_jax, detsim_jax, fee_jax, losses_jax, sim_jax
    from optimize import dataio
    import common as cm
    from oracle import consts as oc
    x64 = os.environ.get("JAXSHIM_X64", "0") == "1"
    tag = "f64" if x64 else "f32"
    fdt = np.float64 if x64 else np.float32
    A = lambda a: np.asarray(a).view(np.ndarray)
    tmp = tempfile.mkdtemp(prefix="refshim_")

    def params(names=(), **over):
        cls = consts_jax.build_params_class(list(names))
        p = consts_jax.load_detector_properties(cls, os.path.join(REF, "src/larndsim/detector_properties/module0.yaml"),
                                                os.path.join(REF, "src/larndsim/pixel_layouts/multi_tile_layout-2.4.16_v4.yaml"))
        over = dict(BASE, **over)
        if "recombination_mode" in over:
            over["recombination_mode"] = consts_jax.RecombinationMode(over["recombination_mode"])
        return p.replace(**over)

    banks = {}

    def bank(shape, p):
        """consts_jax.load_lut on the synthetic response of the test-suite (the real response_44.npy is missing from the
        checkout), cut to the first `ntpl` templates like tests/common.py::synthetic_bank."""
        if shape not in banks:
            ntpl, nx, ny = shape
            path = os.path.join(tmp, "resp_%d_%d.npy" % (nx, ny))
            np.save(path, oc.synthetic_response(nx, ny, 1950))
            pp = p.replace(long_diff_template=jnp.linspace(0.001, 10, 100)[:ntpl])
            banks[shape] = consts_jax.load_lut(path, pp)[0]
        return banks[shape]

    def tpl(p, ntpl):
        return p.replace(long_diff_template=jnp.linspace(0.001, 10, 100)[:ntpl])

    # ------------------------------------------------------------------ LUT mode
    out = {}
    for name, (ifile, ibatch, nseg, pad, prec, bshape, over) in LUT_CASES.items():
        sim_jax.size_history_dict.clear()
        p = tpl(params(**over), bshape[0])
        resp = bank(bshape, p)
        tr = cm.small_batch(nseg, ifile=ifile, ibatch=ibatch, pad=pad, precision=prec).astype(fdt)
        d = sim_jax.simulate_drift_new(p, jnp.array(tr), cm.FIELDS)
        for key, arr in zip(("main_pixels", "pixels", "nelectrons", "t0_after_diff", "long_diff", "currents_idx", "pIDs_neigh",
                             "currents_idx_neigh", "nelectrons_neigh", "t0_neigh"), d):
            out["%s/drift/%s" % (name, key)] = A(arr)
        wfs, upix = sim_jax.simulate_wfs(p, resp, jnp.array(tr), cm.FIELDS)
        hits = sim_jax.simulate_stochastic(p, wfs, upix, 0)
        out[name + "/unique_pixels"] = A(upix)
        out[name + "/wfs"] = A(wfs)
        for k, h in enumerate(hits):
            out["%s/hits/%d" % (name, k)] = A(h)
        integral, ticks = fee_jax.get_adc_values(p, wfs, jax.random.key(0))
        out[name + "/integral"], out[name + "/ticks"] = A(integral), A(ticks)
        # legacy entry points on the same drift arrays (sim_jax.py:456-617, detsim_jax.py:157-205): truncating tick, no
        # sub-tick split, plain template sums
        ren_n = jnp.searchsorted(upix, d[6].ravel())
        ren_n = jnp.where((ren_n < upix.size) & (upix[ren_n] == d[6].ravel()), ren_n, 0)
        legacy = sim_jax.simulate_signals_new(p, upix, d[1], d[3], resp, d[2], d[4], d[5], d[8], ren_n, d[9], d[7])
        out[name + "/legacy_wfs"] = A(legacy)
        if name == "n2_L100_birks":
            out[name + "/tracks"] = tr.astype(np.float32)
            z = np.zeros((len(upix), int(p.time_interval[1] / p.t_sampling) + 1), fdt)
            cum = jnp.cumsum(resp, axis=-1)
            ct = (d[9] / p.t_sampling).astype(int)
            npx = (2 * p.number_pix_neighbors + 1) ** 2
            acc = detsim_jax.accumulate_signals(jnp.array(z), d[7], jnp.repeat(d[8], npx), resp[0], cum, ren_n, jnp.repeat(ct, npx),
                                                p.signal_length)
            out[name + "/accumulate_signals"] = A(acc)
            coords = detsim_jax.get_pixel_coordinates(p, *detsim_jax.id2pixel(p, d[0])[:3])
            t0, cidx = detsim_jax.current_lut(p, resp[0], jnp.array(tr), coords, cm.FIELDS)
            out[name + "/current_lut_t0"], out[name + "/current_lut_idx"] = A(t0), A(cidx)
    np.savez_compressed(os.path.join(HERE, "refshim_lut_%s.npz" % tag), **out)
    print(tag, "lut:", len(out), "arrays")

    # ------------------------------------------------------------------ MC-current mode
    out = {}
    for diff in (True, False):
        sim_jax.size_history_dict.clear()
        p = params(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=diff)
        tr = cm.small_batch(500, ibatch=1, pad=8, precision=0.01).astype(fdt)
        k1, _ = jax.random.split(jax.random.key(0))
        rnd = jax.random.normal(k1, (tr.shape[0], 3))
        electrons, pids = sim_jax.simulate_drift(p, jnp.array(tr), cm.FIELDS, k1)
        res = sim_jax.simulate_parametrized(p, jnp.array(tr), cm.FIELDS, 0)
        pre = "diff%d" % int(diff)
        out[pre + "/rnd"], out[pre + "/electrons"], out[pre + "/pIDs"] = A(rnd), A(electrons), A(pids)
        for k, h in enumerate(res):
            out["%s/hits/%d" % (pre, k)] = A(h)
        # the waveforms behind the hits (simulate_signals_parametrized, sim_jax.py:289-335, up to the front end)
        upix = jnp.unique(pids.ravel())
        coords = detsim_jax.get_pixel_coordinates(p, *detsim_jax.id2pixel(p, pids.ravel())[:3])
        t0, sig = detsim_jax.current_mc(p, electrons, coords, cm.FIELDS)
        w = detsim_jax.accumulate_signals_parametrized(jnp.zeros((upix.shape[0], 2001)), sig, jnp.searchsorted(upix, pids.ravel()),
                                                       t0 - sig.shape[1])
        out[pre + "/unique_pixels"], out[pre + "/wfs_full"], out[pre + "/t0_tick"] = A(upix), A(w), A(t0)
    np.savez_compressed(os.path.join(HERE, "refshim_mc_%s.npz" % tag), **out)
    print(tag, "mc:", len(out), "arrays")

    # ------------------------------------------------------------------ probabilistic front end
    out = {}
    sim_jax.size_history_dict.clear()
    p = tpl(params(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0), 32)
    resp = bank((32, 15, 15), p)
    tr = cm.small_batch(600, ibatch=2, pad=2, precision=0.01).astype(fdt)
    wfs, upix = sim_jax.simulate_wfs(p, resp, jnp.array(tr), cm.FIELDS)
    amp = np.abs(A(wfs)).sum(axis=1)
    sel = np.concatenate([np.argsort(-amp)[:5], np.argsort(amp)[:1]])
    w = jnp.array(A(wfs)[sel])
    lp, qd = fee_jax.get_adc_values_average_noise_vmap(p, w)
    res = sim_jax.simulate_probabilistic(p, w, upix[sel])
    out["wfs"], out["unique_pixels"], out["log_prob"], out["charge"] = A(w), A(upix)[sel], A(lp), A(qd)
    for k, h in enumerate(res):
        out["probabilistic/%d" % k] = A(h)
    np.savez_compressed(os.path.join(HERE, "refshim_prob_%s.npz" % tag), **out)
    print(tag, "prob:", len(out), "arrays")

    # ------------------------------------------------------------------ finite-difference gradients (double only)
    if x64:
        out = {}

        def f32_inputs(p, shape):
            """The gradient of the blend w.r.t. long_diff is a heavily cancelling sum of DIFFERENCES of neighbouring
            templates: rounding the bank to float32 moves it by ~1 % (measured).  The product computes with a float32
            bank, so the double-precision differences are taken on exactly that input: the float32 bank and float32
            template grid of the test-suite (tests/common.py::synthetic_bank), passed to the reference as arguments."""
            grid = oc.linspace_jnp(0.001, 10, 100)[:shape[0]]
            return p.replace(long_diff_template=jnp.array(grid.astype(np.float64))), jnp.array(cm.synthetic_bank(*shape).astype(np.float64))

        for mode in (1, 2, 3):
            sim_jax.size_history_dict.clear()
            over = dict(number_pix_neighbors=1, signal_length=100, recombination_mode=mode, shift_x=0.013, shift_y=-0.021, shift_z=0.017)
            p0, resp = f32_inputs(params(**over), (32, 15, 15))
            tr = jnp.array(cm.small_batch(300, ibatch=2, pad=4, precision=0.01).astype(fdt))
            w0, u0 = sim_jax.simulate_wfs(p0, resp, tr, cm.FIELDS)
            G = weight_field(len(u0), w0.shape[1], 11 + mode)

            def L(p):
                sim_jax.size_history_dict.clear()
                w, u = sim_jax.simulate_wfs(p, resp, tr, cm.FIELDS)
                assert np.array_equal(A(u), A(u0))
                return float((A(w) * G).sum())

            names = [n for n in GRAD_STEPS if not ((mode == 2 and n in ("alpha", "beta", "R_param")) or
                                                   (mode != 2 and n in ("Ab", "kb")) or (mode == 1 and n == "R_param"))]
            out["lut_mode%d/names" % mode] = np.array(names)
            out["lut_mode%d/value" % mode] = L(p0)
            out["lut_mode%d/grad" % mode] = np.array([(L(p0.replace(**{n: getattr(p0, n) + GRAD_STEPS[n]})) -
                                                       L(p0.replace(**{n: getattr(p0, n) - GRAD_STEPS[n]}))) / (2 * GRAD_STEPS[n]) for n in names])
            out["lut_mode%d/unique_pixels" % mode] = A(u0)
        # MC-current mode
        for diff in (True, False):
            p0 = params(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=diff)
            tr = jnp.array(cm.small_batch(300, ibatch=1, pad=0, precision=0.01).astype(fdt))
            k1, _ = jax.random.split(jax.random.key(0))

            def W(p):
                electrons, pids = sim_jax.simulate_drift(p, tr, cm.FIELDS, k1)
                upix = jnp.unique(pids.ravel())
                coords = detsim_jax.get_pixel_coordinates(p, *detsim_jax.id2pixel(p, pids.ravel())[:3])
                t0, sig = detsim_jax.current_mc(p, electrons, coords, cm.FIELDS)
                return A(detsim_jax.accumulate_signals_parametrized(jnp.zeros((upix.shape[0], 2001)), sig,
                                                                    jnp.searchsorted(upix, pids.ravel()), t0 - sig.shape[1])), A(upix)
            w0, u0 = W(p0)
            G = weight_field(len(u0), 2000, 5 + int(diff))
            L = lambda p: float((W(p)[0][:, 1:] * G).sum())
            names = ["Ab", "kb", "eField", "lifetime", "long_diff", "shift_x", "shift_z"]
            out["mc_diff%d/names" % diff] = np.array(names)
            out["mc_diff%d/value" % diff] = L(p0)
            out["mc_diff%d/grad" % diff] = np.array([(L(p0.replace(**{n: getattr(p0, n) + GRAD_STEPS[n]})) -
                                                      L(p0.replace(**{n: getattr(p0, n) - GRAD_STEPS[n]}))) / (2 * GRAD_STEPS[n]) for n in names])
            out["mc_diff%d/unique_pixels" % diff] = u0
        # the fit loss of optimize/fit_test.sh (--lut, n = 2, L = 150): params_loss with mse_adc, target = shifted parameters
        sim_jax.size_history_dict.clear()
        p0, resp = f32_inputs(params(FIT_NAMES, number_pix_neighbors=2, signal_length=150), (32, 25, 25))
        tr = jnp.array(cm.small_batch(500, ibatch=1, pad=6, precision=0.01).astype(fdt))
        target = p0.replace(Ab=0.78, kb=0.05, eField=0.52, lifetime=2100.0, tran_diff=9.2e-6, long_diff=4.4e-6)
        wt, ut = sim_jax.simulate_wfs(target, resp, tr, cm.FIELDS)
        ref = sim_jax.simulate_stochastic(target, wt, ut, 0)

        def Lfit(p):
            sim_jax.size_history_dict.clear()
            return float(losses_jax.params_loss(p, resp, ref[0], ref[1], ref[2], ref[3], ref[4], ref[5], ref[6], tr, cm.FIELDS, rngkey=0,
                                                loss_fn=losses_jax.mse_adc)[0])
        out["fit/names"] = np.array(FIT_NAMES)
        out["fit/target"] = np.array([getattr(target, n) for n in FIT_NAMES])
        out["fit/value"] = Lfit(p0)
        out["fit/grad"] = np.array([(Lfit(p0.replace(**{n: getattr(p0, n) + GRAD_STEPS[n]})) - Lfit(p0.replace(**{n: getattr(p0, n) - GRAD_STEPS[n]})))
                                    / (2 * GRAD_STEPS[n]) for n in FIT_NAMES])
        for k, h in enumerate(ref):
            out["fit/ref/%d" % k] = A(h)
        np.savez_compressed(os.path.join(HERE, "refshim_grad_f64.npz"), **out)
        print(tag, "grad:", {k: v for k, v in out.items() if k.endswith("grad")})
        return

    # ------------------------------------------------------------------ small things (float32 run only)
    out = {}
    p = params()
    out["vdrift"] = np.float64(consts_jax.get_vdrift(p))
    resp = oc.synthetic_response(3, 3, 1950)
    path = os.path.join(tmp, "resp_small.npy")
    np.save(path, resp)
    b, _ = consts_jax.load_lut(path, p)
    out["bank_3x3"] = A(b)[::33]                     # templates 0, 33, 66, 99
    x = jnp.array(np.linspace(-2e5, 4e5, 257).astype(np.float32))
    out["digitize_in"], out["digitize_out"] = A(x), A(fee_jax.digitize(p, x))
    a = jnp.array(np.linspace(0, 256, 513).astype(np.float32))
    out["adc2charge_in"], out["adc2charge_out"] = A(a), A(losses_jax.adc2charge(a, p))
    sim_jax.size_history_dict.clear()
    seq = [1000, 1010, 1100, 990, 1500, 1052, 20, 21, 22]
    out["pad_size_in"], out["pad_size_out"] = np.array(seq), np.array([sim_jax.pad_size(s, "t", 0.2) for s in seq])
    rng = np.random.default_rng(3)
    pts_a, pts_b = rng.normal(size=(40, 3)).astype(np.float32), rng.normal(size=(30, 3)).astype(np.float32)
    wa, wb = rng.uniform(0.5, 2, 40).astype(np.float32), rng.uniform(0.5, 2, 30).astype(np.float32)
    out["mmd_in_a"], out["mmd_in_b"], out["mmd_wa"], out["mmd_wb"] = pts_a, pts_b, wa, wb
    out["mmd_out"] = np.float64(losses_jax.mmd(jnp.array(pts_a), jnp.array(pts_b), jnp.array(wa), jnp.array(wb), 0.7))
    seg = np.load(os.path.join(HERE, "segments_input_0.npz"))["segments"]
    from oracle import larnd_oracle as lo
    raw = lo.structured_to_f32(lo.swap_xz_structured(seg))[:40]
    chopped = dataio.chop_tracks(raw.copy(), cm.FIELDS, 0.05)
    out["chop_in"], out["chop_out"] = raw, np.asarray(chopped)
    # TracksDataset bookkeeping (optimize/dataio.py:107-418) on three prepared inputs, production settings + a cut variant
    for ifile, kw in ((0, {}), (1, {}), (7, {}), (7, dict(nevents=5, max_nbatch=2, max_batch_len=30))):
        ds = dataio.TracksDataset(filename=os.path.join(REF, "prepared_data/input_%d.h5" % ifile), **dict(dict(
            nevents=None, max_nbatch=None, swap_xz=True, random_nevents=False, data_seed=42, max_batch_len=50, chopped=True, pad=False,
            electron_sampling_resolution=0.005, live_selection=False), **kw))
        pre = "dataset_%d%s" % (ifile, "_cut" if kw else "")
        out[pre + "/nsteps"] = np.array(ds.batch_nsteps)
        out[pre + "/tot_len"] = np.float64(ds.tot_data_length)
        for b in range(len(ds)):
            out["%s/rows_%d" % (pre, b)] = np.sort(ds.get_batch_row_indices(b))     # the order inside a trajectory comes from an
            out["%s/events_%d" % (pre, b)] = ds.get_batch_global_event_ids(b)       # unstable argsort in the reference: set only
            if b == 0 and not kw:
                arr = ds[b]
                padded = ds.pad_batch(arr, arr.shape[0] + 7, b)
                order = np.lexsort(arr.T[::-1])                                     # row order normalised for the comparison
                out[pre + "/batch0_sorted"], out[pre + "/batch0_pad_tail"] = arr[order], padded[-7:]
    np.savez_compressed(os.path.join(HERE, "refshim_misc.npz"), **out)
    print(tag, "misc:", len(out), "arrays")

    # ------------------------------------------------------------------ the production driver, end to end
    # python -m optimize.simulate with the settings of optimize/simulate_test.sh on prepared_data/input_0.h5 and the synthetic
    # response; what it writes through h5py is captured by the stand-in (tests/golden/jaxshim/h5py.py)
    import argparse
    import h5py
    from optimize import simulate as ref_simulate
    sim_jax.size_history_dict.clear()
    lut_path = os.path.join(tmp, "response_synthetic.npy")
    np.save(lut_path, oc.synthetic_response(45, 45, 1950))
    cfg = argparse.Namespace(
        input_file=os.path.join(REF, "prepared_data/input_0.h5"), output_file=os.path.join(tmp, "out_0.h5"),
        detector_props=os.path.join(REF, "src/larndsim/detector_properties/module0.yaml"),
        pixel_layouts=os.path.join(REF, "src/larndsim/pixel_layouts/multi_tile_layout-2.4.16_v4.yaml"), mode="lut",
        electron_sampling_resolution=0.005, number_pix_neighbors=4, signal_length=100, lut_file=lut_path, noise=False, seed=None,
        diffusion_in_current_sim=False, batch_size=500, gpu=False, jac=False, mc_diff=False, save_wfs=False, n_events=-1, out_np=False,
        max_batch_len=50., chop=True)
    rc, msg = ref_simulate.main(cfg)
    assert rc == 0, msg
    written = h5py.WRITTEN[cfg.output_file]
    np.savez_compressed(os.path.join(HERE, "refshim_simulate_0.npz"), **written)
    print(tag, "simulate:", len(written), "datasets,", sum(len(v) for k, v in written.items() if k.endswith("/adc")), "hits")


if __name__ == "__main__":
    if os.environ.get("REFSHIM_WORKER") == "1":
        worker()
    else:
        for x64 in ("0", "1"):
            env = dict(os.environ, REFSHIM_WORKER="1", JAXSHIM_X64=x64)
            subprocess.check_call([sys.executable, os.path.abspath(__file__)], env=env)
