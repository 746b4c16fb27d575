"""Parity against outputs of the reference's OWN SOURCE.

tests/golden/refshim_*.npz were produced by importing the unmodified reference modules (sim_jax, detsim_jax, fee_jax,
losses_jax, consts_jax, optimize/dataio) on the numpy stand-in for jax in tests/golden/jaxshim/ (generator:
tests/golden/make_refshim_fixtures.py).  They pin the ALGORITHM the reference states; XLA's own float32 code generation is
not reproduced by numpy, so float32 comparisons carry float32 tolerances, while the double-precision runs (both sides in
float64, identical banks) agree to ~1e-9 — any structural difference between the oracle and the reference would show there.

CPU tests: oracle == reference source.  GPU tests: CUDA kernels == reference source, through the C ABI.
"""
import os

import numpy as np
import pytest

import common as cm
from oracle import consts as oc
from oracle import larnd_oracle as lo

G = cm.GOLD
HAVE = os.path.exists(os.path.join(G, "refshim_lut_f32.npz"))
pytestmark = pytest.mark.skipif(not HAVE, reason="refshim fixtures not generated")

# mirror of tests/golden/make_refshim_fixtures.py::LUT_CASES (the generator cannot be imported on the GPU box's behalf: it
# needs /root/reference) — name: (file, batch, segments, pad, precision, bank shape, overrides)
LUT_CASES = {
    "n2_L100_birks": (0, 1, 600, 8, 0.01, (32, 25, 25), dict(number_pix_neighbors=2, signal_length=100)),
    "n4_L100_birks": (1, 0, 500, 6, 0.01, (32, 45, 45), dict(number_pix_neighbors=4, signal_length=100)),
    "n1_L150_box_shift": (2, 2, 400, 4, 0.01, (32, 15, 15), dict(number_pix_neighbors=1, signal_length=150, recombination_mode=1,
                                                                 shift_x=0.013, shift_y=-0.021, shift_z=0.017)),
    "n0_L100_ellipsoid": (3, 1, 400, 0, 0.01, (32, 5, 5), dict(number_pix_neighbors=0, signal_length=100, recombination_mode=3)),
    "n2_L100_noise": (0, 2, 500, 8, 0.01, (32, 25, 25), dict(number_pix_neighbors=2, signal_length=100, RESET_NOISE_CHARGE=900,
                                                              UNCORRELATED_NOISE_CHARGE=500)),
}
GRAD_STEPS = dict(eField=1e-7, lifetime=1e-2, long_diff=1e-11, tran_diff=1e-11, shift_x=1e-6, shift_y=1e-6, shift_z=1e-6,
                  MeVToElectrons=1e-1, lArDensity=1e-6, Ab=1e-6, kb=1e-7, alpha=1e-6, beta=1e-6, R_param=1e-5)
DRIFT_KEYS = ("main_pixels", "pixels", "nelectrons", "t0_after_diff", "long_diff", "currents_idx", "pIDs_neigh",
              "currents_idx_neigh", "nelectrons_neigh", "t0_neigh")


def _load(name):
    return np.load(os.path.join(G, name))


def weight_field(npix, nticks, seed):
    rng = np.random.default_rng(seed)
    tt = np.arange(nticks)
    return rng.uniform(0.5, 1.5, (npix, 1)) * (1 + 0.5 * np.sin(tt[None, :] / 41.0 + rng.uniform(0, 6, (npix, 1))))


def _case(name, product=False):
    ifile, ibatch, nseg, pad, prec, bshape, over = LUT_CASES[name]
    tr = cm.small_batch(nseg, ifile=ifile, ibatch=ibatch, pad=pad, precision=prec)
    p = (cm.product_params if product else cm.oracle_params)(**over)
    if product and "recombination_mode" in over:
        import larndsim_b200 as lb
        p = p.replace(recombination_mode=lb.RecombinationMode(over["recombination_mode"]))
    p = p.replace(long_diff_template=np.asarray(p.long_diff_template)[:bshape[0]])
    return tr, p, bshape


def _bank(bshape, op, dt):
    """The template bank in the working precision: float32 = the cached bank of the whole test-suite; float64 = the same
    construction evaluated in double with the double template grid (what jax_enable_x64 gives the reference)."""
    if dt is np.float32:
        return cm.synthetic_bank(*bshape), op
    op = op.replace(long_diff_template=oc.linspace_jnp(0.001, 10, 100, dtype=np.float64)[:bshape[0]])
    return oc.build_response_template(oc.synthetic_response(bshape[1], bshape[2], 1950), op, n_templates=bshape[0], dtype=np.float64), op


# ------------------------------------------------------------------------------------------------ CPU: oracle == reference
@pytest.mark.parametrize("name", list(LUT_CASES))
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_lut_path_equals_reference_source(name, prec):
    """simulate_drift_new -> unique / renumber -> simulate_signals -> get_adc_values -> parse_output."""
    z = _load("refshim_lut_%s.npz" % prec)
    dt = np.float64 if prec == "f64" else np.float32
    tr, op, bshape = _case(name)
    bank, op = _bank(bshape, op, dt)
    # -- drift stage: integers bit for bit, floats to rounding
    d = lo.simulate_drift_new(op, tr.astype(dt), cm.FIELDS, dt=dt)
    ftol = 1e-9 if prec == "f64" else 3e-6
    for k in DRIFT_KEYS:
        ref, got = z["%s/drift/%s" % (name, k)], np.asarray(d[k])
        assert ref.shape == got.shape, k
        if ref.dtype.kind in "iu":
            assert np.array_equal(ref, got), k
        else:
            assert np.abs(got - ref).max() <= ftol * max(np.abs(ref).max(), 1e-30), (k, np.abs(got - ref).max())
    # -- waveforms (garbage column dropped by simulate_wfs, rows of the -1 padding compared too)
    upix = z[name + "/unique_pixels"]
    w, u = lo.simulate_wfs(op, bank, tr.astype(dt), cm.FIELDS, dt=dt, pad_to=len(upix))
    assert np.array_equal(u, upix)
    ref = z[name + "/wfs"]
    scale = np.abs(ref).max()
    # float32: elementwise float32 arithmetic on both sides, exactly rounded scatter sums on both sides
    assert np.abs(w - ref).max() <= (1e-9 if prec == "f64" else 1e-5) * scale
    if over_noise(name):
        return            # the noisy front end is compared in test_oracle_noisy_front_end_equals_reference_source
    # -- front end on the REFERENCE's waveforms: ticks bit for bit, charge / ADC to rounding
    integral, ticks = lo.get_adc_values(op, ref.astype(dt), dt=dt)
    assert np.array_equal(ticks.astype(np.int64), z[name + "/ticks"].astype(np.int64))
    assert np.abs(integral - z[name + "/integral"]).max() <= (1e-9 if prec == "f64" else 2e-6) * np.abs(z[name + "/integral"]).max()
    hits = lo.simulate_stochastic(op, ref.astype(dt), upix, dt=dt)
    for k in range(8):
        r, g = z["%s/hits/%d" % (name, k)], np.asarray(hits[k])
        assert r.shape == g.shape, k
        if r.dtype.kind in "iu" or k == 4:
            assert np.array_equal(r.astype(np.int64), g.astype(np.int64)), k
        else:
            assert np.abs(g - r).max() <= (1e-9 if prec == "f64" else 2e-6) * max(np.abs(r).max(), 1.0), k


def over_noise(name):
    return LUT_CASES[name][6].get("RESET_NOISE_CHARGE", 0) > 0


def _noise_dict(seed, npix, nmax=10):
    from oracle import jax_random as jr
    buf = jr.fee_noise(seed, npix, nmax)
    base, rest = buf[:npix], buf[npix:].reshape(3, nmax, npix)
    return dict(base=base, extra=rest[0], **{"pass": rest[1], "fail": rest[2]})


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_noisy_front_end_equals_reference_source(prec):
    """get_adc_values with RESET / UNCORRELATED noise (fee_jax.py:170-279): the reference draws through jax.random, the
    oracle gets the same threefry normals from oracle/jax_random.py; the trigger logic on top is what is compared."""
    name = "n2_L100_noise"
    z = _load("refshim_lut_%s.npz" % prec)
    dt = np.float64 if prec == "f64" else np.float32
    _, op, _ = _case(name)
    ref_w, upix = z[name + "/wfs"].astype(dt), z[name + "/unique_pixels"]
    noise = _noise_dict(0, len(upix))
    integral, ticks = lo.get_adc_values(op, ref_w, dt=dt, noise=noise)
    assert np.array_equal(ticks.astype(np.int64), z[name + "/ticks"].astype(np.int64))
    assert (z[name + "/ticks"] < 1998).sum() >= 15          # the case does fire
    assert np.abs(integral - z[name + "/integral"]).max() <= (1e-9 if prec == "f64" else 2e-6) * np.abs(z[name + "/integral"]).max()
    hits = lo.simulate_stochastic(op, ref_w, upix, dt=dt, noise=noise)
    for k in range(8):
        r, g = z["%s/hits/%d" % (name, k)], np.asarray(hits[k])
        assert r.shape == g.shape and np.abs(g.astype(np.float64) - r).max() <= 2e-6 * max(np.abs(r).max(), 1.0), k


@pytest.mark.parametrize("diff", [1, 0])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_mc_current_mode_equals_reference_source(prec, diff):
    """simulate_drift (mc_diff) -> current_mc -> accumulate_signals_parametrized -> front end (sim_jax.py:120-139,289-372)."""
    z = _load("refshim_mc_%s.npz" % prec)
    dt = np.float64 if prec == "f64" else np.float32
    pre = "diff%d" % diff
    op = cm.oracle_params(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=bool(diff))
    tr = cm.small_batch(500, ibatch=1, pad=8, precision=0.01).astype(dt)
    rnd = z[pre + "/rnd"]
    # the draw itself: random.normal(split(key(0))[0], (N, 3))
    from oracle import jax_random as jr
    assert np.allclose(rnd, jr.normal(jr.split(jr.key(0))[0], (tr.shape[0], 3)), rtol=0, atol=1e-6)
    e, pids = lo.simulate_drift_mc(op, tr, cm.FIELDS, rnd, dt)
    assert np.array_equal(pids, z[pre + "/pIDs"])
    assert np.abs(e - z[pre + "/electrons"]).max() <= (1e-9 if prec == "f64" else 3e-6) * np.abs(z[pre + "/electrons"]).max()
    upix = z[pre + "/unique_pixels"]
    out, wfull, uniq = lo.simulate_parametrized(op, tr, cm.FIELDS, rnd, dt=dt, pad_to=len(upix), return_wfs=True)
    assert np.array_equal(uniq, upix)
    ref = z[pre + "/wfs_full"]
    valid = upix >= 0
    scale = np.abs(ref[valid]).max()
    assert np.abs(wfull[valid] - ref[valid]).max() <= (1e-9 if prec == "f64" else 1e-4) * scale   # float32: exp / erf of numpy vs the oracle's own order
    # hits of the reference's own simulate_parametrized (its padding = pad_size(..., 0.05): same hits, -1 rows dropped)
    for k in range(8):
        r, g = z["%s/hits/%d" % (pre, k)], np.asarray(out[k])
        assert r.shape == g.shape, k
        if k in (4, 6, 7):
            assert np.array_equal(r.astype(np.int64), g.astype(np.int64)), k
        else:
            assert np.abs(g - r).max() <= (1e-8 if prec == "f64" else 2e-5) * max(np.abs(r).max(), 1.0), k


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_probabilistic_front_end_equals_reference_source(prec):
    """get_adc_values_average_noise_vmap + simulate_probabilistic (fee_jax.py:334-461, sim_jax.py:772-812)."""
    from oracle import prob_fee as pf
    z = _load("refshim_prob_%s.npz" % prec)
    dt = np.float64 if prec == "f64" else np.float32
    op = cm.oracle_params(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0)
    w = z["wfs"].astype(dt)
    lp, q = pf.get_adc_values_average_noise(op, w, dt=dt)
    assert lp.shape == z["log_prob"].shape and q.shape == z["charge"].shape
    assert np.abs(q - z["charge"]).max() <= (1e-9 if prec == "f64" else 1e-6) * np.abs(z["charge"]).max() + (0 if prec == "f64" else 1e-2)
    if prec == "f64":
        # log-probabilities down to exp(-100): compared as logs wherever they are not the -1000 floor
        live = z["log_prob"] > -90
        assert np.abs(lp - z["log_prob"])[live].max() < 1e-6
        assert np.abs(np.exp(lp) - np.exp(z["log_prob"])).max() < 1e-9
    else:
        assert np.abs(np.exp(lp) - np.exp(z["log_prob"])).max() < 5e-4
    out = pf.simulate_probabilistic(op, w, z["unique_pixels"], dt=dt) if "dt" in pf.simulate_probabilistic.__code__.co_varnames else \
        pf.simulate_probabilistic(op, w, z["unique_pixels"])
    for k in (1, 2, 4):
        assert np.allclose(np.asarray(out[k], dtype=np.float64), z["probabilistic/%d" % k], rtol=2e-6, atol=1e-6), k


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_legacy_entry_points_equal_reference_source(prec):
    """simulate_signals_new (sim_jax.py:456-617), accumulate_signals (detsim_jax.py:157-205), current_lut (:642-660)."""
    z = _load("refshim_lut_%s.npz" % prec)
    dt = np.float64 if prec == "f64" else np.float32
    for name in ("n2_L100_birks", "n1_L150_box_shift", "n0_L100_ellipsoid"):
        tr, op, bshape = _case(name)
        bank, op = _bank(bshape, op, dt)
        d = lo.simulate_drift_new(op, tr.astype(dt), cm.FIELDS, dt=dt)
        upix = z[name + "/unique_pixels"]
        _, ren = lo.unique_and_renumber(d, pad_to=len(upix))
        w = lo.simulate_signals_new(op, upix, d, ren, bank, dt=dt)
        ref = z[name + "/legacy_wfs"]
        assert w.shape == ref.shape and np.abs(w - ref).max() <= (1e-9 if prec == "f64" else 1e-5) * np.abs(ref).max()
        if name == "n2_L100_birks":
            P2 = 25
            ct = (np.repeat(d["t0_neigh"], P2) / dt(op.t_sampling)).astype(np.int32)
            acc = lo.accumulate_signals(np.zeros_like(ref), d["currents_idx_neigh"], np.repeat(d["nelectrons_neigh"], P2), bank[0],
                                        lo.response_cumsum(bank), ren, ct, 100, dt)
            ra = z[name + "/accumulate_signals"]
            assert np.abs(acc - ra).max() <= (1e-9 if prec == "f64" else 1e-5) * np.abs(ra).max()
            px, py, plane, _ = lo.id2pixel(op, d["main_pixels"])
            t0, cidx = lo.current_lut(op, bank[0], tr.astype(dt), lo.get_pixel_coordinates(op, px, py, plane, dt), cm.FIELDS, dt)
            ok = d["main_pixels"] >= 0             # padding rows (pixel id -1) sit exactly on a bin edge: 1 ulp decides
            assert np.array_equal(cidx[ok], z[name + "/current_lut_idx"][ok]) and ok.sum() >= 590
            assert np.abs(t0 - z[name + "/current_lut_t0"]).max() <= 1e-4


def _fd(L, p, names):
    return np.array([(L(p.replace(**{n: getattr(p, n) + GRAD_STEPS[n]})) - L(p.replace(**{n: getattr(p, n) - GRAD_STEPS[n]}))) /
                     (2 * GRAD_STEPS[n]) for n in names])


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_oracle_finite_differences_equal_reference_finite_differences(mode):
    """d/d(leaf) sum(G * simulate_wfs) for every fitted leaf of the recombination model: double-precision central
    differences of the reference source == the same differences of the oracle (they are what the CUDA backward is held to)."""
    z = _load("refshim_grad_f64.npz")
    over = dict(number_pix_neighbors=1, signal_length=100, recombination_mode=mode, shift_x=0.013, shift_y=-0.021, shift_z=0.017)
    op = cm.oracle_params(**over)
    # identical inputs on both sides: the float32 bank / template grid of the test-suite (see f32_inputs in the generator)
    op = op.replace(long_diff_template=np.asarray(op.long_diff_template)[:32])
    bank = cm.synthetic_bank(32, 15, 15).astype(np.float64)
    tr = cm.small_batch(300, ibatch=2, pad=4, precision=0.01).astype(np.float64)
    u0 = z["lut_mode%d/unique_pixels" % mode]
    Gw = weight_field(len(u0), 2000, 11 + mode)

    def L(p):
        w, u = lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, pad_to=len(u0))
        assert np.array_equal(u, u0)
        return float((w * Gw).sum())

    names = [str(n) for n in z["lut_mode%d/names" % mode]]
    assert abs(L(op) - float(z["lut_mode%d/value" % mode])) <= 1e-9 * abs(float(z["lut_mode%d/value" % mode]))
    got, ref = _fd(L, op, names), z["lut_mode%d/grad" % mode]
    assert (np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-7 * np.abs(ref).max()).all(), dict(zip(names, zip(got, ref)))
    assert (ref != 0).all() and len(names) == {1: 11, 2: 11, 3: 12}[mode]


def test_oracle_fit_loss_equals_reference_params_loss():
    """losses_jax.params_loss with mse_adc (the loss of optimize/fit_test.sh) and its six finite-difference gradients."""
    z = _load("refshim_grad_f64.npz")
    names = [str(n) for n in z["fit/names"]]
    op = cm.oracle_params(number_pix_neighbors=2, signal_length=150)
    op = op.replace(long_diff_template=np.asarray(op.long_diff_template)[:32])
    bank = cm.synthetic_bank(32, 25, 25).astype(np.float64)
    tr = cm.small_batch(500, ibatch=1, pad=6, precision=0.01).astype(np.float64)
    ref = [z["fit/ref/%d" % k] for k in range(8)]
    # the target hits themselves (simulate_stochastic at the shifted parameters)
    target = op.replace(**dict(zip(names, z["fit/target"])))
    wt, ut = lo.simulate_wfs(target, bank, tr, cm.FIELDS, dt=np.float64, history={})
    rt = lo.simulate_stochastic(target, wt, ut, dt=np.float64)
    for k in range(8):
        assert rt[k].shape == ref[k].shape and np.allclose(np.asarray(rt[k], np.float64), ref[k], rtol=1e-8, atol=1e-8), k

    def L(p):
        w, u = lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, history={})
        o = lo.simulate_stochastic(p, w, u, dt=np.float64)
        Q, rQ = lo.adc2charge(o[0], p, np.float64), lo.adc2charge(ref[0], p, np.float64)
        return float(lo.mse_adc(p, Q, o[1], o[2], o[3], o[5], o[6], rQ, ref[1], ref[2], ref[3], ref[5], ref[6]))

    assert abs(L(op) - float(z["fit/value"])) <= 1e-8 * abs(float(z["fit/value"]))
    got, want = _fd(L, op, names), z["fit/grad"]
    assert (np.abs(got - want) <= 2e-4 * np.abs(want) + 1e-9).all(), dict(zip(names, zip(got, want)))


def test_small_functions_equal_reference_source():
    """consts_jax.load_lut / get_vdrift, fee_jax.digitize, losses_jax.adc2charge / mmd, sim_jax.pad_size, dataio.chop_tracks."""
    z = _load("refshim_misc.npz")
    op = cm.oracle_params()
    assert abs(oc.get_vdrift(op) - float(z["vdrift"])) < 1e-12
    bank = oc.build_response_template(oc.synthetic_response(3, 3, 1950), cm.oracle_params(RESET_NOISE_CHARGE=0))
    assert np.abs(bank[::33] - z["bank_3x3"]).max() <= 3e-7 * np.abs(z["bank_3x3"]).max()
    assert np.abs(lo.digitize(op, z["digitize_in"]) - z["digitize_out"]).max() <= 3e-5          # 1 ulp at 256
    assert np.abs(lo.adc2charge(z["adc2charge_in"], op) - z["adc2charge_out"]).max() <= 2e-5
    hist = {}
    assert [lo.pad_size(int(s), "t", 0.2, hist) for s in z["pad_size_in"]] == list(z["pad_size_out"])
    assert np.array_equal(lo.chop_tracks(z["chop_in"], cm.FIELDS, 0.05), z["chop_out"])
    a, b, wa, wb = (z[k].astype(np.float64) for k in ("mmd_in_a", "mmd_in_b", "mmd_wa", "mmd_wb"))
    k = lambda u, v: np.exp(-((u[:, None] - v[None]) ** 2).sum(-1) / (2 * 0.7 ** 2))
    mmd = (k(a, a) * wa[:, None] * wa).sum() / wa.sum() ** 2 + (k(b, b) * wb[:, None] * wb).sum() / wb.sum() ** 2 - \
        2 * (k(a, b) * wa[:, None] * wb).sum() / (wa.sum() * wb.sum())
    assert abs(mmd - float(z["mmd_out"])) <= 1e-5 * abs(mmd)


# ------------------------------------------------------------------------------------------ GPU: CUDA kernels == reference
WFS_RTOL = 1e-5      # of the row maximum, against the reference's float32 run (float32 accumulation order differs here)
WFS_RTOL64 = 2e-4    # against the reference evaluated in DOUBLE: float32 itself (sub-tick fraction of t0 / t_sampling ~ 1000
                     # ticks, template blend) sits ~5e-5 away from double on sharp pulses
ADC_ATOL = 2e-3      # ADC counts (reference's own acceptance bar: 1e-2, optimize/comparison.py:183)
GRAD_RTOL = 2e-3


@pytest.fixture(scope="module")
def torch_dev(cuda_lib):
    import torch
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("impl", ["chunk", "sorted"])
@pytest.mark.parametrize("name", [n for n in LUT_CASES if not LUT_CASES[n][6].get("RESET_NOISE_CHARGE")])
def test_cuda_lut_path_equals_reference_source(torch_dev, name, impl, monkeypatch):
    """simulate_wfs + simulate_stochastic through the C ABI against the reference source's outputs: unique pixels, hit
    pixels / ticks / events bit for bit; waveforms against the reference's DOUBLE evaluation; ADC within 2e-3 counts."""
    import torch
    from larndsim_b200 import sim
    monkeypatch.setenv("LARND_ACC_IMPL", impl)
    z32, z64 = _load("refshim_lut_f32.npz"), _load("refshim_lut_f64.npz")
    tr, pp, bshape = _case(name, product=True)
    bank = torch.as_tensor(cm.synthetic_bank(*bshape), device=torch_dev)
    upix = z32[name + "/unique_pixels"]
    st = sim.lut_forward(pp, bank, torch.as_tensor(tr, device=torch_dev), cm.FIELDS, npix_capacity=len(upix))
    assert np.array_equal(st.unique_pixels.cpu().numpy(), upix)
    assert np.array_equal(sim.record_fields(st)["MAINPIX"].cpu().numpy(), z32[name + "/drift/main_pixels"])
    w = st.wfs_full[:, 1:].cpu().numpy()
    real = upix >= 0
    for ref, tol in ((z32[name + "/wfs"], WFS_RTOL), (z64[name + "/wfs"], WFS_RTOL64)):
        scale = np.abs(ref[real]).max(axis=1, keepdims=True)
        err = np.abs(w[real] - ref[real])
        assert (err <= tol * scale + 1e-3).all(), (tol, (err / (scale + 1e-30)).max())
    hits = [t.detach().cpu().numpy() for t in sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0)]
    for k in range(8):
        r = z32["%s/hits/%d" % (name, k)]
        assert r.shape == hits[k].shape, (k, r.shape, hits[k].shape)
        if k == 0:
            assert np.abs(hits[0] - r).max() <= ADC_ATOL
        elif k in (4, 5, 6, 7):
            assert np.array_equal(hits[k].astype(np.int64), r.astype(np.int64)), k
        else:
            assert np.abs(hits[k] - r).max() <= 1e-5, k
    # the front end alone on the reference's own float32 waveforms: integrated charge and ticks
    integral, ticks = _fee_on(pp, z32[name + "/wfs"], upix, torch_dev)
    assert np.array_equal(ticks.astype(np.int64), z32[name + "/ticks"].astype(np.int64))
    assert np.abs(integral - z32[name + "/integral"]).max() <= 2e-6 * np.abs(z32[name + "/integral"]).max()


def _fee_on(pp, wfs, upix, dev, noise=None):
    import torch
    from larndsim_b200 import fee
    out = fee.get_adc_values(pp, torch.as_tensor(np.ascontiguousarray(wfs), device=dev), noise)
    return out[0].detach().cpu().numpy(), out[1].cpu().numpy()


@pytest.mark.gpu
def test_cuda_noisy_front_end_equals_reference_source(torch_dev):
    """Front end with reset / uncorrelated noise drawn by the device threefry stream (csrc/rng.cu) for jax.random.key(0)."""
    import torch
    from larndsim_b200 import sim
    name = "n2_L100_noise"
    z = _load("refshim_lut_f32.npz")
    _, pp, _ = _case(name, product=True)
    upix = z[name + "/unique_pixels"]
    w = torch.as_tensor(z[name + "/wfs"], device=torch_dev)
    hits = [t.detach().cpu().numpy() for t in sim.simulate_stochastic(pp, w, torch.as_tensor(upix, device=torch_dev), 0)]
    for k in range(8):
        r = z["%s/hits/%d" % (name, k)]
        assert r.shape == hits[k].shape, k
        if k == 0:
            assert np.abs(hits[0] - r).max() <= 5e-3         # normals agree to ~1e-6 x 900 e- of noise
        elif k in (4, 5, 6, 7):
            assert np.array_equal(hits[k].astype(np.int64), r.astype(np.int64)), k
        else:
            assert np.abs(hits[k] - r).max() <= 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("impl", ["chunk", "sorted"])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_cuda_gradients_equal_reference_finite_differences(torch_dev, mode, impl, monkeypatch):
    """The backward kernels against double-precision central differences of the REFERENCE source for every fitted leaf."""
    import torch
    import larndsim_b200 as lb
    from larndsim_b200 import _lib, sim
    monkeypatch.setenv("LARND_ACC_IMPL", impl)
    z = _load("refshim_grad_f64.npz")
    over = dict(number_pix_neighbors=1, signal_length=100, shift_x=0.013, shift_y=-0.021, shift_z=0.017)
    pp = cm.product_params(**over).replace(recombination_mode=lb.RecombinationMode(mode))
    pp = pp.replace(long_diff_template=np.asarray(pp.long_diff_template)[:32])
    bank = torch.as_tensor(cm.synthetic_bank(32, 15, 15), device=torch_dev)
    tr = cm.small_batch(300, ibatch=2, pad=4, precision=0.01)
    u0 = z["lut_mode%d/unique_pixels" % mode]
    st = sim.lut_forward(pp, bank, torch.as_tensor(tr, device=torch_dev), cm.FIELDS, npix_capacity=len(u0))
    assert np.array_equal(st.unique_pixels.cpu().numpy(), u0)
    Gw = weight_field(len(u0), 2000, 11 + mode).astype(np.float32)
    value = float((st.wfs_full[:, 1:].double() * torch.as_tensor(Gw, device=torch_dev).double()).sum())
    assert abs(value - float(z["lut_mode%d/value" % mode])) <= 2e-5 * abs(float(z["lut_mode%d/value" % mode]))
    grad = sim.lut_backward(st, torch.as_tensor(Gw, device=torch_dev)).cpu().numpy()
    for n, fd in zip(z["lut_mode%d/names" % mode], z["lut_mode%d/grad" % mode]):
        g = grad[_lib.PARAM_ORDER.index(str(n))]
        assert abs(g - fd) <= GRAD_RTOL * abs(fd) + 1e-6 * np.abs(grad).max(), (mode, str(n), g, fd)


@pytest.mark.gpu
def test_cuda_fit_loss_and_gradients_equal_reference_params_loss(torch_dev):
    """losses.params_loss (mse_adc) and its autograd gradients against the reference's params_loss and the double-precision
    finite differences of it — the quantity every optimize/ fit of the reference minimises."""
    import torch
    from larndsim_b200 import losses
    z = _load("refshim_grad_f64.npz")
    names = [str(n) for n in z["fit/names"]]
    P = cm.product_params(grad=names, number_pix_neighbors=2, signal_length=150)
    P = P.replace(long_diff_template=np.asarray(P.long_diff_template)[:32])
    bank = torch.as_tensor(cm.synthetic_bank(32, 25, 25), device=torch_dev)
    tr = torch.as_tensor(cm.small_batch(500, ibatch=1, pad=6, precision=0.01), device=torch_dev)
    ref = [torch.as_tensor(z["fit/ref/%d" % k].astype(np.float32 if k != 6 else np.int32), device=torch_dev) for k in range(7)]
    loss, _ = losses.params_loss(P, bank, ref[0], ref[1], ref[2], ref[3], ref[4], ref[5], ref[6], tr, cm.FIELDS, rngkey=0)
    assert abs(float(loss) - float(z["fit/value"])) <= 2e-3 * abs(float(z["fit/value"]))
    loss.backward()
    for n, fd in zip(names, z["fit/grad"]):
        g = float(getattr(P, n).grad)
        assert abs(g - fd) <= 5e-3 * abs(fd) + 1e-9, (n, g, fd)


@pytest.mark.gpu
@pytest.mark.parametrize("diff", [1, 0])
def test_cuda_mc_current_mode_equals_reference_source(torch_dev, diff):
    import torch
    from larndsim_b200 import _lib, sim
    z32, z64, zg = _load("refshim_mc_f32.npz"), _load("refshim_mc_f64.npz"), _load("refshim_grad_f64.npz")
    pre = "diff%d" % diff
    pp = cm.product_params(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=bool(diff))
    tr = cm.small_batch(500, ibatch=1, pad=8, precision=0.01)
    trd, rnd = torch.as_tensor(tr, device=torch_dev), torch.as_tensor(z32[pre + "/rnd"], device=torch_dev)
    upix = z32[pre + "/unique_pixels"]
    st = sim.mc_forward(pp, trd, cm.FIELDS, rnd, npix_capacity=len(upix))
    assert np.array_equal(st.unique_pixels.cpu().numpy(), upix)
    w = st.wfs_full.cpu().numpy()
    for ref, tol in ((z32[pre + "/wfs_full"], 1e-4), (z64[pre + "/wfs_full"], 2e-4)):   # float32 exp / erf / erfc on both sides
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(w - ref)[:, 1:] <= tol * scale + 1e-2).all(), tol
    # the public entry point with the seed (device threefry draw) == the reference's simulate_parametrized(seed 0)
    out = [t.cpu().numpy() for t in sim.simulate_parametrized(pp, trd, cm.FIELDS, rngseed=0)]
    for k in range(8):
        r = z32["%s/hits/%d" % (pre, k)]
        assert r.shape == out[k].shape, k
        if k == 0:
            assert np.abs(out[0] - r).max() <= ADC_ATOL
        elif k in (4, 5, 6, 7):
            assert np.array_equal(out[k].astype(np.int64), r.astype(np.int64)), k
        else:
            assert np.abs(out[k] - r).max() <= 1e-5, k
    # gradients against the reference's finite differences (300 un-padded segments, seed-0 draw)
    tr2 = cm.small_batch(300, ibatch=1, pad=0, precision=0.01)
    from oracle import jax_random as jr
    rnd2 = jr.normal(jr.split(jr.key(0))[0], (tr2.shape[0], 3))
    u0 = zg["mc_diff%d/unique_pixels" % diff]
    t2 = torch.as_tensor(tr2, device=torch_dev)
    st2 = sim.mc_forward(pp, t2, cm.FIELDS, torch.as_tensor(rnd2, device=torch_dev), npix_capacity=len(u0))
    assert np.array_equal(st2.unique_pixels.cpu().numpy(), u0)
    Gw = weight_field(len(u0), 2000, 5 + diff).astype(np.float32)
    grad = sim.mc_backward(st2, t2, torch.as_tensor(Gw, device=torch_dev)).cpu().numpy()
    for n, fd in zip(zg["mc_diff%d/names" % diff], zg["mc_diff%d/grad" % diff]):
        g = grad[_lib.PARAM_ORDER.index(str(n))]
        assert abs(g - fd) <= 5e-3 * abs(fd) + 1e-6 * np.abs(grad).max(), (diff, str(n), g, fd)


@pytest.mark.gpu
def test_cuda_probabilistic_front_end_equals_reference_source(torch_dev):
    import torch
    from larndsim_b200 import fee, sim
    z32, z64 = _load("refshim_prob_f32.npz"), _load("refshim_prob_f64.npz")
    pp = cm.product_params(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0)
    w = torch.as_tensor(z32["wfs"], device=torch_dev)
    lp, qd = fee.get_adc_values_average_noise_vmap(pp, w)
    lp, qd = lp.cpu().numpy(), qd.cpu().numpy()
    assert lp.shape == z64["log_prob"].shape
    assert np.abs(qd - z32["charge"]).max() <= 1e-6 * np.abs(z32["charge"]).max() + 1e-2       # plain float32 arithmetic
    assert np.abs(qd - z64["charge"]).max() <= 5e-6 * np.abs(z64["charge"]).max() + 1e-2
    assert np.abs(np.exp(lp) - np.exp(z64["log_prob"])).max() < 5e-4          # against the reference in double
    out = sim.simulate_probabilistic(pp, w, torch.as_tensor(z32["unique_pixels"], device=torch_dev))
    assert np.abs(out[0].cpu().numpy() - z32["probabilistic/0"]).max() < 2e-3
    for k in (1, 2, 4):
        assert np.allclose(out[k].cpu().numpy().astype(np.float64), z32["probabilistic/%d" % k], rtol=2e-6, atol=1e-6), k


@pytest.mark.gpu
def test_cuda_legacy_entry_points_equal_reference_source(torch_dev):
    """sim.simulate_signals_new, detsim.accumulate_signals, detsim.current_lut (kept with the reference's argument lists)."""
    import torch
    from larndsim_b200 import detsim, sim
    z32 = _load("refshim_lut_f32.npz")
    T = lambda a, **k: torch.as_tensor(np.ascontiguousarray(a), device=torch_dev, **k)
    for name in ("n2_L100_birks", "n1_L150_box_shift", "n0_L100_ellipsoid"):
        tr, pp, bshape = _case(name, product=True)
        bank = T(cm.synthetic_bank(*bshape))
        d = {k: z32["%s/drift/%s" % (name, k)] for k in DRIFT_KEYS}
        upix = z32[name + "/unique_pixels"]
        ren = np.searchsorted(upix, d["pIDs_neigh"].ravel())
        ren = np.where((ren < len(upix)) & (upix[np.minimum(ren, len(upix) - 1)] == d["pIDs_neigh"].ravel()), ren, 0)
        w = sim.simulate_signals_new(pp, T(upix), T(d["pixels"]), T(d["t0_after_diff"]), bank, T(d["nelectrons"]), T(d["long_diff"]),
                                     T(d["currents_idx"]), T(d["nelectrons_neigh"]), T(ren.astype(np.int32)), T(d["t0_neigh"]),
                                     T(d["currents_idx_neigh"])).cpu().numpy()
        ref = z32[name + "/legacy_wfs"]
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
        real = upix >= 0
        assert w.shape == ref.shape and (np.abs(w - ref)[real][:, 1:] <= 1e-5 * scale[real] + 1e-3).all()
        assert (np.abs(w - ref)[:, 0] <= 1e-3 * np.maximum(np.abs(ref[:, 0]), scale[:, 0]) + 1e-3).all()     # garbage column
        if name == "n2_L100_birks":
            ts = np.float32(pp.t_sampling)
            ct = (np.repeat(d["t0_neigh"], 25) / ts).astype(np.int32)
            acc = detsim.accumulate_signals(torch.zeros(ref.shape, device=torch_dev), T(d["currents_idx_neigh"]),
                                            T(np.repeat(d["nelectrons_neigh"], 25)), bank[0], None, T(ren.astype(np.int32)), T(ct), 100)
            ra = z32[name + "/accumulate_signals"]
            sc = np.maximum(np.abs(ra).max(axis=1, keepdims=True), 1.0)
            assert (np.abs(acc.cpu().numpy() - ra)[:, 1:] <= 1e-5 * sc + 1e-3).all()
            px, py, plane, _ = detsim.id2pixel(pp, T(d["main_pixels"]))
            t0, cidx = detsim.current_lut(pp, bank[0], T(z32[name + "/tracks"]), detsim.get_pixel_coordinates(pp, px, py, plane), cm.FIELDS)
            ok = d["main_pixels"] >= 0             # padding rows (pixel id -1) sit exactly on a bin edge: 1 ulp decides
            assert np.array_equal(cidx.cpu().numpy()[ok], z32[name + "/current_lut_idx"][ok])
            assert np.abs(t0.cpu().numpy() - z32[name + "/current_lut_t0"]).max() <= 1e-4
