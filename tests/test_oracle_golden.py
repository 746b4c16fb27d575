"""Pins the CPU oracle against the reference's own golden outputs (output/jax_ref/output_{0..4}.h5, committed as
tests/golden/golden_lut_*.npz by tests/golden/make_fixtures.py).

The goldens were produced with the real response_44.npy, which is missing from the reference checkout, so they cannot
pin the waveform arithmetic; they do pin — bit for bit — the geometry, the pixel-id packing, the pixel coordinates, the
hit z, the ADC<->charge maps, the event batching, and (as a set inclusion) the main-pixel computation of the drift stage
(SURVEY.md §8c).  CPU only."""
import os

import numpy as np
import pytest

import common as cm
from oracle import consts as oc
from oracle import larnd_oracle as lo


def _golden(i):
    g = np.load(os.path.join(cm.GOLD, "golden_lut_%d.npz" % i))
    out = {}
    for k in g.files:
        b, e, ds = k.split("/")
        out.setdefault(int(b[1:]), {}).setdefault(int(e[1:]), {})[ds] = g[k]
    return out


@pytest.mark.parametrize("ifile", range(5))
def test_geometry_and_adc_maps_are_bit_exact(ifile):
    p = cm.oracle_params()
    v = oc.get_vdrift(p)
    thr_adc = lo.digitize(p, np.float32(p.DISCRIMINATION_THRESHOLD))
    n = 0
    for batch in _golden(ifile).values():
        for ev in batch.values():
            pix = ev["pixels"]
            assert pix.dtype == np.int32                       # x64 disabled in the reference: ids are int32
            xp, yp, plane, event = lo.id2pixel(p, pix)
            xy = lo.get_pixel_coordinates(p, xp, yp, plane)
            assert np.array_equal(xy[:, 0], ev["pix_x"]) and np.array_equal(xy[:, 1], ev["pix_y"])
            assert np.array_equal(lo.get_hit_z(p, ev["ticks"], plane, v), ev["pix_z"])
            assert np.array_equal(lo.adc2charge(ev["adc"], p), ev["Q"])
            assert np.array_equal(ev["adc"] - thr_adc, ev["adc_clean"])
            assert np.array_equal(lo.pixel2id(p, xp, yp, plane, event), pix)   # pack(unpack(id)) == id
            assert (ev["ticks"] == np.round(ev["ticks"])).all() and (ev["ticks"] < 1997).all()
            n += len(pix)
    assert n == (777, 786, 479, 538, 748)[ifile]               # hit counts quoted in SURVEY.md Appendix A


@pytest.mark.parametrize("ifile", range(5))
def test_batching_and_main_pixels_against_goldens(ifile):
    """Replay of TracksDataset batching reproduces the golden batch->event partition; every fired golden pixel is a
    main pixel of the oracle's drift stage for that batch (ids carry the batch-local event id)."""
    gold = _golden(ifile)
    batches = cm.fixture_batches(ifile, 0.005)
    assert len(batches) == len(gold)
    p = cm.oracle_params()
    for ib, (arr, gids) in enumerate(batches):
        assert sorted(gold[ib].keys()) == sorted(set(gold[ib].keys()) & set(int(g) for g in gids))
        d = lo.simulate_drift_new(p, arr, cm.FIELDS)
        main = np.unique(d["main_pixels"])
        fired = np.unique(np.concatenate([e["pixels"] for e in gold[ib].values()]))
        assert np.isin(fired, main).all()     # 1 767 fired golden pixels over the 18 batches, every one a main pixel
        local = {int(g): i for i, g in enumerate(gids)}
        for gid, ev in gold[ib].items():
            _, _, _, event = lo.id2pixel(p, ev["pixels"])
            assert (event == local[gid]).all()


@pytest.mark.parametrize("ifile,ib", [(0, 1), (3, 0)])
def test_per_pixel_charge_against_goldens_with_synthetic_response(ifile, ib):
    """Coarse known-answer test of the WHOLE chain: the goldens were produced with the real
    response_44.npy, which is absent, so the oracle runs with the synthetic response (unit-integral collecting pulse peaking
    ~70 ticks before the end of the LUT axis, like the real one).  Pulse shapes differ, charge bookkeeping must not: the fired
    pixel set, the collected charge per pixel and the time of the first hit have to agree with the golden within the
    stated bands (multi-hit splitting depends on the true pulse shape, so hit COUNTS are not compared)."""
    p = cm.oracle_params()
    bank = cm.synthetic_bank(32, 45, 45, 1950)
    gold = _golden(ifile)[ib]
    arr, _ = cm.fixture_batches(ifile, 0.005)[ib]
    wfs, uniq = lo.simulate_wfs(p, bank, arr, cm.FIELDS, history={})
    hits = lo.simulate_stochastic(p, wfs, uniq)
    q, ticks, pix = lo.adc2charge(hits[0], p), hits[4], hits[7]
    gp = np.concatenate([e["pixels"] for e in gold.values()])
    gq = np.concatenate([e["Q"] for e in gold.values()])
    gt = np.concatenate([e["ticks"] for e in gold.values()])
    fired_o, fired_g = np.unique(pix), np.unique(gp)
    common = np.intersect1d(fired_o, fired_g)
    assert len(common) >= 0.95 * len(fired_g) and len(fired_o) <= 1.05 * len(fired_g)
    ratio = np.array([q[pix == c].sum() / gq[gp == c].sum() for c in common])
    assert 0.9 < np.median(ratio) < 1.15                                  # survey: 0.99 / 1.10
    lo16, hi84 = np.percentile(ratio, [16, 84])
    assert lo16 > 0.7 and hi84 < 1.5                                      # survey: 0.87 .. 1.29
    assert abs(q.sum() / gq.sum() - 1) < 0.1                              # total collected charge
    dt = np.array([ticks[pix == c].min() - gt[gp == c].min() for c in common])
    assert abs(np.median(dt)) <= 6                                        # first-hit time: a few ticks, no constant offset


def test_input0_batch_sizes_match_survey():
    sizes = [a.shape[0] for a, _ in cm.fixture_batches(0, 0.005)]
    assert sizes == [7373, 10879, 9980, 11776]                 # SURVEY.md §8 (replay of the simulate_test.sh batching)
    assert [len(g) for _, g in cm.fixture_batches(0, 0.005)] == [10, 3, 7, 6]


def test_vdrift_and_constants():
    p = cm.oracle_params()
    assert abs(oc.get_vdrift(p) - 0.159645) < 1e-6             # consts_jax.py:248 vdrift_static
    assert abs(float(oc.get_vdrift(p, traced=True)) - oc.get_vdrift(p)) < 1e-7
    assert lo.hold_interval(p) == 18
    assert int(p.time_interval[1] / p.t_sampling) + 1 == 2001
    assert (p.n_pixels_x, p.n_pixels_y) == (140, 280) and abs(p.pixel_pitch - 0.4434) < 1e-12
    tv = np.asarray(p.long_diff_template)
    assert tv.shape == (100,) and tv[0] == np.float32(0.001) and tv[-1] == np.float32(10)


def test_float_divmod_semantics():
    a = np.array([0.3, -0.3, 5.0, -5.0, 0.0443, 31.038, 1e-8], dtype=np.float32)
    w = np.float32(0.04434)
    q, r = lo.jnp_floor_divide_f(a, w), lo.jnp_remainder_f(a, w)
    # jnp's float floor-divide is round((a - fmod(a, b)) / b) with a sign fix-up: for these values (none within an ulp of a
    # multiple of w) it equals the exact floor of the real quotient
    assert np.array_equal(q, np.floor(a.astype(np.float64) / np.float64(w)).astype(np.float32))
    assert (r >= 0).all() and (r < w).all()
    assert np.allclose(q * w + r, a, atol=1e-5)
    assert lo.jnp_floor_divide_f(np.float32([-0.01]), w)[0] == -1.0


def test_oracle_end_to_end_is_self_consistent():
    """Charge bookkeeping of the oracle: with no neighbours the summed waveform charge of the valid rows equals the
    collected charge of the segments (unit-integral collecting response), and hits only come from main pixels."""
    p = cm.oracle_params(number_pix_neighbors=0, signal_length=150)
    bank = cm.synthetic_bank(32, 5, 5, 1950)
    tr = cm.small_batch(400, ibatch=1, pad=5, precision=0.01)
    wfs, uniq, d, full = lo.simulate_wfs(p, bank, tr, cm.FIELDS, history={}, return_aux=True)
    q_tot = d["nelectrons_neigh"].sum()
    collected = (wfs[uniq >= 0].astype(np.float64).sum() + 0) * p.t_sampling
    assert abs(collected - q_tot) < 0.05 * q_tot
    hits = lo.simulate_stochastic(p, wfs, uniq)
    assert np.isin(hits[7], uniq[uniq >= 0]).all() and len(hits[0]) > 0


def test_jax_random_restatement_reproduces_published_values():
    """oracle/jax_random.py against (i) the Random123 known-answer vectors JAX's own suite checks
    (jax tests/random_test.py::testThreefry2x32) and (ii) outputs printed in the JAX documentation / README for both
    counter layouts: random.normal(random.key(42)) = -0.028304616 (threefry_partitionable, JAX >= 0.5) and -0.18471177
    (original), random.normal(PRNGKey(0), (3,)) = [1.8160863, -0.48262316, 0.33988908] and
    random.split(PRNGKey(0)) = [[4146024105, 967050713], [2718843009, 1272950319]] (original)."""
    from oracle import jax_random as jr
    kat = [((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
           ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
           ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
    for k, x, exp in kat:
        o0, o1 = jr.threefry2x32(k[0], k[1], np.array([x[0]], np.uint32), np.array([x[1]], np.uint32))
        assert (int(o0[0]), int(o1[0])) == exp
    assert abs(float(jr.normal(jr.key(42), (), True)) - (-0.028304616)) < 1e-7
    assert abs(float(jr.normal(jr.key(42), (), False)) - (-0.18471177)) < 1e-7
    assert abs(float(jr.normal(jr.key(0), (), False)) - (-0.20584226)) < 1e-7
    assert np.allclose(jr.normal(jr.key(0), (3,), False), [1.8160863, -0.48262316, 0.33988908], rtol=0, atol=2e-7)
    assert [tuple(int(v) for v in k) for k in jr.split(jr.key(0), 2, False)] == [(4146024105, 967050713), (2718843009, 1272950319)]
