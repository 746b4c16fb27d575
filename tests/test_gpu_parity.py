"""Parity of the CUDA path (through the C ABI / ctypes host mirror) against the numpy oracle.  GPU only.

Bars (north star): pixel ids, tick indices, template indices and hit sets bit-exact; waveforms within
WFS_RTOL of the row maximum (float32 accumulation order differs from the oracle's float64 scatter-add);
ADC values within ADC_ATOL counts; gradients within GRAD_RTOL of float64 central differences."""
import os

import numpy as np
import pytest

import common as cm
from oracle import consts as oc
from oracle import larnd_oracle as lo

pytestmark = pytest.mark.gpu

WFS_RTOL = 5e-6     # |wfs_cuda - wfs_oracle| <= WFS_RTOL * max|row| (+ tiny absolute floor)
ADC_ATOL = 2e-3     # ADC counts (reference's own acceptance bar is 1e-2, optimize/comparison.py:183)
GRAD_RTOL = 2e-3


@pytest.fixture(scope="module")
def torch_dev(cuda_lib):
    import torch
    return torch.device("cuda", 0)


def _run_lut(torch_dev, tr, bank, npix=None, **kw):
    import torch
    from larndsim_b200 import sim
    pp = cm.product_params(**kw)
    st = sim.lut_forward(pp, torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev), cm.FIELDS,
                         npix_capacity=npix)
    torch.cuda.synchronize()
    return pp, st


GARBAGE_RTOL = 1e-3  # column 0 (the garbage tick the reference drops, sim_jax.py:736) is a cancellation-prone float32 sum:
                     # the oracle's own float32 and float64 evaluations differ by ~2e-4 there


GARBAGE_ROW_RTOL = 1e-4  # rows of pixel ids < 0 (discarded by parse_output) collect contributions from every segment through
                         # many separate float32 atomic flushes; their rounding noise grows with the number of flushes


def _check_wfs(w, ref, rows=None, uniq=None):
    if uniq is not None:
        bad = np.asarray(uniq) < 0
        if bad.any():
            sc = np.abs(ref[bad]).max(axis=1, keepdims=True)
            assert (np.abs(w[bad] - ref[bad]) <= GARBAGE_ROW_RTOL * sc + 1e-3).all()
        rows = ~bad if rows is None else (rows & ~bad)
    if rows is not None:
        w, ref = w[rows], ref[rows]
    if w.shape[1] == 2001:  # full rows: column 0 is the garbage tick
        g, gr = w[:, 0], ref[:, 0]
        assert (np.abs(g - gr) <= GARBAGE_RTOL * np.maximum(np.abs(gr), np.abs(ref).max(axis=1)) + 1e-3).all()
        w, ref = w[:, 1:], ref[:, 1:]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    err = np.abs(w - ref)
    assert (err <= WFS_RTOL * scale + 1e-3).all(), "max rel-to-row-max error %.3g" % (err / (scale + 1e-30)).max()


def _check_records(st, d, op, L, nt=1950):
    from larndsim_b200 import sim
    rec = {k: v.cpu().numpy() for k, v in sim.record_fields(st).items()}
    assert np.array_equal(rec["MAINPIX"], d["main_pixels"])
    assert np.array_equal(rec["BX"], d["bins_pitches"][:, 0]) and np.array_equal(rec["BY"], d["bins_pitches"][:, 1])
    ts = np.float32(op.t_sampling)
    ft = d["t0_neigh"] / ts
    ct = np.clip(np.floor(ft).astype(np.int32), 0, nt - 1)
    assert np.array_equal(rec["T0"], nt - L - ct)
    assert np.array_equal(rec["FRAC"], (ft - ct).astype(np.float32))
    tv = np.asarray(op.long_diff_template, dtype=np.float32)
    assert np.array_equal(rec["IDX"], np.clip(np.searchsorted(tv, d["long_diff_seg"]), 1, tv.shape[0] - 2))
    q = d["nelectrons_neigh"]
    assert np.allclose(rec["Q"], q, rtol=2e-6, atol=1e-6)
    wx = np.stack([rec["WX%d" % i] for i in range(5)], 1)
    wy = np.stack([rec["WY%d" % i] for i in range(5)], 1)
    assert np.abs(wx - d["wx"]).max() < 5e-7 and np.abs(wy - d["wy"]).max() < 5e-7


def _hits_equal(out_o, out_p):
    names = ["adc", "x", "y", "z", "ticks", "hit_prob", "event", "pixel"]
    got = [t.detach().cpu().numpy() for t in out_p]
    assert len(got[0]) == len(out_o[0]), "hit count %d vs oracle %d" % (len(got[0]), len(out_o[0]))
    for n, a, b in zip(names, out_o, got):
        if n == "adc":
            assert np.abs(a - b).max() <= ADC_ATOL if len(a) else True
        else:
            assert np.array_equal(a, b), n


@pytest.fixture(params=["chunk", "sorted"])
def acc_impl(request, monkeypatch):
    """Both accumulate kernels: the chunk kernel (accumulate.cu, small batches) and the class-sorted kernel
    (accumulate_sorted.cu, picked automatically for >= 200k segments; forced here through LARND_ACC_IMPL)."""
    monkeypatch.setenv("LARND_ACC_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("cfg", [
    dict(n=4, L=100, prec=0.005, nseg=2500, pad=60, ibatch=1),    # optimize/simulate_test.sh settings
    dict(n=2, L=150, prec=0.01, nseg=2000, pad=0, ibatch=0),      # optimize/fit_test.sh --lut settings
    dict(n=0, L=100, prec=0.01, nseg=800, pad=10, ibatch=2),      # no neighbours
    dict(n=1, L=30, prec=0.01, nseg=800, pad=10, ibatch=3),       # short window -> 4-slot kernel
    dict(n=2, L=400, prec=0.05, nseg=600, pad=0, ibatch=0),       # optimize/simulate_fwd.sh window -> 16-slot kernel
])
def test_lut_forward_matches_oracle(torch_dev, cfg, acc_impl):
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=cfg["n"], signal_length=cfg["L"])
    nx = max(10 * cfg["n"] + 5, 5)
    bank = cm.synthetic_bank(32, nx, nx, 1950)
    tr = cm.small_batch(cfg["nseg"], ibatch=cfg["ibatch"], pad=cfg["pad"], precision=cfg["prec"])
    op = cm.oracle_params(**kw)
    wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={}, return_aux=True)
    pp, st = _run_lut(torch_dev, tr, bank, npix=len(uniq_o), **kw)
    assert st.counts.cpu().numpy()[2] == 0
    _check_records(st, d, op, cfg["L"])
    assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
    _check_wfs(st.wfs_full.cpu().numpy(), full_o)
    out_o = lo.simulate_stochastic(op, wfs_o, uniq_o)
    out_p = sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0)
    _hits_equal(out_o, out_p)


def test_exact_shape_mode_reproduces_reference_padding(torch_dev):
    """npix_capacity=None follows pad_size(n_unique+1,'unique_pixels',0.2) (sim_jax.py:717-721)."""
    import torch
    from larndsim_b200 import sim
    sim.size_history_dict.clear()
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    kw = dict(number_pix_neighbors=2, signal_length=100)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    hist = {}
    for ib in (0, 1, 2):
        tr = cm.small_batch(700 + 150 * ib, ibatch=ib, pad=5, precision=0.01)
        wfs_o, uniq_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history=hist)
        wfs, upix = sim.simulate_wfs(pp, torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev), cm.FIELDS)
        assert tuple(wfs.shape) == wfs_o.shape
        assert np.array_equal(upix.cpu().numpy(), uniq_o)
        _check_wfs(wfs.cpu().numpy(), wfs_o)


def test_unsorted_segments_and_edges(torch_dev, acc_impl):
    """Shuffled rows (no runs, constant row/window flushes), segments next to the anode (windows sticking out of the
    readout -> garbage tick 0) and outside every TPC."""
    rng = np.random.default_rng(3)
    tr = cm.small_batch(900, ibatch=1, pad=0, precision=0.01)
    c = cm.FIELDS.index
    near = tr[:300].copy()
    shift = 30.45 - np.abs(near[:, c("z")]).max()
    for col in ("z", "z_start", "z_end"):
        near[:, c(col)] += np.sign(near[:, c(col)]) * shift     # last ~1 cm before the anode: T0 < 2
    far = tr[300:360].copy()
    far[:, c("x")] += 100.0                                      # outside every TPC: masked, pixel id -1
    allr = np.concatenate([tr, near, far])
    allr = allr[rng.permutation(len(allr))]
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = cm.synthetic_bank(48, 25, 25, 1950)   # masked segments drift "through" the other TPC: larger sigma_L, higher templates
    op = cm.oracle_params(**kw)
    wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, allr, cm.FIELDS, history={}, return_aux=True)
    pp, st = _run_lut(torch_dev, allr, bank, npix=len(uniq_o), **kw)
    assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
    w = st.wfs_full.cpu().numpy()
    assert np.abs(full_o[:, 0]).max() > 0, "test must exercise the garbage tick"
    _check_wfs(w, full_o, uniq=uniq_o)


def test_skip_garbage_flag_only_changes_garbage_rows(torch_dev, acc_impl):
    from larndsim_b200 import sim
    import torch
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(800, ibatch=0, pad=12, precision=0.01)
    pp = cm.product_params(**kw)
    b, t = torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev)
    full = sim.lut_forward(pp, b, t, cm.FIELDS)
    skip = sim.lut_forward(pp, b, t, cm.FIELDS, npix_capacity=full.npix, flags=1)
    valid = (full.unique_pixels >= 0).cpu().numpy()
    a, s = full.wfs_full.cpu().numpy(), skip.wfs_full.cpu().numpy()
    _check_wfs(s, a, rows=valid)
    assert np.abs(s[~valid]).max() == 0.0


def test_waveform_row_stride_variants_agree(torch_dev, acc_impl):
    """The C ABI takes a waveform row stride: the padded stride (multiple of 4, n_ticks + 3: frames leave as 16-byte vector
    reductions in the tile kernel), the bare n_ticks stride (scalar reductions) and a misaligned base must give the same
    rows; padding columns stay zero."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    pp = cm.product_params(**kw)
    tr = cm.small_batch(3000, pad=16, precision=0.01)
    b, t = torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev)
    ref = sim.lut_forward(pp, b, t, cm.FIELDS)
    npix, nt = ref.npix, ref.pod.n_ticks
    assert ref.wfs_buf.shape[1] % 4 == 0 and ref.wfs_buf.shape[1] >= nt + 3
    assert float(ref.wfs_buf[:, nt:].abs().max()) == 0.0
    a = ref.wfs_full.cpu().numpy()
    scale = np.abs(a).max(axis=1, keepdims=True) + 1e-30
    for buf in (torch.empty((npix, nt), device=torch_dev),                                  # reference layout, stride 2001
                torch.empty(npix * (nt + 7) + 1, device=torch_dev)[1:].view(npix, nt + 7)):   # base not 16-byte aligned
        st = sim.lut_forward(pp, b, t, cm.FIELDS, npix_capacity=npix, out=(torch.empty(npix, dtype=torch.int32, device=torch_dev), buf))
        assert np.array_equal(st.unique_pixels.cpu().numpy(), ref.unique_pixels.cpu().numpy())
        w = st.wfs_full.cpu().numpy()
        assert (np.abs(w - a)[:, 1:] <= 2 * WFS_RTOL * scale + 1e-4).all()


def test_packed_columns_and_raw_entry_match_the_full_batch(torch_dev):
    """`fields` is an argument of the API: a batch reduced to the ten columns the simulation reads (dataio.pack_columns,
    40 instead of 104 bytes per segment on the host-to-device link) must give bit-identical pixel lists, waveforms and hits;
    and dataio.simulate_from_raw (raw rows uploaded, chopped on the device) must reproduce the hits of the host-chopped batch."""
    import torch
    from larndsim_b200 import dataio, sim
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = torch.as_tensor(cm.synthetic_bank(32, 25, 25, 1950), device=torch_dev)
    pp = cm.product_params(**kw).replace(electron_sampling_resolution=0.01)
    seg = lo.swap_xz_structured(np.load(os.path.join(cm.GOLD, "segments_input_2.npz"))["segments"])
    rows, gids = lo.make_batches(seg, 50.0)[0]
    raw = lo.batch_array(seg, rows, gids, cm.FIELDS, False, 0.01)          # un-chopped rows of the first batch, local event ids
    chopped = lo.chop_tracks(raw, cm.FIELDS, 0.01)
    t = torch.as_tensor(chopped, device=torch_dev)
    ref = sim.lut_forward(pp, bank, t, cm.FIELDS)
    packed, pfields = dataio.pack_columns(chopped, cm.FIELDS)
    assert packed.shape == (chopped.shape[0], 10) and pfields == dataio.PACKED_FIELDS
    st = sim.lut_forward(pp, bank, torch.as_tensor(packed, device=torch_dev), pfields, npix_capacity=ref.npix)
    assert torch.equal(st.unique_pixels, ref.unique_pixels)
    scale = ref.wfs_full.abs().amax(dim=1, keepdim=True) + 1e-30
    assert bool(((st.wfs_full - ref.wfs_full).abs() <= 2 * WFS_RTOL * scale + 1e-3).all())   # float atomics: order differs
    hits_ref = sim.simulate_stochastic(pp, ref.wfs_full[:, 1:], ref.unique_pixels, 0)
    hits_raw = dataio.simulate_from_raw(pp, bank, raw, cm.FIELDS, precision=0.01)
    assert len(hits_raw[0]) == len(hits_ref[0]) > 0
    for k in (4, 6, 7):
        assert torch.equal(hits_raw[k], hits_ref[k])
    assert float((hits_raw[0] - hits_ref[0]).abs().max()) <= ADC_ATOL


def test_empty_and_all_padding_batches(torch_dev):
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=1, signal_length=50)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    pp = cm.product_params(**kw)
    b = torch.as_tensor(bank, device=torch_dev)
    empty = torch.zeros((0, 26), device=torch_dev)
    st = sim.lut_forward(pp, b, empty, cm.FIELDS, npix_capacity=4, n_events=0)
    assert st.counts.cpu().numpy()[0] == 0 and (st.unique_pixels.cpu().numpy() == -1).all() and float(st.wfs_full.abs().max()) == 0.0
    pad = lo.pad_batch(np.zeros((0, 26), np.float32), 50, cm.FIELDS)
    op = cm.oracle_params(**kw)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, pad, cm.FIELDS, history={})
    st = sim.lut_forward(pp, b, torch.as_tensor(pad, device=torch_dev), cm.FIELDS, npix_capacity=len(uniq_o))
    assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
    assert float(st.wfs_full.abs().max()) == 0.0
    hits = sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0)
    assert all(len(h) == 0 for h in hits)


def test_capacity_overflow_and_bad_event_ids_are_flagged(torch_dev):
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=1, signal_length=50)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    pp = cm.product_params(**kw)
    tr = cm.small_batch(1500, ibatch=0, pad=0, precision=0.01)   # spans several events
    assert tr[:, 0].max() >= 1
    b, t = torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev)
    st = sim.lut_forward(pp, b, t, cm.FIELDS, npix_capacity=2)
    assert st.counts.cpu().numpy()[2] & 1
    # nothing was written: the caller sees "no pixel" and zero waveforms, and check_state turns the flag into an exception
    assert (st.unique_pixels.cpu().numpy() == -1).all() and float(st.wfs_full.abs().max()) == 0.0
    from larndsim_b200 import LarndError
    with pytest.raises(LarndError):
        sim.check_state(st)
    st = sim.lut_forward(pp, b, t, cm.FIELDS, npix_capacity=64, n_events=1)   # events 1.. are out of the declared range
    assert st.counts.cpu().numpy()[2] & 2
    with pytest.raises(ValueError):
        sim.check_state(st)
    with pytest.raises(ValueError):
        sim.lut_forward(pp, b, t, cm.FIELDS, n_events=1)
    from larndsim_b200 import LarndError
    with pytest.raises(LarndError):
        sim.lut_forward(pp, b, torch.as_tensor(tr), cm.FIELDS)                # CPU tensor: no CPU path
    with pytest.raises(LarndError):
        sim.lut_forward(cm.product_params(number_pix_neighbors=3, signal_length=50), b, t, cm.FIELDS)  # LUT too small


def test_fee_with_injected_noise_matches_oracle(torch_dev):
    import torch
    from larndsim_b200 import sim
    rng = np.random.default_rng(11)
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(1500, ibatch=1, pad=0, precision=0.01)
    op = cm.oracle_params(**kw).replace(RESET_NOISE_CHARGE=900, UNCORRELATED_NOISE_CHARGE=500)
    pp = cm.product_params(**kw).replace(RESET_NOISE_CHARGE=900, UNCORRELATED_NOISE_CHARGE=500)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={})
    npix = len(uniq_o)
    z = rng.normal(size=(31, npix)).astype(np.float32)
    noise = dict(base=z[0], extra=z[1:11], **{"pass": z[11:21], "fail": z[21:31]})
    adc_o, ticks_o = lo.get_adc_values(op, wfs_o, noise=noise)
    fs = sim.fee_forward(pp, torch.as_tensor(wfs_o, device=torch_dev), torch.as_tensor(uniq_o, device=torch_dev),
                         torch.as_tensor(z.reshape(-1), device=torch_dev), compact=False)
    assert np.array_equal(fs.ticks.cpu().numpy(), ticks_o)
    assert np.array_equal(fs.adc.cpu().numpy(), lo.digitize(op, adc_o))


def test_get_adc_values_returns_the_integrated_charge_and_its_gradient(torch_dev):
    """fee.get_adc_values (fee_jax.py:170-279) returns the charge integrals themselves — also for hits the digitiser
    clips (ADC 0 or 256), where un-digitising the ADC would be wrong — and is differentiable w.r.t. the waveforms:
    d adc_k / d wfs[t] = t_sampling on the hit's integration interval, checked against central differences
    of the float64 oracle."""
    import torch
    from larndsim_b200 import fee
    kw = dict(number_pix_neighbors=1, signal_length=100)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    tr = cm.small_batch(1200, ibatch=1, pad=0, precision=0.01)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={})
    wfs_o = wfs_o.copy()
    big = int(np.argmax(lo.get_adc_values(op, wfs_o)[0].max(axis=1)))
    wfs_o[big] *= 300.0   # one fired row far beyond the ADC range: its hits saturate at ADC_COUNTS
    q_o, ticks_o = lo.get_adc_values(op, wfs_o)
    assert (lo.digitize(op, q_o) >= op.ADC_COUNTS).any()
    w = torch.as_tensor(wfs_o, device=torch_dev).requires_grad_(True)
    q, ticks = fee.get_adc_values(pp, w)
    assert np.array_equal(ticks.cpu().numpy(), ticks_o)
    assert np.array_equal(q.detach().cpu().numpy(), q_o)                       # bit for bit, saturated hits included
    assert np.array_equal(fee.digitize(pp, q.detach()).cpu().numpy(), lo.digitize(op, q_o))
    rng = np.random.default_rng(3)
    G = rng.uniform(0.5, 1.5, q_o.shape).astype(np.float32)
    (q * torch.as_tensor(G, device=torch_dev)).sum().backward()
    g = w.grad.cpu().numpy()
    rows = np.where((q_o != 0).any(axis=1))[0][:3]
    for r in rows:                                                              # finite differences on a few fired rows
        for t in (int(ticks_o[r, 0]) - 5, int(ticks_o[r, 0]) + 10, 1500):
            if not 0 <= t < wfs_o.shape[1]:
                continue
            h = 1e-3 * max(abs(float(wfs_o[r, t])), 1.0)
            wp, wm = wfs_o[r:r + 1].astype(np.float64), wfs_o[r:r + 1].astype(np.float64)
            wp, wm = wp.copy(), wm.copy()
            wp[0, t] += h
            wm[0, t] -= h
            fd = ((lo.get_adc_values(op, wp, dt=np.float64)[0] - lo.get_adc_values(op, wm, dt=np.float64)[0]) * G[r:r + 1]).sum() / (2 * h)
            assert abs(g[r, t] - fd) <= 1e-3 * abs(fd) + 1e-6, (r, t, g[r, t], fd)


def test_lut_gradients_match_float64_finite_differences(torch_dev, acc_impl):
    import torch
    from larndsim_b200 import _lib, sim
    kw = dict(number_pix_neighbors=2, signal_length=150)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(300, ibatch=1, pad=0, precision=0.01)
    _, uniq, d, full = lo.simulate_wfs(op, bank, tr, cm.FIELDS, dt=np.float64, history={}, return_aux=True)
    npix = len(uniq)
    rng = np.random.default_rng(5)
    t = np.arange(2000)
    G = (rng.uniform(0.5, 1.5, (npix, 1)) * (1 + 0.5 * np.sin(t[None, :] / 37.0 + rng.uniform(0, 6, (npix, 1))))).astype(np.float32)

    def L(p):
        w, _ = lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, pad_to=npix)
        return float((w * G.astype(np.float64)).sum())

    st = sim.lut_forward(pp, torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev), cm.FIELDS, npix_capacity=npix)
    grad = sim.lut_backward(st, torch.as_tensor(G, device=torch_dev)).cpu().numpy()
    steps = dict(Ab=1e-6, kb=1e-7, eField=1e-7, lifetime=1e-2, long_diff=1e-11, tran_diff=1e-11, shift_x=1e-6, shift_z=1e-6,
                 MeVToElectrons=1e-1)
    for name, h in steps.items():
        base = getattr(op, name)
        fd = (L(op.replace(**{name: base + h})) - L(op.replace(**{name: base - h}))) / (2 * h)
        g = grad[_lib.PARAM_ORDER.index(name)]
        assert abs(g - fd) <= GRAD_RTOL * abs(fd) + 1e-6 * abs(grad).max(), (name, g, fd)
    assert grad[_lib.PARAM_ORDER.index("vdrift")] == 0.0   # params.vdrift is never read by the reference


@pytest.mark.parametrize("mode", [1, 2, 3])   # BOX, BIRKS, ELLIPSOID (consts_jax.py:20-23)
def test_all_fitted_leaves_in_every_recombination_model(torch_dev, mode, acc_impl):
    """Forward parity of the FUSED prepare/accumulate kernels in the Box and Ellipsoid models (quenching_jax.py:18-35) and
    finite-difference checks of every leaf of LARND_P_* the model reads: the Birks pair (Ab, kb) or the Box/Ellipsoid set
    (alpha, beta, R_param), lArDensity, MeVToElectrons, eField, lifetime, the two diffusion coefficients and all three
    shifts — 15/15 leaves over the three modes (vdrift is identically zero: params.vdrift has no reader)."""
    import torch
    import larndsim_b200 as lb
    from larndsim_b200 import _lib, sim
    kw = dict(number_pix_neighbors=1, signal_length=100)
    extra = dict(recombination_mode=mode, shift_x=0.013, shift_y=-0.021, shift_z=0.017)
    op = cm.oracle_params(**kw).replace(**extra)
    pp = cm.product_params(**kw).replace(**dict(extra, recombination_mode=lb.RecombinationMode(mode)))
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    tr = cm.small_batch(300, ibatch=2, pad=4, precision=0.01)
    # forward: records of the fused prepare kernel + waveforms + hit set against the float32 oracle
    wfs_o, uniq_o, d_o, full_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={}, return_aux=True)
    b, t = torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev)
    st = sim.lut_forward(pp, b, t, cm.FIELDS, npix_capacity=len(uniq_o))
    assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
    _check_records(st, d_o, op, kw["signal_length"])
    _check_wfs(st.wfs_full.cpu().numpy(), full_o)
    _hits_equal(lo.simulate_stochastic(op, wfs_o, uniq_o), sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0))
    # gradients against float64 central differences of the oracle
    npix = len(uniq_o)
    rng = np.random.default_rng(11 + mode)
    tt = np.arange(2000)
    G = (rng.uniform(0.5, 1.5, (npix, 1)) * (1 + 0.5 * np.sin(tt[None, :] / 41.0 + rng.uniform(0, 6, (npix, 1))))).astype(np.float32)

    def L(p):
        w, _ = lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, pad_to=npix)
        return float((w * G.astype(np.float64)).sum())

    grad = sim.lut_backward(st, torch.as_tensor(G, device=torch_dev)).cpu().numpy()
    steps = dict(eField=1e-7, lifetime=1e-2, long_diff=1e-11, tran_diff=1e-11, shift_x=1e-6, shift_y=1e-6, shift_z=1e-6,
                 MeVToElectrons=1e-1, lArDensity=1e-6)
    steps.update(dict(Ab=1e-6, kb=1e-7) if mode == 2 else dict(alpha=1e-6, beta=1e-6))
    if mode == 3:
        steps["R_param"] = 1e-5
    for name, h in steps.items():
        base = getattr(op, name)
        fd = (L(op.replace(**{name: base + h})) - L(op.replace(**{name: base - h}))) / (2 * h)
        g = grad[_lib.PARAM_ORDER.index(name)]
        assert abs(g - fd) <= GRAD_RTOL * abs(fd) + 1e-6 * abs(grad).max(), (mode, name, g, fd)
    unused = {1: ("Ab", "kb", "R_param"), 2: ("alpha", "beta", "R_param"), 3: ("Ab", "kb")}[mode] + ("vdrift",)
    for name in unused:   # leaves the model does not read
        assert grad[_lib.PARAM_ORDER.index(name)] == 0.0, name


def test_autograd_end_to_end(torch_dev, acc_impl):
    """build_params_class leaves get gradients through simulate_wfs + simulate_stochastic, like jax.grad in the reference."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=2, signal_length=150)
    names = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
    P = cm.product_params(grad=names, **kw)
    op = cm.oracle_params(**kw)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(400, ibatch=1, pad=0, precision=0.01)
    wfs, upix = sim.simulate_wfs(P, torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev), cm.FIELDS)
    out = sim.simulate_stochastic(P, wfs, upix, 0)
    loss = (out[0] ** 2).sum() * 1e-4 + (out[3] ** 2).sum() * 1e-3
    loss.backward()

    def L(p):
        w, u = lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, history={})
        o = lo.simulate_stochastic(p, w, u, dt=np.float64)
        return float((o[0] ** 2).sum() * 1e-4 + (o[3] ** 2).sum() * 1e-3)

    assert abs(float(loss) - L(op)) < 1e-4 * abs(L(op))
    for name, h in dict(Ab=1e-6, kb=1e-7, eField=1e-7, lifetime=1e-2, tran_diff=1e-11, long_diff=1e-11).items():
        base = getattr(op, name)
        fd = (L(op.replace(**{name: base + h})) - L(op.replace(**{name: base - h}))) / (2 * h)
        g = float(getattr(P, name).grad)
        # d/d(long_diff) is a cancelling sum of three nearly equal template terms (da + db + dc = 0): float32 leaves ~3e-3
        tol = 5e-3 if name == "long_diff" else GRAD_RTOL
        assert abs(g - fd) <= tol * abs(fd) + 1e-9, (name, g, fd)


def test_mc_mode_matches_oracle(torch_dev):
    import torch
    from larndsim_b200 import _lib, sim
    rng = np.random.default_rng(9)
    kw = dict(number_pix_neighbors=0, signal_length=150, mc_diff=True)
    for diff_in_current in (True, False):
        op = cm.oracle_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
        pp = cm.product_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
        tr = cm.small_batch(500, ibatch=1, pad=8, precision=0.01)
        rnd = rng.normal(size=(tr.shape[0], 3)).astype(np.float32)
        out_o, wfull_o, uniq_o = lo.simulate_parametrized(op, tr, cm.FIELDS, rnd, history={}, return_wfs=True)
        trd, rndd = torch.as_tensor(tr, device=torch_dev), torch.as_tensor(rnd, device=torch_dev)
        st = sim.mc_forward(pp, trd, cm.FIELDS, rndd, npix_capacity=len(uniq_o))
        assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
        valid = uniq_o >= 0
        w = st.wfs_full.cpu().numpy()
        scale = np.abs(wfull_o[valid]).max(axis=1, keepdims=True)
        assert (np.abs(w[valid] - wfull_o[valid]) <= 2e-5 * scale + 1e-2).all()
        out_p = sim.simulate_parametrized(pp, trd, cm.FIELDS, rnd=rndd, npix_capacity=len(uniq_o))
        got = [t.cpu().numpy() for t in out_p]
        assert len(got[0]) == len(out_o[0])
        assert np.array_equal(got[4], out_o[4]) and np.array_equal(got[7], out_o[7])
        assert np.abs(got[0] - out_o[0]).max() <= ADC_ATOL
    # gradients of both current models (diffusion inside the current model / plain exponentials + smeared z)
    for diff_in_current in (True, False):
        _mc_gradient_check(torch_dev, kw, diff_in_current, rng)


def _mc_gradient_check(torch_dev, kw, diff_in_current, rng):
    import torch
    from larndsim_b200 import _lib, sim
    op = cm.oracle_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
    pp = cm.product_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
    tr = cm.small_batch(300, ibatch=1, pad=0, precision=0.01)
    rnd = rng.normal(size=(tr.shape[0], 3)).astype(np.float32)
    _, wfull, uniq = lo.simulate_parametrized(op, tr, cm.FIELDS, rnd, history={}, return_wfs=True)
    G = (rng.uniform(0.5, 1.5, (len(uniq), 1)) * (1 + 0.5 * np.sin(np.arange(2000)[None, :] / 37.0))).astype(np.float32)
    G[uniq < 0] = 0
    trd = torch.as_tensor(tr, device=torch_dev)
    st = sim.mc_forward(pp, trd, cm.FIELDS, torch.as_tensor(rnd, device=torch_dev), npix_capacity=len(uniq))
    grad = sim.mc_backward(st, trd, torch.as_tensor(G, device=torch_dev)).cpu().numpy()

    def L(p):
        _, w, _ = lo.simulate_parametrized(p, tr, cm.FIELDS, rnd, dt=np.float64, pad_to=len(uniq), return_wfs=True)
        return float((w[:, 1:] * G.astype(np.float64)).sum())

    for name, h in dict(Ab=1e-6, eField=1e-7, lifetime=1e-2, long_diff=1e-11, shift_z=1e-6, shift_x=1e-6).items():
        base = getattr(op, name)
        fd = (L(op.replace(**{name: base + h})) - L(op.replace(**{name: base - h}))) / (2 * h)
        g = grad[_lib.PARAM_ORDER.index(name)]
        assert abs(g - fd) <= 5e-3 * abs(fd) + 1e-6 * abs(grad).max(), (name, g, fd)


def test_full_fixture_batch_and_size_independent_properties(torch_dev, acc_impl):
    """A complete simulate_test.sh batch (input_0, batch 1: 10 879 segments + reference padding) against the oracle,
    then properties that hold at any size: linearity in the charge scale and additivity over disjoint event sets."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=4, signal_length=100)
    bank = cm.synthetic_bank(32, 45, 45, 1950)
    arr, _ = cm.fixture_batches(0, 0.005)[1]
    tr = lo.pad_batch(arr, int(arr.shape[0] * 1.25 + 0.5), cm.FIELDS)
    op = cm.oracle_params(**kw)
    wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={}, response_cum=cm.synthetic_bank_cum(32, 45, 45, 1950),
                                               return_aux=True)
    pp, st = _run_lut(torch_dev, tr, bank, npix=len(uniq_o), **kw)
    assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o)
    _check_wfs(st.wfs_full.cpu().numpy(), full_o)
    _hits_equal(lo.simulate_stochastic(op, wfs_o, uniq_o), sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0))
    # linearity: MeVToElectrons x2 -> waveforms x2 (exactly, power of two)
    b, t = torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev)
    st2 = sim.lut_forward(pp.replace(MeVToElectrons=2 * pp.MeVToElectrons), b, t, cm.FIELDS, npix_capacity=st.npix)
    assert torch.allclose(st2.wfs_full, 2 * st.wfs_full, rtol=1e-5, atol=1e-3)
    # additivity: events are independent (the event id is part of the pixel key)
    ev = t[:, cm.FIELDS.index("eventID")]
    lo_half, hi_half = t[(ev >= 0) & (ev < 2)], t[ev >= 2]
    sa = sim.lut_forward(pp, b, lo_half, cm.FIELDS, n_events=3)
    sb = sim.lut_forward(pp, b, hi_half, cm.FIELDS, n_events=3)
    tot = {int(p): w for p, w in zip(st.unique_pixels.cpu().numpy(), st.wfs_full.cpu().numpy()) if p >= 0}
    for part in (sa, sb):
        for p, w in zip(part.unique_pixels.cpu().numpy(), part.wfs_full.cpu().numpy()):
            if p >= 0:
                _check_wfs(w[None, 1:], tot[int(p)][None, 1:])


def _oracle_batch(job):
    """Worker of test_all_prepared_inputs_forward_matches_oracle (module level: picklable for the process pool)."""
    ifile, ib = job
    arr, _ = cm.fixture_batches(ifile, 0.005)[ib]
    op = cm.oracle_params()
    bank = cm.synthetic_bank(32, 45, 45, 1950)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, arr, cm.FIELDS, history={})
    hits = lo.simulate_stochastic(op, wfs_o, uniq_o)
    real = uniq_o >= 0
    return ifile, ib, uniq_o, [np.asarray(h) for h in hits], wfs_o[real].astype(np.float32)


def test_all_prepared_inputs_forward_matches_oracle(torch_dev):
    """BASELINE config 2: every batch of all 22 un-shifted prepared_data inputs (optimize/simulate_test.sh settings: 0.005 cm,
    n = 4, L = 100, max_batch_len 50, no noise; ~0.84 M chopped segments) through simulate_wfs + simulate_stochastic, against
    the oracle evaluated in a process pool on the host cores: unique-pixel lists and hit sets (pixel, tick, event) bit for
    bit, ADCs to ADC_ATOL, waveform rows of real pixels to WFS_RTOL of the row maximum."""
    import concurrent.futures as cf
    import multiprocessing as mp
    import torch
    from larndsim_b200 import sim
    jobs = [(f, ib) for f in range(22) for ib in range(len(cm.fixture_batches(f, 0.005)))]
    cm.synthetic_bank(32, 45, 45, 1950)   # built once here, inherited by the forked workers
    pp = cm.product_params()
    bank_d = torch.as_tensor(cm.synthetic_bank(32, 45, 45, 1950), device=torch_dev)
    nseg = nhits = 0
    with cf.ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1), mp_context=mp.get_context("fork")) as pool:
        for ifile, ib, uniq_o, hits_o, wreal_o in pool.map(_oracle_batch, jobs, chunksize=1):
            arr, _ = cm.fixture_batches(ifile, 0.005)[ib]
            st = sim.lut_forward(pp, bank_d, torch.as_tensor(arr, device=torch_dev), cm.FIELDS, npix_capacity=len(uniq_o))
            sim.check_state(st)
            assert np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o), (ifile, ib)
            w = st.wfs_full[:, 1:].cpu().numpy()[uniq_o >= 0]
            scale = np.abs(wreal_o).max(axis=1, keepdims=True)
            assert (np.abs(w - wreal_o) <= WFS_RTOL * scale + 1e-3).all(), (ifile, ib)
            _hits_equal(hits_o, sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0))
            nseg += arr.shape[0]
            nhits += len(hits_o[0])
    assert nseg > 700000 and nhits > 10000, (nseg, nhits)


def test_spill_sized_batch_properties(torch_dev, monkeypatch):
    """Benchmark-shaped batch (synthetic spill, ~2 M segments — far beyond what the numpy oracle finishes in seconds):
    the class-sorted and the chunk kernels are two independent decompositions of the same sums and must agree to the
    waveform tolerance, forward and backward; the result must not depend on the order of the segments (the reference's
    segment_sum does not); hits of the two forward paths must be identical."""
    import torch
    import larndsim_b200 as lb
    from larndsim_b200 import sim, synthetic, dataio
    from larndsim_b200.consts import build_response_template
    dev = torch_dev
    params = cm.product_params(number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.01)
    raw, nev = synthetic.synthetic_raw_tracks(2_000_000, seed=77, precision=0.01)
    tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
    assert tracks.shape[0] > 1_500_000
    bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
    out = {}
    for impl in ("sorted", "chunk"):
        monkeypatch.setenv("LARND_ACC_IMPL", impl)
        st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
        fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
        g = sim.fee_backward(fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
        grad = sim.lut_backward(st, g)
        nv = int(fs.n_valid.item())
        out[impl] = (st.wfs_full, st.unique_pixels, [h[:, :nv].clone() for h in fs.hits], grad.double().cpu().numpy(), st.npix)
        assert int(st.counts.cpu()[2]) == 0
    ws, us, hs, gs, npix = out["sorted"]
    wc, uc, hc, gc, _ = out["chunk"]
    assert torch.equal(us, uc)
    real = us >= 0
    scale = ws[real].abs().amax(dim=1, keepdim=True)
    assert float(((ws[real][:, 1:] - wc[real][:, 1:]).abs() / (scale + 1e-30)).max()) < 2 * WFS_RTOL
    # hits: identical up to counted threshold-edge cases (the two waveform sets differ by ~1e-6 of the row maximum, so a
    # crossing that close to the threshold may move by a tick or appear / disappear: at most 3 in ~1e5 hits)
    if hs[0].shape[1] == hc[0].shape[1]:
        same = (hs[1][1] == hc[1][1]) & (hs[0][4] == hc[0][4])
        assert int((~same).sum()) <= 3
        assert float((hs[0][0] - hc[0][0])[same].abs().max()) <= ADC_ATOL
    else:
        assert abs(hs[0].shape[1] - hc[0].shape[1]) <= 3
    nz = np.abs(gs) > 0
    assert np.abs(gs[nz] - gc[nz]).max() / np.abs(gs[nz]).max() < 1e-4 and (np.abs(gs[nz] / gc[nz] - 1) < GRAD_RTOL).all()
    # permutation invariance (class-sorted path): same pixel list, same waveforms to rounding
    monkeypatch.setenv("LARND_ACC_IMPL", "sorted")
    perm = torch.randperm(tracks.shape[0], device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    sp = sim.lut_forward(params, bank, tracks[perm], synthetic.FIELDS, npix_capacity=npix, n_events=nev)
    assert torch.equal(sp.unique_pixels, us)
    assert float(((sp.wfs_full[real][:, 1:] - ws[real][:, 1:]).abs() / (scale + 1e-30)).max()) < 2 * WFS_RTOL
    # self-cleaning front end at spill size (tile kernels, garbage row 0 fed by every CTA): after fee_forward(clear_wfs=True)
    # the whole padded buffer is bitwise zero, and a forward pass into it without the memset reproduces the waveforms
    fz = sim.fee_forward(params, sp.wfs_full[:, 1:], sp.unique_pixels, None, compact=True, clear_wfs=True)
    assert int(torch.count_nonzero(sp.wfs_buf.view(torch.int32))) == 0
    s2 = sim.lut_forward(params, bank, tracks[perm], synthetic.FIELDS, npix_capacity=npix, n_events=nev,
                         out=(sp.unique_pixels, sp.wfs_buf), wfs_zero=True)
    assert float(((s2.wfs_full[real][:, 1:] - ws[real][:, 1:]).abs() / (scale + 1e-30)).max()) < 2 * WFS_RTOL
    f2 = sim.fee_forward(params, s2.wfs_full[:, 1:], s2.unique_pixels, None, compact=True)
    assert abs(int(f2.n_valid.item()) - int(fz.n_valid.item())) <= 3


_SPLIT_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/larnd-sim-jax_b200"); sys.path.insert(0, sys.argv[1] + "/tests")
import common as cm
from larndsim_b200 import sim, synthetic, dataio
from larndsim_b200.consts import build_response_template
dev = torch.device("cuda", 0)
params = cm.product_params(number_pix_neighbors=4, signal_length=100, electron_sampling_resolution=0.01)
raw, nev = synthetic.synthetic_raw_tracks(400_000, seed=5, precision=0.01)
tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
g = sim.fee_backward(fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
grad = sim.lut_backward(st, g)
np.savez(sys.argv[2], wfs=st.wfs_full.cpu().numpy(), uniq=st.unique_pixels.cpu().numpy(), grad=grad.double().cpu().numpy(),
         ticks=fs.ticks.cpu().numpy())
"""


def test_sorted_kernel_variants_agree(torch_dev, tmp_path):
    """The class-sorted kernels serve the tile table with one, two or three register-footprint variants (LARND_SORTED_SPLIT =
    0 / 1 / 2, read once per process): same runs, same per-window arithmetic, different launch partition — the waveforms
    must agree to the float32 reduction-order tolerance, the hit ticks exactly, the gradients to GRAD_RTOL."""
    import os
    import subprocess
    import sys
    res = {}
    for mode in ("0", "1", "2"):
        out = str(tmp_path / ("split%s.npz" % mode))
        env = dict(os.environ, LARND_SORTED_SPLIT=mode, LARND_ACC_IMPL="sorted", LARND_BWD_IMPL="sorted")
        subprocess.run([sys.executable, "-c", _SPLIT_SCRIPT, cm.ROOT, out], check=True, env=env, timeout=600)
        res[mode] = np.load(out)
    ref = res["0"]
    real = ref["uniq"] >= 0
    scale = np.abs(ref["wfs"][real]).max(axis=1, keepdims=True)
    for mode in ("1", "2"):
        r = res[mode]
        assert np.array_equal(r["uniq"], ref["uniq"])
        assert (np.abs(r["wfs"][real][:, 1:] - ref["wfs"][real][:, 1:]) <= 2 * WFS_RTOL * scale + 1e-3).all()
        assert (r["ticks"][real] != ref["ticks"][real]).sum() <= 3
        nz = np.abs(ref["grad"]) > 0
        assert (np.abs(r["grad"][nz] / ref["grad"][nz] - 1) < GRAD_RTOL).all()


_SHARD_SCRIPT = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "larnd-sim-jax_b200"))
import numpy as np, torch, torch.distributed as dist
import larndsim_b200 as lb
from larndsim_b200 import fit, parallel, synthetic
from larndsim_b200.consts import build_response_template
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
dev = torch.device("cuda", rank if ngpu >= world else 0)
torch.cuda.set_device(dev)
# one rank per GPU over NCCL when the box has two GPUs; on a one-GPU box both ranks share cuda:0 and the two small
# collectives go through gloo (NCCL refuses two ranks on one device) - the kernels and the two-phase reduction are the same
dist.init_process_group("nccl" if ngpu >= world else "gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                        rank=rank, world_size=world)
GEOM = os.path.join(sys.argv[1], "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
names = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
base = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
fields = synthetic.FIELDS
tracks_all, nev_all = synthetic.synthetic_tracks(19800, seed=5, precision=0.01)
p = lb.load_geometry_json(lb.build_params_class(list(names)), GEOM).replace(**base)
bank = build_response_template(synthetic.synthetic_response(25, 25, 1950), p, device=dev)
target = dict(Ab=0.83, kb=0.055, eField=0.52, lifetime=1.8e3, long_diff=5.0e-6, tran_diff=10e-6)
nominal = dict(Ab=0.8, kb=0.0486, eField=0.5, lifetime=2.2e3, long_diff=4.0e-6, tran_diff=8.8e-6)
loc, nev, _ = parallel.shard_tracks(tracks_all, fields, rank, world)
prob = fit.FitProblem.from_target_params(names, p, target, bank, torch.as_tensor(loc, device=dev), fields, nev)
loss, g = prob.loss_and_grads(nominal)
floss, fg = fit.FusedFitStep(prob)(nominal)          # the autograd-free chain, same sharding and collectives
if rank == 0:
    single = fit.FitProblem.from_target_params(names, p, target, bank, torch.as_tensor(tracks_all, device=dev), fields, nev_all, distributed=False)
    loss1, g1 = single.loss_and_grads(nominal)
    floss1, fg1 = fit.FusedFitStep(single)(nominal)
    np.savez(sys.argv[2], loss=float(loss), g=g.cpu().numpy(), loss1=float(loss1), g1=g1.cpu().numpy(), nseg=len(tracks_all),
             nloc=len(loc), backend=dist.get_backend(), floss=floss, fg=fg, floss1=floss1, fg1=fg1)
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_fit_step_equals_single_gpu(torch_dev, tmp_path):
    """BASELINE config 4 on hardware: the fit step of optimize/fit_test.sh --lut (n = 2, L = 150, ~19.8 k segments, mse_adc,
    six leaves) with the events sharded over two ranks must give the single-GPU loss and gradients: the MMD and charge
    terms are normalised by GLOBAL sums, so this holds only if the two-phase reduction (7 loss sums, then 6 gradients) is
    right.  Two ranks on two GPUs over NCCL when the box has them, else both on cuda:0 with the collectives over gloo."""
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "shard.npz")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", _SHARD_SCRIPT, cm.ROOT, out], env=env))
    for pr in procs:
        assert pr.wait(timeout=900) == 0
    r = np.load(out)
    assert 0 < r["nloc"] < r["nseg"]
    assert abs(r["loss"] - r["loss1"]) <= 1e-4 * abs(r["loss1"]), (r["loss"], r["loss1"])
    scale = np.abs(r["g1"]).max()
    assert (np.abs(r["g"] - r["g1"]) <= 2e-3 * np.abs(r["g1"]) + 1e-5 * scale).all(), (r["g"], r["g1"])
    assert (r["g1"] != 0).all()
    # the fused (autograd-free, dense-loss) step: same loss and gradients, sharded and not
    for lf, gf in ((r["floss1"], r["fg1"]), (r["floss"], r["fg"])):
        assert abs(lf - r["loss1"]) <= 2e-4 * abs(r["loss1"]), (lf, r["loss1"])
        assert (np.abs(gf - r["g1"]) <= 3e-3 * np.abs(r["g1"]) + 1e-5 * scale).all(), (gf, r["g1"])


def test_fit_and_scan_drivers(torch_dev):
    """The reference's convergence criterion (tests/test_fit_convergence.py:21-45): no NaN and the mean of the last 5
    losses is below the mean of the first 5; plus a 3x3 likelihood scan whose minimum sits at the nominal point."""
    import os
    import sys
    sys.path.insert(0, os.path.join(cm.ROOT, "examples"))
    import fit_demo
    import scan_2d
    hist, target = fit_demo.run_fit(("Ab", "eField"), iterations=16, n_segments=12000, lr=0.01, device=torch_dev, verbose=False)
    losses = np.array([h[0] for h in hist])
    assert np.isfinite(losses).all() and losses[-5:].mean() < losses[:5].mean()
    theta0, theta1 = hist[0][1], hist[-1][1]
    tgt = np.array([target["Ab"], target["eField"]])
    assert np.abs(theta1 - tgt).sum() < np.abs(theta0 - tgt).sum()          # moved towards the target parameters
    a1, a2, out = scan_2d.run_scan("Ab", "lifetime", grid=3, n_segments=8000, device=torch_dev)
    assert np.isfinite(out).all()
    assert np.unravel_index(np.argmin(out[..., 0]), (3, 3))[0] in (0, 1)        # nominal Ab = 0.8 lies between grid points 0 and 1


def test_device_chop_tracks_is_bit_identical_to_numpy(torch_dev):
    """k_chop_* (csrc/chop.cu) against the oracle's restatement of chop_tracks (optimize/dataio.py:63-106) on every raw
    fixture row, at the three sampling resolutions the reference scripts use; plus degenerate rows (zero length)."""
    import torch
    from larndsim_b200 import dataio
    seg = lo.swap_xz_structured(np.load(cm.GOLD + "/segments_input_0.npz")["segments"])
    raw = lo.structured_to_f32(seg)
    raw = np.concatenate([raw, raw[:3]])
    c = cm.FIELDS.index
    for ax in "xyz":                       # a zero-length row: one piece, direction 0/1e-10
        raw[-1, c(ax + "_end")] = raw[-1, c(ax + "_start")]
    raw[-2, c("x_end")] = raw[-2, c("x_start")] + np.float32(0.02)   # exactly 2 pieces at 0.01 on one axis
    raw[-2, c("y_end")] = raw[-2, c("y_start")]
    raw[-2, c("z_end")] = raw[-2, c("z_start")]
    t = torch.as_tensor(raw, device=torch_dev)
    for prec in (0.005, 0.01, 0.05):
        ref = lo.chop_tracks(raw, cm.FIELDS, prec)
        got = dataio.chop_tracks(t, cm.FIELDS, prec).cpu().numpy()
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "chopped rows differ at precision %g" % prec
        off = dataio.chop_offsets(t, cm.FIELDS, prec).cpu().numpy()
        assert off[-1] == ref.shape[0] and off[0] == 0
    # capacity too small: nothing is written, the total is still reported
    out = torch.full((10, raw.shape[1]), -7.0, device=torch_dev)
    dataio.chop_tracks(t, cm.FIELDS, 0.01, out=out)
    assert (out == -7.0).all()
    assert dataio.chop_tracks(t[:0], cm.FIELDS, 0.01).shape[0] == 0


def test_template_bank_builder_matches_oracle(torch_dev):
    """k_build_bank (load_lut's Gaussian template bank, consts_jax.py:427-447) against the oracle's scipy construction."""
    from larndsim_b200.consts import build_response_template
    op, pp = cm.oracle_params(), cm.product_params()
    resp = oc.synthetic_response(7, 6, 1950)
    ref = oc.build_response_template(resp, op)
    got = build_response_template(resp, pp, device=torch_dev).cpu().numpy()
    assert got.shape == ref.shape == (100, 7, 6, 1950)
    assert np.array_equal(got[0], resp)                      # row 0 is the raw response, bit for bit
    scale = np.abs(ref).max(axis=-1, keepdims=True) + 1e-30
    assert (np.abs(got - ref) / scale).max() < 2e-6


def test_threefry_normals_match_the_jax_restatement(torch_dev):
    """csrc/rng.cu against oracle/jax_random.py (itself pinned to values published by JAX): normals for both counter
    layouts, odd sizes and an (N,3) shape, and the complete noise buffer of get_adc_values."""
    from larndsim_b200 import jrandom
    from oracle import jax_random as jr
    for part in (True, False):
        for seed, shape in ((0, (1,)), (42, (7,)), (5, (1001,)), (9, (333, 3))):
            ref = jr.normal(jr.key(seed), shape, part)
            got = jrandom.normal(jrandom.key(seed), shape, torch_dev, part).cpu().numpy()
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 2e-6 * (1 + np.abs(ref).max()), (part, seed)
        ref = jr.fee_noise(3, 37, 10, part)
        got = jrandom.fee_noise(jrandom.key(3), 37, 10, torch_dev, part).cpu().numpy()
        assert np.abs(got - ref).max() <= 4e-6
    assert abs(float(jrandom.normal(jrandom.key(42), (), torch_dev, True)) - (-0.028304616)) < 1e-6   # value printed in the JAX docs


def test_noisy_fee_uses_the_jax_noise_stream(torch_dev):
    """simulate_stochastic with noise switched on draws jax.random.key(rngseed)'s normals: same hits as the oracle fed with
    the restated jax stream."""
    import torch
    from larndsim_b200 import sim
    from oracle import jax_random as jr
    kw = dict(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0, UNCORRELATED_NOISE_CHARGE=500.0)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    tr = cm.small_batch(500, ibatch=2, pad=4, precision=0.01)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={})
    z = jr.fee_noise(11, len(uniq_o), 10, True).reshape(31, len(uniq_o))
    noise = dict(base=z[0], extra=z[1:11], **{"pass": z[11:21], "fail": z[21:31]})
    out_o = lo.simulate_stochastic(op, wfs_o, uniq_o, noise=noise)
    w, u = torch.as_tensor(wfs_o, device=torch_dev), torch.as_tensor(uniq_o, device=torch_dev)
    out_p = sim.simulate_stochastic(pp, w, u, 11)
    got = [t.detach().cpu().numpy() for t in out_p]
    assert len(got[0]) == len(out_o[0]) and len(got[0]) > 0
    assert np.array_equal(got[4], out_o[4]) and np.array_equal(got[7], out_o[7])
    assert np.abs(got[0] - out_o[0]).max() < 5e-3


def test_probabilistic_front_end_forward_matches_oracle(torch_dev):
    """k_prob_* (get_adc_values_average_noise_vmap, fee_jax.py:334-461) against oracle/prob_fee.py.  The log-probabilities
    contain log(Phi(a) - Phi(b)) of nearly equal arguments, which is float32 noise wherever the probability itself is
    ~1e-7 (in the reference too), so values are compared as probabilities and through get_average_hit_values."""
    import torch
    from larndsim_b200 import fee, sim
    from oracle import prob_fee as pf
    kw = dict(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    tr = cm.small_batch(600, ibatch=2, pad=2, precision=0.01)
    wfs_o, uniq_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={})
    amp = np.abs(wfs_o).sum(axis=1)
    sel = np.concatenate([np.argsort(-amp)[:9], np.argsort(amp)[:3]])
    w = np.ascontiguousarray(wfs_o[sel])
    lp_o, q_o, top_o = pf.get_adc_values_average_noise(op, w, return_state=True)
    lp, qd, top = fee.get_adc_values_average_noise_vmap(pp, torch.as_tensor(w, device=torch_dev), return_top_ticks=True)
    lp, qd, top = lp.cpu().numpy(), qd.cpu().numpy(), top.cpu().numpy()
    assert lp.shape == lp_o.shape == (len(sel), 10, 1999)
    assert np.abs(qd - q_o).max() <= 1e-6 * np.abs(q_o).max() + 1e-2          # esperance_value: plain float32 arithmetic
    assert np.abs(np.exp(lp) - np.exp(lp_o)).max() < 5e-4
    et_o, eq_o, lam_o = pf.get_average_hit_values(np.exp(lp_o), q_o)
    et, eq, lam = [t.cpu().numpy() for t in fee.get_average_hit_values(torch.as_tensor(np.exp(lp)), torch.as_tensor(qd))]
    assert np.abs(lam - lam_o).max() < 2e-3
    big = lam_o > 1e-2
    assert big.sum() >= 9 and np.abs(et - et_o)[big].max() < 0.05 and (np.abs(eq - eq_o)[big] / np.abs(eq_o[big])).max() < 1e-3
    strong = lam_o[:, 0] > 0.5
    assert np.array_equal(np.sort(top[strong, 0], axis=1), np.sort(top_o[strong, 0], axis=1))   # beam of the first hit
    # simulate_probabilistic wrapper: shapes and the digitised distribution
    out = sim.simulate_probabilistic(pp, torch.as_tensor(w, device=torch_dev), torch.as_tensor(uniq_o[sel], device=torch_dev))
    ref = pf.simulate_probabilistic(op, w, uniq_o[sel])
    assert np.abs(out[0].cpu().numpy() - ref[0]).max() < 2e-3 and np.array_equal(out[4].cpu().numpy(), ref[4])
    assert np.allclose(out[1].cpu().numpy(), ref[1]) and np.allclose(out[2].cpu().numpy(), ref[2])


def test_probabilistic_front_end_gradient_matches_finite_differences(torch_dev):
    """VJP of the beam search w.r.t. the waveforms (k_prob_bwd_*) against central differences of the float64 oracle with
    the discrete choices (beam ticks, stop flags) pinned to the kernel's forward.  A small positive tilt keeps the running
    sum strictly increasing so that the running maximum has a unique arg-max (no kinks)."""
    import torch
    from larndsim_b200 import fee
    from oracle import prob_fee as pf
    kw = dict(number_pix_neighbors=1, signal_length=100, RESET_NOISE_CHARGE=900.0)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 15, 15, 1950)
    tr = cm.small_batch(600, ibatch=2, pad=2, precision=0.01)
    wfs_o, _ = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={})
    sel = np.argsort(-np.abs(wfs_o).sum(axis=1))[:3]
    w = np.ascontiguousarray(wfs_o[sel] + 2.0).astype(np.float32)
    rng = np.random.default_rng(2)
    wt = torch.as_tensor(w, device=torch_dev).requires_grad_(True)
    lp, qd = fee.get_adc_values_average_noise_vmap(pp, wt)
    _, _, top = fee.get_adc_values_average_noise_vmap(pp, wt.detach(), return_top_ticks=True)
    # loss: probability-weighted sums, the kind of quantity the probabilistic losses build (bounded weights)
    A = rng.uniform(0.5, 1.5, size=lp.shape).astype(np.float32)
    B = rng.uniform(-1, 1, size=lp.shape).astype(np.float32) * 1e-3
    loss = (torch.exp(lp) * torch.as_tensor(A, device=torch_dev)).sum() + (torch.exp(lp) * qd * torch.as_tensor(B, device=torch_dev)).sum()
    loss.backward()
    g = wt.grad.cpu().numpy().astype(np.float64)
    tops = top.cpu().numpy()

    def L(x):
        l, q = pf.get_adc_values_average_noise(op, x, dt=np.float64, force_tops=tops)
        return float((np.exp(l) * A).sum() + (np.exp(l) * q * B).sum())

    assert abs(float(loss) - L(w.astype(np.float64))) < 2e-3 * abs(L(w.astype(np.float64)))
    checked = 0
    for pix in range(len(sel)):
        peak = int(np.argmax(w[pix]))
        for t in (peak - 30, peak - 5, peak, peak + 7, peak + 40, 100):
            h = 0.05 * max(1.0, abs(w[pix, t]))
            xp, xm = w.astype(np.float64).copy(), w.astype(np.float64).copy()
            xp[pix, t] += h
            xm[pix, t] -= h
            fd = (L(xp) - L(xm)) / (2 * h)
            assert abs(g[pix, t] - fd) <= 2e-2 * abs(fd) + 2e-3 * np.abs(g[pix]).max(), (pix, t, g[pix, t], fd)
            checked += 1
    assert checked == 18 and np.abs(g).max() > 0


def test_mmd_kernel_matches_dense_evaluation(torch_dev):
    """k_rbf_field (tiled, far tiles skipped) against the dense float64 evaluation of the reference's mmd
    (losses_jax.py:14-39) on event-offset hit clouds, values and gradients w.r.t. positions and weights."""
    import torch
    from larndsim_b200 import losses
    rng = np.random.default_rng(4)
    def cloud(n, nev):
        ev = np.sort(rng.integers(0, nev, n))
        return np.stack([rng.normal(0, 3, n) + ev * 1e5, rng.normal(0, 3, n), rng.normal(0, 3, n)], 1), rng.uniform(0.5, 2, n)
    (x, px), (y, py) = cloud(700, 5), cloud(650, 5)
    # the event offset of 1e5 per event costs float32 position bits (in the reference too): compare on the float32-rounded inputs
    x, px, y, py = [a.astype(np.float32).astype(np.float64) for a in (x, px, y, py)]
    sigma = 1.3
    xt = torch.tensor(x, dtype=torch.float32, device=torch_dev, requires_grad=True)
    pt = torch.tensor(px, dtype=torch.float32, device=torch_dev, requires_grad=True)
    yt, qt = torch.tensor(y, dtype=torch.float32, device=torch_dev), torch.tensor(py, dtype=torch.float32, device=torch_dev)
    val = losses.mmd(xt, yt, pt, qt, sigma)
    val.backward()
    xd = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    pd = torch.tensor(px, dtype=torch.float64, requires_grad=True)
    yd, qd = torch.tensor(y, dtype=torch.float64), torch.tensor(py, dtype=torch.float64)
    K = lambda a, b: torch.exp(-((a[:, None, :] - b[None, :, :]) ** 2).sum(-1) / (2 * sigma ** 2))
    ref = ((K(xd, xd) * pd[:, None] * pd[None, :]).sum() / pd.sum() ** 2 + (K(yd, yd) * qd[:, None] * qd[None, :]).sum() / qd.sum() ** 2
           - 2 * (K(xd, yd) * pd[:, None] * qd[None, :]).sum() / (pd.sum() * qd.sum()))
    ref.backward()
    assert abs(float(val) - float(ref)) < 2e-5 * max(abs(float(ref)), 1e-3) + 1e-7
    gx, gp = xt.grad.cpu().double(), pt.grad.cpu().double()
    assert (gx - xd.grad).abs().max() < 1e-4 * xd.grad.abs().max() + 1e-9
    assert (gp - pd.grad).abs().max() < 1e-4 * pd.grad.abs().max() + 1e-9


def test_simulate_drift_new_view_and_mirror_helpers(torch_dev):
    """The reference-shaped per-segment arrays rebuilt from the prepare kernel's records (sim.simulate_drift_new) and the
    detsim mirror helpers (get_bin_shifts, get_pixels, density_2d, shift_tracks, generate_electrons) against the oracle."""
    import torch
    from larndsim_b200 import detsim, jrandom, sim
    from oracle import jax_random as jr
    kw = dict(number_pix_neighbors=2, signal_length=100)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    tr = cm.small_batch(700, ibatch=1, pad=6, precision=0.01)
    d = lo.simulate_drift_new(op, tr, cm.FIELDS)
    t = torch.as_tensor(tr, device=torch_dev)
    out = sim.simulate_drift_new(pp, t, cm.FIELDS)
    names = ("main_pixels", "pixels", "nelectrons", "t0_after_diff", "long_diff", "currents_idx", "pIDs_neigh", "currents_idx_neigh",
             "nelectrons_neigh", "t0_neigh")
    for name, got in zip(names, out):
        ref, g = d[name], got.cpu().numpy()
        assert g.shape == ref.shape, name
        if ref.dtype.kind in "iu":
            assert np.array_equal(g, ref), name
        else:
            assert np.allclose(g, ref, rtol=3e-6, atol=1e-6 * np.abs(ref).max()), name
    # helpers on the drifted tracks
    drifted = torch.as_tensor(d["tracks"], device=torch_dev)
    assert np.array_equal(detsim.get_bin_shifts(pp, drifted, cm.FIELDS).cpu().numpy(), d["bins_pitches"])
    assert np.array_equal(detsim.get_pixels(pp, drifted, cm.FIELDS)[:, 2, 2].cpu().numpy(), d["main_pixels"])
    edges = torch.as_tensor(oc.linspace_jnp(np.float32(-2 * op.pixel_pitch / 10), np.float32(3 * op.pixel_pitch / 10), 6), device=torch_dev)
    w2d = detsim.density_2d(edges, torch.as_tensor(d["x0"], device=torch_dev), torch.as_tensor(d["y0"], device=torch_dev),
                            torch.as_tensor(d["sigma_t"], device=torch_dev)).cpu().numpy()
    ok = d["sigma_t"] > 0
    assert np.abs(w2d[ok] - (d["wx"][:, :, None] * d["wy"][:, None, :])[ok]).max() < 1e-6
    assert np.allclose(sim.shift_tracks(pp.replace(shift_x=0.3, shift_z=-0.2), t, cm.FIELDS).cpu().numpy(),
                       lo.shift_tracks(op.replace(shift_x=0.3, shift_z=-0.2), tr, cm.FIELDS, np.float32))
    k1 = jrandom.split(jrandom.key(5), 2)[0]
    el = detsim.generate_electrons(drifted, cm.FIELDS, k1, apply_long_diffusion=False).cpu().numpy()
    rnd = jr.normal(jr.split(jr.key(5), 2)[0], (len(tr), 3))
    c = cm.FIELDS.index
    assert np.allclose(el[:, c("x")], d["tracks"][:, c("x")] + rnd[:, 0] * d["tracks"][:, c("tran_diff")], atol=1e-5)
    assert np.array_equal(el[:, c("z")], d["tracks"][:, c("z")])


def test_deterministic_mode_is_bitwise_reproducible_and_reference_accurate(torch_dev):
    """sim.set_deterministic / lut_forward(deterministic=True): the counterpart of --xla_gpu_deterministic_ops
    (optimize/example_run.py:44-47).  Five runs of a 60 k-segment batch (where the default kernels' float reductions do
    reorder) give bit-identical waveforms and hit lists; the result meets the same bars against the oracle."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=4, signal_length=100)
    bank = torch.as_tensor(cm.synthetic_bank(32, 45, 45, 1950), device=torch_dev)
    # event ids stay local per source batch: offset them so that the concatenation is one batch of distinct events
    off, parts = 0, []
    for i, src in enumerate([cm.fixture_batches(0, 0.005)[j][0] for j in range(4)] + [cm.fixture_batches(1, 0.005)[j][0] for j in range(2)]):
        p = src.copy()
        p[:, 0] += off
        off = int(p[:, 0].max()) + 1
        parts.append(p)
    tr = torch.as_tensor(np.concatenate(parts), device=torch_dev)
    assert tr.shape[0] > 50000
    pp = cm.product_params(**kw)
    ref = None
    for rep in range(5):
        st = sim.lut_forward(pp, bank, tr, cm.FIELDS, deterministic=True)
        hits = sim.simulate_stochastic(pp, st.wfs_full[:, 1:], st.unique_pixels, 0)
        cur = [st.wfs_full.clone()] + [h.clone() for h in hits]
        if ref is None:
            ref = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref, cur)), rep
    # accuracy: the deterministic waveforms against the default kernels (both within the oracle bar of each other)
    st0 = sim.lut_forward(pp, bank, tr, cm.FIELDS, npix_capacity=st.npix)
    w, w0 = ref[0].cpu().numpy(), st0.wfs_full.cpu().numpy()
    real = st.unique_pixels.cpu().numpy() >= 0
    scale = np.abs(w0[real]).max(axis=1, keepdims=True)
    assert (np.abs(w - w0)[real][:, 1:] <= WFS_RTOL * scale + 1e-3).all()
    _hits_equal([h.cpu().numpy() for h in sim.simulate_stochastic(pp, st0.wfs_full[:, 1:], st0.unique_pixels, 0)], ref[1:])
    # ... and against the oracle on a small batch through the same path
    small = cm.small_batch(600, pad=8, precision=0.01)
    op = cm.oracle_params(**kw)
    wo, uo = lo.simulate_wfs(op, cm.synthetic_bank(32, 45, 45, 1950), small, cm.FIELDS, history={})
    sd = sim.lut_forward(pp, bank, torch.as_tensor(small, device=torch_dev), cm.FIELDS, npix_capacity=len(uo), deterministic=True)
    assert np.array_equal(sd.unique_pixels.cpu().numpy(), uo)
    _check_wfs(sd.wfs_full[:, 1:].cpu().numpy(), wo, uniq=uo)


@pytest.mark.parametrize("cfg", [dict(number_pix_neighbors=4, signal_length=100), dict(number_pix_neighbors=2, signal_length=150)])
@pytest.mark.parametrize("bwd", ["chunk", "sorted"])
def test_steps_backward_equals_dense_backward(torch_dev, cfg, bwd, monkeypatch):
    """sim.hits_backward (larnd_fee_backward_steps -> larnd_lut_backward_steps: the front end's VJP as a step list per row, the
    correlations as running-sum differences) against fee_backward + lut_backward (dense gradient array) for the same upstream
    gradient, on a batch with segments next to the anode (windows sticking out at the low end) and outside every TPC; all 15
    leaves, both kernel families."""
    import torch
    from larndsim_b200 import _lib, sim
    monkeypatch.setenv("LARND_BWD_IMPL", bwd)
    rng = np.random.default_rng(5)
    tr = cm.small_batch(1500, ibatch=1, pad=0, precision=0.01)
    c = cm.FIELDS.index
    near = tr[:300].copy()
    shift = 30.45 - np.abs(near[:, c("z")]).max()
    for col in ("z", "z_start", "z_end"):
        near[:, c(col)] += np.sign(near[:, c(col)]) * shift
    far = tr[300:330].copy()
    far[:, c("x")] += 100.0
    allr = np.concatenate([tr, near, far, cm.small_batch(1200, ifile=1, ibatch=0, pad=6, precision=0.01)])
    pp = cm.product_params(**cfg)
    nbins = 10 * cfg["number_pix_neighbors"] + 5
    bank = torch.as_tensor(cm.synthetic_bank(48, nbins, nbins, 1950), device=torch_dev)
    st = sim.lut_forward(pp, bank, torch.as_tensor(allr, device=torch_dev), cm.FIELDS)
    fs = sim.fee_forward(pp, st.wfs_full[:, 1:], st.unique_pixels, None, compact=False)
    assert int((fs.ticks < 1997).sum()) > 80
    g_adc = torch.as_tensor(rng.normal(size=tuple(fs.adc.shape)).astype(np.float32), device=torch_dev)
    g_adc = g_adc * (st.unique_pixels >= 0).unsqueeze(1)          # hits of invalid rows never reach a loss (parse_output)
    dense = sim.lut_backward(st, sim.fee_backward(fs, g_adc)).cpu().numpy().astype(np.float64)
    steps = sim.hits_backward(st, fs, g_adc).cpu().numpy().astype(np.float64)
    used = [i for i, n in enumerate(_lib.PARAM_ORDER) if n not in ("alpha", "beta", "R_param", "vdrift")]   # Birks model
    assert (dense[used] != 0).all()
    err = np.abs(steps - dense) / np.maximum(np.abs(dense), 1e-30)
    assert (err[used] <= 1e-3).all(), dict(zip(_lib.PARAM_ORDER, zip(steps, dense)))
    # raw-charge form (gradient w.r.t. get_adc_values' integrated charge)
    dense_q = sim.lut_backward(st, sim.fee_backward(fs, g_adc, raw_charge=True)).cpu().numpy().astype(np.float64)
    steps_q = sim.hits_backward(st, fs, g_adc, raw_charge=True).cpu().numpy().astype(np.float64)
    assert (np.abs(steps_q - dense_q)[used] <= 1e-3 * np.abs(dense_q)[used]).all()


@pytest.mark.gpu
def test_autograd_hands_the_front_end_vjp_over_as_step_events(torch_dev, monkeypatch):
    """simulate_wfs -> simulate_stochastic chained under torch.autograd: from STEPS_AUTOGRAD_MIN_SEGMENTS on, the front end's
    node hands its VJP to simulate_wfs' node as step events (sim._StepsLink) instead of a dense (Npix, Nticks) gradient.  Same
    gradients as the dense route; with a second consumer of the waveforms the placeholder is summed away by autograd and the
    dense front-end gradient is added back."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=2, signal_length=150)
    names = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
    bank = torch.as_tensor(cm.synthetic_bank(32, 25, 25, 1950), device=torch_dev)
    tr = torch.as_tensor(cm.small_batch(1200, ibatch=1, pad=0, precision=0.01), device=torch_dev)

    def grads(min_segments, second_consumer):
        monkeypatch.setattr(sim, "STEPS_AUTOGRAD_MIN_SEGMENTS", min_segments)
        P = cm.product_params(grad=names, **kw)
        calls = {"steps": 0, "dense": 0}
        hb, lb_ = sim.hits_backward, sim.lut_backward
        monkeypatch.setattr(sim, "hits_backward", lambda *a, **k: (calls.__setitem__("steps", calls["steps"] + 1), hb(*a, **k))[1])
        monkeypatch.setattr(sim, "lut_backward", lambda *a, **k: (calls.__setitem__("dense", calls["dense"] + 1), lb_(*a, **k))[1])
        wfs, upix = sim.simulate_wfs(P, bank, tr, cm.FIELDS)
        out = sim.simulate_stochastic(P, wfs, upix, 0)
        loss = (out[0] ** 2).sum() * 1e-4 + (out[3] ** 2).sum() * 1e-3
        if second_consumer:
            loss = loss + (wfs[1:] ** 2).sum() * 1e-6
        loss.backward()
        monkeypatch.setattr(sim, "hits_backward", hb)
        monkeypatch.setattr(sim, "lut_backward", lb_)
        return np.array([float(getattr(P, n).grad) for n in names]), calls, float(loss)

    for second in (False, True):
        g_dense, c_dense, l_dense = grads(1 << 40, second)
        g_steps, c_steps, l_steps = grads(0, second)
        assert c_dense == {"steps": 0, "dense": 1}
        assert c_steps == ({"steps": 0, "dense": 1} if second else {"steps": 1, "dense": 0})
        assert abs(l_dense - l_steps) <= 1e-5 * abs(l_dense)
        tol = np.where(np.array(names) == "long_diff", 5e-3, 1e-3)   # cancelling template sums in float32 (see above)
        assert (np.abs(g_steps - g_dense) <= tol * np.abs(g_dense) + 1e-12).all(), (second, g_steps, g_dense)


def test_fused_raw_prepare_is_bit_identical_to_chop_then_prepare(torch_dev):
    """larnd_lut_prepare_raw (chop_tracks fused into the prepare kernel: the chopped batch is never written) against
    chop_tracks -> pad_batch -> larnd_lut_prepare: every per-segment record bit for bit (including the invalid rows of the
    padded slots), the same pixel list and hits, the same gradients through the step-event backward; degenerate raw rows
    (zero length, exactly two pieces); too few segment slots are flagged on the device."""
    import torch
    from larndsim_b200 import _lib, dataio, sim
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = torch.as_tensor(cm.synthetic_bank(32, 25, 25, 1950), device=torch_dev)
    pp = cm.product_params(**kw).replace(electron_sampling_resolution=0.01)
    seg = lo.swap_xz_structured(np.load(os.path.join(cm.GOLD, "segments_input_2.npz"))["segments"])
    rows, gids = lo.make_batches(seg, 50.0)[0]
    raw = lo.batch_array(seg, rows, gids, cm.FIELDS, False, 0.01)
    raw = np.concatenate([raw, raw[:2]])
    c = cm.FIELDS.index
    for ax in "xyz":
        raw[-1, c(ax + "_end")] = raw[-1, c(ax + "_start")]
    raw[-2, c("x_end")] = raw[-2, c("x_start")] + np.float32(0.02)
    raw[-2, c("y_end")] = raw[-2, c("y_start")]
    raw[-2, c("z_end")] = raw[-2, c("z_start")]
    rt = torch.as_tensor(raw, device=torch_dev)
    for prec in (0.01, 0.05):
        chopped = dataio.chop_tracks(rt, cm.FIELDS, prec)
        total = chopped.shape[0]
        for n_seg in (total, total + 37):
            padded = dataio.pad_batch(chopped, n_seg, cm.FIELDS)
            nev = sim.n_events_of(rt, cm.FIELDS)
            ref = sim.lut_forward(pp, bank, padded, cm.FIELDS, n_events=nev)
            got = sim.lut_forward(pp, bank, rt, cm.FIELDS, n_events=nev, npix_capacity=ref.npix, raw=(prec, n_seg))
            assert got.n == ref.n == n_seg
            assert int(got.counts[2].item()) == 0
            ra, rb = sim.record_fields(ref), sim.record_fields(got)
            for name in ra:
                assert torch.equal(ra[name].view(torch.int32), rb[name].view(torch.int32)), (prec, n_seg, name)
            assert torch.equal(got.unique_pixels, ref.unique_pixels)
            scale = ref.wfs_full.abs().amax(dim=1, keepdim=True) + 1e-30
            assert bool(((got.wfs_full - ref.wfs_full).abs() <= 2 * WFS_RTOL * scale + 1e-3).all())
            fa = sim.fee_forward(pp, ref.wfs_full[:, 1:], ref.unique_pixels, None, compact=False)
            fb = sim.fee_forward(pp, got.wfs_full[:, 1:], got.unique_pixels, None, compact=False)
            assert torch.equal(fa.ticks, fb.ticks)
            g_adc = fa.adc * (ref.unique_pixels >= 0).unsqueeze(1)
            ga = sim.hits_backward(ref, fa, g_adc).cpu().numpy().astype(np.float64)
            gb = sim.hits_backward(got, fb, g_adc).cpu().numpy().astype(np.float64)
            assert (np.abs(ga - gb) <= 1e-4 * np.abs(ga) + 1e-12).all(), (ga, gb)
    # the reference-facing entry: hits of the fused path == hits of the chop-first path
    ha = dataio.simulate_from_raw(pp, bank, raw, cm.FIELDS, precision=0.01, fused=False)
    hb = dataio.simulate_from_raw(pp, bank, raw, cm.FIELDS, precision=0.01, fused=True)
    assert len(ha[0]) == len(hb[0]) > 0
    for k in (4, 6, 7):
        assert torch.equal(ha[k], hb[k])
    assert float((ha[0] - hb[0]).abs().max()) <= ADC_ATOL
    # fewer slots than pieces: flagged, nothing simulated
    st = sim.lut_forward(pp, bank, rt, cm.FIELDS, n_events=nev, npix_capacity=ref.npix, raw=(0.01, 100))
    assert int(st.counts[2].item()) & 8
    with pytest.raises(_lib.LarndError):
        sim.check_state(st)


def test_hits_only_pipeline_cleans_its_waveform_buffer(torch_dev):
    """sim.simulate_hits: simulate_wfs + simulate_stochastic over a persistent arena.  The front-end kernel zeroes every sample
    it has read (LARND_FEE_CLEAR_WFS), so (a) the padded waveform buffer is bitwise all-zero after every call — including the
    garbage row 0, the garbage column and the padding columns — and the next accumulate may skip its memset
    (LARND_FLAG_WFS_ZERO); (b) repeated calls on different batches return exactly the hits of the two-call API; (c) FEE
    outputs with and without cleaning are identical."""
    import torch
    from larndsim_b200 import dataio, sim
    kw = dict(number_pix_neighbors=2, signal_length=100)
    bank = torch.as_tensor(cm.synthetic_bank(32, 25, 25, 1950), device=torch_dev)
    pp = cm.product_params(**kw)
    batches = [cm.small_batch(900, ifile=0, ibatch=0, pad=5, precision=0.01), cm.small_batch(1300, ifile=1, ibatch=1, pad=0, precision=0.01),
               cm.small_batch(700, ifile=2, ibatch=0, pad=9, precision=0.01)]
    arena = sim.HitsArena()
    npix = 1024
    for rep in range(2):
        for ib, tr in enumerate(batches):
            t = torch.as_tensor(tr, device=torch_dev)
            nev = sim.n_events_of(t, cm.FIELDS)
            wfs, upix = sim.simulate_wfs(pp, bank, t, cm.FIELDS, npix_capacity=npix, n_events=nev)
            ref = sim.simulate_stochastic(pp, wfs, upix, ib)
            got = sim.simulate_hits(pp, bank, t, cm.FIELDS, rngseed=ib, npix_capacity=npix, n_events=nev, arena=arena)
            assert arena.clean and int(torch.count_nonzero(arena.wfs.view(torch.int32))) == 0, (rep, ib)
            assert len(got[0]) == len(ref[0]) > 0
            for k in (4, 6, 7):
                assert torch.equal(got[k], ref[k])
            assert float((got[0] - ref[0]).abs().max()) <= ADC_ATOL
    # same FEE outputs with and without the cleaning stores
    st = sim.lut_forward(pp, bank, torch.as_tensor(batches[0], device=torch_dev), cm.FIELDS, npix_capacity=npix)
    keep = st.wfs_buf.clone()
    fa = sim.fee_forward(pp, st.wfs_full[:, 1:], st.unique_pixels, None, compact=False)
    assert torch.equal(st.wfs_buf, keep)
    fb = sim.fee_forward(pp, st.wfs_full[:, 1:], st.unique_pixels, None, compact=False, clear_wfs=True)
    assert int(torch.count_nonzero(st.wfs_buf.view(torch.int32))) == 0
    assert torch.equal(fa.adc, fb.adc) and torch.equal(fa.ticks, fb.ticks) and torch.equal(fa.saved, fb.saved)
    with pytest.raises(ValueError):
        sim.fee_forward(pp, keep[:, 1:1950].double(), st.unique_pixels, None, clear_wfs=True)
