"""Parity of the stage-by-stage ("stream-form") operators — quench, drift, simulate_signals, simulate_drift, current_mc,
accumulate_signals_parametrized, simulate_signals_parametrized — against the numpy oracle and against the fused kernels.
GPU only.  Same bars as test_gpu_parity.py: integers bit-exact, waveforms within 5e-6 of the row maximum, gradients within
2e-3 of float64 central differences of the oracle."""
import numpy as np
import pytest

import common as cm
from oracle import consts as oc
from oracle import larnd_oracle as lo

pytestmark = pytest.mark.gpu

WFS_RTOL = 5e-6
GRAD_RTOL = 2e-3


@pytest.fixture(scope="module")
def torch_dev(cuda_lib):
    import torch
    return torch.device("cuda", 0)


@pytest.mark.parametrize("mode", [oc.BIRKS, oc.BOX, oc.ELLIPSOID])
def test_quench_and_drift_stages_match_oracle(torch_dev, mode):
    """quench / drift / shift as separate calls on (N, 26) tracks, column for column against the oracle."""
    import torch
    import larndsim_b200 as lb
    from larndsim_b200 import drifting, quenching, stream_ops
    kw = dict(number_pix_neighbors=2, signal_length=100)
    op = cm.oracle_params(**kw).replace(recombination_mode=mode, shift_x=0.11, shift_z=-0.07)
    pp = cm.product_params(**kw).replace(recombination_mode=lb.RecombinationMode(mode), shift_x=0.11, shift_z=-0.07)
    tr = cm.small_batch(900, ibatch=2, pad=12, precision=0.01)
    t = torch.as_tensor(tr, device=torch_dev)
    f32 = np.float32
    shifted = lo.shift_tracks(op, tr, cm.FIELDS, f32)
    quenched = lo.quench(op, shifted, cm.FIELDS, f32)
    drifted = lo.drift(op, quenched, cm.FIELDS, f32, oc.get_vdrift(op))
    got_s = stream_ops.tracks_stage(pp, t, cm.FIELDS, 1)
    assert np.array_equal(got_s.cpu().numpy(), shifted)
    got_q = quenching.quench(pp, got_s, cm.FIELDS).cpu().numpy()
    assert np.allclose(got_q, quenched, rtol=2e-6, atol=0, equal_nan=True)
    got_d = drifting.drift(pp, torch.as_tensor(quenched, device=torch_dev), cm.FIELDS).cpu().numpy()
    c = cm.FIELDS.index
    assert np.array_equal(got_d[:, c("pixel_plane")], drifted[:, c("pixel_plane")])
    assert np.allclose(got_d, drifted, rtol=3e-6, atol=1e-7, equal_nan=True)
    # all three at once == what the fused prepare kernel starts from
    all3 = stream_ops.tracks_stage(pp, t, cm.FIELDS, 7).cpu().numpy()
    assert np.allclose(all3, drifted, rtol=3e-6, atol=1e-7, equal_nan=True)
    # untouched columns are copied bit for bit
    same = [i for i, n in enumerate(cm.FIELDS) if n not in ("n_electrons", "long_diff", "tran_diff", "pixel_plane", "t", "t_start", "t_end")]
    assert np.array_equal(all3[:, same], shifted[:, same])


def _streams(d, dev):
    import torch
    T = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
    return dict(pixels=T(d["pixels"], torch.int32), t0_after_diff=T(d["t0_after_diff"], torch.float32),
                nelectrons=T(d["nelectrons"], torch.float32), long_diff=T(d["long_diff"], torch.float32),
                currents_idx=T(d["currents_idx"], torch.int32), nelectrons_neigh=T(d["nelectrons_neigh"], torch.float32),
                t0_neigh=T(d["t0_neigh"], torch.float32), currents_idx_neigh=T(d["currents_idx_neigh"], torch.int32))


@pytest.mark.parametrize("cfg", [dict(n=2, L=150, nseg=500), dict(n=4, L=100, nseg=300), dict(n=0, L=100, nseg=300)])
def test_simulate_signals_stream_form_matches_oracle_and_fused_kernels(torch_dev, cfg):
    """simulate_signals with the reference's argument list (materialised streams) == oracle == fused simulate_wfs."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=cfg["n"], signal_length=cfg["L"])
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 45 if cfg["n"] == 4 else 25, 45 if cfg["n"] == 4 else 25, 1950)
    tr = cm.small_batch(cfg["nseg"], ibatch=1, pad=7, precision=0.01)
    d = lo.simulate_drift_new(op, tr, cm.FIELDS)
    uniq, ren = lo.unique_and_renumber(d, history={})
    ref = lo.simulate_signals(op, uniq, d, ren, bank)
    s = _streams(d, torch_dev)
    bank_d = torch.as_tensor(bank, device=torch_dev)
    up = torch.as_tensor(uniq, device=torch_dev)
    w = sim.simulate_signals(pp, up, s["pixels"], s["t0_after_diff"], bank_d, s["nelectrons"], s["long_diff"], s["currents_idx"],
                             s["nelectrons_neigh"], torch.as_tensor(ren, device=torch_dev), s["t0_neigh"], s["currents_idx_neigh"])
    w = w.cpu().numpy()
    assert w.shape == ref.shape
    valid = uniq >= 0
    scale = np.abs(ref[valid][:, 1:]).max(axis=1, keepdims=True)
    assert (np.abs(w[valid][:, 1:] - ref[valid][:, 1:]) <= WFS_RTOL * scale + 1e-3).all()
    g_sc = np.maximum(np.abs(ref[:, 0]), np.abs(ref).max(axis=1))
    assert (np.abs(w[:, 0] - ref[:, 0]) <= 1e-3 * g_sc + 1e-3).all()          # garbage column
    bad_sc = np.abs(ref[~valid]).max(axis=1, keepdims=True)
    # rows of id < 0 (discarded by parse_output): every neighbour that is not a main pixel lands in row 0, a cancellation-prone
    # sum of ~N*P^2*L float32 atomics in arbitrary order (the fused kernels pre-reduce it per CTA and hold 1e-4)
    assert (np.abs(w[~valid] - ref[~valid]) <= 1e-3 * bad_sc + 1e-3).all()
    # the device-built streams (simulate_drift_new view) through the stream kernel == the fused path
    trd = torch.as_tensor(tr, device=torch_dev)
    mp, pixels, nel, t0, ld, ci, pn, cin, qn, t0n = sim.simulate_drift_new(pp, trd, cm.FIELDS, response_template=bank_d)
    wf, upf = sim.simulate_wfs(pp, bank_d, trd, cm.FIELDS, npix_capacity=len(uniq))
    assert np.array_equal(upf.cpu().numpy(), uniq)
    ren_d = torch.searchsorted(upf.to(torch.int64), pn.reshape(-1).to(torch.int64))
    ren_d = torch.where((ren_d < len(uniq)) & (upf[ren_d.clamp(max=len(uniq) - 1)] == pn.reshape(-1)), ren_d, torch.zeros_like(ren_d))
    w2 = sim.simulate_signals(pp, upf, pixels, t0, bank_d, nel, ld, ci, qn, ren_d, t0n, cin)[:, 1:].cpu().numpy()
    wf = wf.cpu().numpy()
    sc = np.abs(wf[valid]).max(axis=1, keepdims=True)
    assert (np.abs(w2[valid] - wf[valid]) <= 2 * WFS_RTOL * sc + 1e-3).all()


def test_simulate_signals_stream_gradients_match_finite_differences(torch_dev):
    """VJP of the stream kernel w.r.t. nelectrons, t0_after_diff, long_diff, nelectrons_neigh, t0_neigh against float64
    central differences of the oracle along random directions (inside one tick: the waveform is piecewise linear in t0)."""
    import torch
    from larndsim_b200 import sim
    kw = dict(number_pix_neighbors=2, signal_length=100)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(150, ibatch=1, pad=3, precision=0.01)
    d = lo.simulate_drift_new(op, tr, cm.FIELDS)
    uniq, ren = lo.unique_and_renumber(d, history={})
    rng = np.random.default_rng(3)
    G = (rng.uniform(0.5, 1.5, (len(uniq), 1)) * (1 + 0.5 * np.sin(np.arange(2001)[None, :] / 23.0))).astype(np.float32)
    s = _streams(d, torch_dev)
    names = ("t0_after_diff", "nelectrons", "long_diff", "nelectrons_neigh", "t0_neigh")
    for n in names:
        s[n].requires_grad_(True)
    bank_d = torch.as_tensor(bank, device=torch_dev)
    w = sim.simulate_signals(pp, torch.as_tensor(uniq, device=torch_dev), s["pixels"], s["t0_after_diff"], bank_d, s["nelectrons"],
                             s["long_diff"], s["currents_idx"], s["nelectrons_neigh"], torch.as_tensor(ren, device=torch_dev),
                             s["t0_neigh"], s["currents_idx_neigh"])
    (w * torch.as_tensor(G, device=torch_dev)).sum().backward()
    grads = {n: s[n].grad.cpu().numpy().astype(np.float64) for n in names}
    bank64, cum64 = bank.astype(np.float64), np.cumsum(bank.astype(np.float64), axis=-1)
    d64 = {k: (np.asarray(v, dtype=np.float64) if np.asarray(v).dtype.kind == "f" else v) for k, v in d.items()}

    def L(dd):
        return float((lo.simulate_signals(op, uniq, dd, ren, bank64, cum64, dt=np.float64) * G.astype(np.float64)).sum())

    steps = dict(t0_after_diff=1e-5, t0_neigh=1e-5, nelectrons=1e-2, nelectrons_neigh=1e-2, long_diff=1e-5)
    for n in names:
        ts = float(np.float32(op.t_sampling))
        base = d64[n]
        direction = rng.uniform(0.5, 1.0, base.shape) * rng.choice([-1.0, 1.0], base.shape)
        if n.startswith("t0"):   # keep every entry inside its tick for +-h
            fr = base / ts - np.floor(base / ts)
            direction = np.where((fr > 0.02) & (fr < 0.98), direction, 0.0)
        if n == "long_diff":     # ... and inside its template interval (the node index is discrete)
            tv = np.asarray(op.long_diff_template, dtype=np.float64)
            gap = np.minimum(np.abs(base[:, None] - tv[None, :]).min(axis=1), 1.0)
            direction = np.where(gap > 1e-3, direction, 0.0)
        h = steps[n]
        dp, dm = dict(d64), dict(d64)
        dp[n], dm[n] = base + h * direction, base - h * direction
        fd = (L(dp) - L(dm)) / (2 * h)
        an = float((grads[n] * direction).sum())
        assert abs(an - fd) <= GRAD_RTOL * abs(fd) + 1e-7 * np.abs(grads[n]).sum(), (n, an, fd)


def test_mc_stage_operators_match_oracle(torch_dev):
    """simulate_drift -> current_mc -> accumulate_signals_parametrized called one by one == oracle == fused MC kernels."""
    import torch
    from larndsim_b200 import detsim, jrandom, sim
    from oracle import jax_random as jr
    kw = dict(number_pix_neighbors=0, signal_length=150, mc_diff=True)
    for diff_in_current in (True, False):
        op = cm.oracle_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
        pp = cm.product_params(**kw).replace(diffusion_in_current_sim=diff_in_current)
        tr = cm.small_batch(400, ibatch=1, pad=5, precision=0.01)
        k1 = jrandom.split(jrandom.key(11), 2)[0]
        rnd = jr.normal(jr.split(jr.key(11), 2)[0], (len(tr), 3)).astype(np.float32)
        el_o, pid_o = lo.simulate_drift_mc(op, tr, cm.FIELDS, rnd)
        trd = torch.as_tensor(tr, device=torch_dev)
        el, pids = sim.simulate_drift(pp, trd, cm.FIELDS, k1)
        assert np.allclose(el.cpu().numpy(), el_o, rtol=3e-6, atol=2e-6, equal_nan=True)
        # from here on use the oracle's electrons so that integer outputs can be compared bit for bit
        eld = torch.as_tensor(el_o, device=torch_dev)
        assert np.array_equal(detsim.get_pixels(pp, eld, cm.FIELDS).reshape(-1).cpu().numpy(), pid_o.ravel())
        pid_o = pid_o.ravel()
        px, py, plane, _ = lo.id2pixel(op, pid_o)
        coords = lo.get_pixel_coordinates(op, px, py, plane)
        tick_o, sig_o = lo.current_mc(op, el_o, coords, cm.FIELDS)
        tick, sig = detsim.current_mc(pp, eld, torch.as_tensor(coords, device=torch_dev), cm.FIELDS)
        assert np.array_equal(tick.cpu().numpy(), tick_o)
        sc = np.abs(sig_o).max(axis=1, keepdims=True)
        assert (np.abs(sig.cpu().numpy() - sig_o) <= 3e-5 * sc + 1e-3).all()
        uniq = np.unique(pid_o)
        uniq = np.sort(np.pad(uniq, (0, 5), constant_values=-1)).astype(np.int32)
        ren = np.searchsorted(uniq, pid_o)
        w_o = lo.accumulate_signals_parametrized(np.zeros((len(uniq), 2001), np.float32), sig_o, ren, tick_o.astype(np.int64) - 51)
        w = detsim.accumulate_signals_parametrized(torch.zeros((len(uniq), 2001), device=torch_dev), torch.as_tensor(sig_o, device=torch_dev),
                                                   torch.as_tensor(ren, device=torch_dev), torch.as_tensor(tick_o - 51, device=torch_dev))
        sc = np.abs(w_o).max(axis=1, keepdims=True)
        assert (np.abs(w.cpu().numpy() - w_o) <= WFS_RTOL * sc + 1e-3).all()
        # the whole stage-by-stage chain == the fused simulate_parametrized
        up = torch.as_tensor(uniq, device=torch_dev)
        adcs, x, y, z, ticks, hp, ev = sim.simulate_signals_parametrized(pp, eld, torch.as_tensor(pid_o, device=torch_dev), up, None, cm.FIELDS)
        out = sim.parse_output(pp, adcs, x, y, z, ticks, hp, ev, up)
        fused = sim.simulate_parametrized(pp, trd, cm.FIELDS, rnd=torch.as_tensor(rnd, device=torch_dev), npix_capacity=len(uniq))
        assert out[8] == len(fused[0])
        assert np.array_equal(out[4].cpu().numpy(), fused[4].cpu().numpy()) and np.array_equal(out[7].cpu().numpy(), fused[7].cpu().numpy())
        assert np.abs(out[0].cpu().numpy() - fused[0].cpu().numpy()).max() <= 2e-3


def test_current_mc_and_accumulate_gradients(torch_dev):
    """VJPs of current_mc (x, y, z, long_diff, n_electrons columns, pixel centres) and of accumulate_signals_parametrized
    against float64 central differences of the oracle."""
    import torch
    from larndsim_b200 import detsim
    kw = dict(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=True)
    op, pp = cm.oracle_params(**kw), cm.product_params(**kw)
    tr = cm.small_batch(200, ibatch=1, pad=0, precision=0.01)
    rng = np.random.default_rng(5)
    rnd = rng.normal(size=(len(tr), 3)).astype(np.float32)
    el_o, pid_o = lo.simulate_drift_mc(op, tr, cm.FIELDS, rnd)
    pid_o = pid_o.ravel()
    px, py, plane, _ = lo.id2pixel(op, pid_o)
    coords = lo.get_pixel_coordinates(op, px, py, plane)
    G = rng.uniform(0.5, 1.5, (len(tr), 51))
    eld = torch.as_tensor(el_o, device=torch_dev).requires_grad_(True)
    pcd = torch.as_tensor(coords, device=torch_dev).requires_grad_(True)
    tick, sig = detsim.current_mc(pp, eld, pcd, cm.FIELDS)
    (sig * torch.as_tensor(G, device=torch_dev, dtype=torch.float32)).sum().backward()
    g_el, g_pc = eld.grad.cpu().numpy().astype(np.float64), pcd.grad.cpu().numpy().astype(np.float64)
    tick_o, _ = lo.current_mc(op, el_o, coords, cm.FIELDS)

    def L(e, c):
        t, s = lo.current_mc(op, e, c, cm.FIELDS, dt=np.float64)
        assert np.array_equal(t, tick_o)
        return float((s * G).sum())

    e64, c64 = el_o.astype(np.float64), coords.astype(np.float64)
    c = cm.FIELDS.index
    for name, h in dict(x=1e-6, y=1e-6, z=1e-6, long_diff=1e-7, n_electrons=1e-2).items():
        direction = rng.uniform(0.5, 1.0, len(tr)) * rng.choice([-1.0, 1.0], len(tr))
        ep, em = e64.copy(), e64.copy()
        ep[:, c(name)] += h * direction
        em[:, c(name)] -= h * direction
        fd = (L(ep, c64) - L(em, c64)) / (2 * h)
        an = float((g_el[:, c(name)] * direction).sum())
        assert abs(an - fd) <= 5e-3 * abs(fd) + 1e-6 * np.abs(g_el[:, c(name)]).sum(), (name, an, fd)
    assert np.allclose(g_pc[:, 0], -g_el[:, c("x")]) and np.allclose(g_pc[:, 1], -g_el[:, c("y")])
    others = [i for i, n in enumerate(cm.FIELDS) if n not in ("x", "y", "z", "long_diff", "n_electrons")]
    assert not g_el[:, others].any()
    # accumulate: d out / d signals is a gather of the upstream gradient, d out / d wfs the identity
    uniq = np.sort(np.unique(pid_o)).astype(np.int32)
    ren = torch.as_tensor(np.searchsorted(uniq, pid_o), device=torch_dev)
    sigd = torch.as_tensor(sig.detach(), device=torch_dev).requires_grad_(True)
    w0 = torch.zeros((len(uniq), 2001), device=torch_dev, requires_grad=True)
    Gw = torch.as_tensor(rng.uniform(0.5, 1.5, (len(uniq), 2001)).astype(np.float32), device=torch_dev)
    start = tick.to(torch.int32) - 51
    (detsim.accumulate_signals_parametrized(w0, sigd, ren, start) * Gw).sum().backward()
    tt = start[:, None] + torch.arange(51, device=torch_dev)[None, :]
    tt = torch.where((tt < 0) | (tt >= 2000), torch.zeros_like(tt), tt + 1)
    assert torch.equal(sigd.grad, Gw[ren[:, None], tt])
    assert torch.equal(w0.grad, Gw)
