#!/usr/bin/env python
"""Benchmark of the B200-native larnd-sim hot path (LUT mode), see DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--segments S]

One *step* = one pass of the hot path over one synthetic spill batch of the prepared_data shape
(SURVEY.md §8d config 5): prepare -> unique/renumber -> LUT accumulate -> fused FEE/ADC (+ hit compaction).
`value`      : forward segments/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`        : the same through the reference-facing entry (dataio.simulate_from_raw's path): the RAW, un-chopped rows
               of the batch are copied from pinned host memory, chopped on the device (inside the prepare kernel),
               simulated, and the hit list is read back — all inside the timed region.  `e2e_chopped` / `e2e_packed`: a caller that holds CHOPPED
               batches on the host uploads all 26 columns (104 B/segment) / the 10 columns the simulation reads (40 B).
`fwd_grad`   : forward + loss + backward (FEE VJP -> accumulate VJP -> 15 parameter gradients, all-reduced over ranks).
`mc_mode`    : BASELINE config 3 (MC-current mode) forward and forward+grad.
`fit_step`   : BASELINE config 4: one Adam step of the reference's --lut fit (n = 2, L = 150, ~19.8 k segments per rank,
               mse_adc, six fitted parameters), events sharded over the ranks, steps/s.
`prepared_inputs`: BASELINE config 2: all 22 prepared inputs through the production-driver loop (N = 1 only).
`scan_2d`    : BASELINE config 5b: 16 x 16 likelihood scan (loss + 2 gradients per point), grid points x event shards
               distributed over the ranks, wall time.
`roofline`   : dominant kernels (forward and backward tile kernels): algorithmic HBM bytes / CUDA-event kernel time vs the
               measured peak, DRAM traffic from the committed ncu capture.
`cpu_baseline`: the numpy oracle (a port of the reference algorithm; JAX is absent from this image AND from the GPU box,
               profiles/r2_probe_jax.txt) on host cores.
--impl reference times that oracle port on all host cores on a bounded sample of the SAME workload and config.
Multi-GPU: events are independent, so every rank simulates its own batch (weak scaling); the only collectives are the
16-float all-reduce of (loss, gradients) in fwd_grad, two small all-reduces per fit step and one all-gather per scan.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))

import numpy as np  # noqa: E402

METRIC = "segments/s fwd (LUT mode)"
NEIGH, SIGLEN, PRECISION = 4, 100, 0.01
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")


MC_SEGMENTS = 2_000_000  # size of the MC-current-mode side measurement
FIT_SEGMENTS = 19_800    # optimize/fit_test.sh --lut: 200 cm of track at 0.01 cm per batch
FIT_NAMES = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
FIT_NOMINAL = dict(Ab=0.8, kb=0.0486, eField=0.5, lifetime=2.2e3, long_diff=4.0e-6, tran_diff=8.8e-6)
FIT_TARGET = dict(Ab=0.83, kb=0.055, eField=0.52, lifetime=1.8e3, long_diff=5.0e-6, tran_diff=10e-6)
SCAN_GRID = 16
SCAN_RANGES = {"eField": (0.45, 0.55), "lifetime": (500.0, 5000.0)}  # optimize/ranges.py down/up


def workload_config(nseg, n_events):
    """`config` of the JSON line: a function of the generator arguments only, so that both arms print the same dict."""
    return {"workload": "synthetic_spill: %d segments/GPU (straight tracks chopped at %.2f cm, %d events), LUT mode, "
                        "synthetic (45,45,1950) response x 100 templates" % (nseg, PRECISION, n_events),
            "number_pix_neighbors": NEIGH, "signal_length": SIGLEN,
            "l2": "inputs (%.0f MB tracks + waveforms > 1 GB) exceed the 126 MB L2" % (nseg * 104 / 1e6),
            "garbage_row": "computed (reference-identical)"}


def chopped_count(raw, fields):
    """Number of rows chop_tracks(raw, fields, PRECISION) produces (float32 length, as numpy evaluates the reference)."""
    c = lambda n: fields.index(n)
    seg = np.stack([raw[:, c(a + "_end")] - raw[:, c(a + "_start")] for a in "xyz"], axis=1).astype(np.float32)
    length = np.sqrt(np.sum(seg ** 2, axis=1))
    return int(np.maximum(np.ceil(length / PRECISION), 1).astype(np.int64).sum())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=10_000_000, help="segments per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=12_000, help="segments of the workload timed on the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-garbage", action="store_true", help="also report the variant that drops the garbage row")
    ap.add_argument("--no-extras", action="store_true", help="skip the fit-step / scan / MC-mode side measurements")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in (self.rows[self.first:] or self.rows[-3:]):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_params(n_neigh=NEIGH, siglen=SIGLEN):
    from oracle import consts as oc
    return oc.params_from_geometry_json(GEOM).replace(number_pix_neighbors=n_neigh, signal_length=siglen,
                                                      electron_sampling_resolution=PRECISION, RESET_NOISE_CHARGE=0,
                                                      UNCORRELATED_NOISE_CHARGE=0, time_window=siglen)


_G = {}  # large read-only arrays shared with forked workers (not pickled per job)


def _oracle_worker(job):
    """One shard of the CPU baseline: full oracle forward (simulate_wfs + simulate_stochastic) on a slice of segments."""
    tracks, fields = job
    bank, cum = _G["bank"], _G["cum"]
    from oracle import larnd_oracle as lo
    p = oracle_params()
    ev = tracks[:, fields.index("eventID")]
    tracks = tracks.copy()
    tracks[:, fields.index("eventID")] = ev - ev.min()
    wfs, uniq = lo.simulate_wfs(p, bank, tracks, fields, history={}, response_cum=cum)
    out = lo.simulate_stochastic(p, wfs, uniq)
    return len(out[0])


def cpu_oracle_rate(tracks, bank32, fields, nproc, per_proc):
    """segments/s of the oracle port over `nproc` processes each handling `per_proc` segments."""
    from oracle import larnd_oracle as lo
    if _G.get("bank") is not bank32:
        _G["bank"], _G["cum"] = bank32, lo.response_cumsum(bank32)
    jobs = [(tracks[i * per_proc:(i + 1) * per_proc], fields) for i in range(nproc)]
    jobs = [j for j in jobs if len(j[0])]
    nseg = sum(len(j[0]) for j in jobs)
    t = time.time()
    if len(jobs) == 1:
        _oracle_worker(jobs[0])
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            pool.map(_oracle_worker, jobs)
    return nseg / (time.time() - t), nseg


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores.  jax is installed neither in this image nor on
    the GPU box (profiles/r2_probe_jax.txt), so this is the numpy oracle port (kind "port"), on all host cores, each step a
    bounded sample of OUR arm's workload (same generator, same seed, same chopping, same physics configuration)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from larndsim_b200 import synthetic
    from oracle import consts as oc
    from oracle import larnd_oracle as lo
    cores = os.cpu_count() or 1
    per_proc = 2500
    fields = synthetic.FIELDS
    raw, n_events = synthetic.synthetic_raw_tracks(args.segments, seed=1234, precision=PRECISION)
    nseg_full = chopped_count(raw, fields)
    # the first raw tracks of the workload, chopped like the reference does, until the sample holds per_proc * cores segments
    need, rows, k = per_proc * cores, [], 0
    while sum(len(r) for r in rows) < need and k < raw.shape[0]:
        rows.append(lo.chop_tracks(raw[k:k + 8], fields, PRECISION))
        k += 8
    tracks = np.concatenate(rows, axis=0)[:need]
    p = oracle_params()
    bank32 = oc.build_response_template(synthetic.synthetic_response(), p, n_templates=32)
    rates = []
    for i in range(args.warmup + args.steps):
        r, nseg = cpu_oracle_rate(tracks, bank32, fields, cores, per_proc)
        if i >= args.warmup:
            rates.append(r)
    val = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nseg / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(nseg_full, n_events),
        "cpu_baseline": {"value": val, "unit": "segments/s", "cores": cores, "kind": "port",
                         "sample": "the first %d segments of the workload per step (%d processes x %d segments; only the first 32 of "
                                   "the 100 templates are built: none of the sample's segments needs a later one), numpy oracle port of "
                                   "the reference algorithm (the reference needs jax: absent from the image and from the GPU box)"
                                   % (nseg, cores, per_proc)},
        "e2e": {"value": val, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import larndsim_b200 as lb
    from larndsim_b200 import _lib, dataio, fit, parallel, sim, synthetic
    from larndsim_b200.consts import build_response_template

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU: keep this rank's pinned host buffers and copy threads on the GPU's own NUMA node
        numa_cpus = parallel.bind_to_gpu_numa_node(local) if os.environ.get("LARND_NO_NUMA_BIND") is None else None
    else:
        numa_cpus = None
    lb.build_library()
    lib = lb.get_lib()

    fields = synthetic.FIELDS
    Params = lb.build_params_class([])
    params = lb.load_geometry_json(Params, GEOM).replace(number_pix_neighbors=NEIGH, signal_length=SIGLEN,
                                                          electron_sampling_resolution=PRECISION, RESET_NOISE_CHARGE=0,
                                                          UNCORRELATED_NOISE_CHARGE=0, time_window=SIGLEN)
    t_gen = time.time()
    # raw tracks (one row per track, the prepared_data shape) -> chopped on the DEVICE (csrc/chop.cu, bit-identical to the
    # reference's host-side chop_tracks); the chopped batch is the resident input of `value`, the raw rows the input of `e2e`
    raw_np, n_events = synthetic.synthetic_raw_tracks(args.segments, seed=1234 + rank, precision=PRECISION)
    raw_host = torch.from_numpy(raw_np).pin_memory()
    tracks = dataio.chop_tracks(raw_host.to(dev), fields, PRECISION)
    nseg = tracks.shape[0]
    assert rank != 0 or nseg == chopped_count(raw_np, fields)
    tracks_host = tracks.cpu().pin_memory()
    packed_dev, pfields = dataio.pack_columns(tracks, fields)
    packed_host = packed_dev.cpu().pin_memory()
    del packed_dev
    tracks_np = tracks_host.numpy()
    bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    # size the outputs once (the reference pads the pixel list per batch; the capacity mode needs no host sync per step)
    st0 = sim.lut_forward(params, bank, tracks, fields, n_events=n_events)
    n_unique = int(st0.counts[0].item())
    npix = st0.npix
    out = (st0.unique_pixels, st0.wfs_buf)
    pod = st0.pod
    del st0

    # hits-only pipeline (sim.simulate_hits' kernels with preallocated outputs): the waveform buffer is scratch, the front-end
    # kernel zeroes what it has read (LARND_FEE_CLEAR_WFS) and the next accumulate skips the 2 GB memset (LARND_FLAG_WFS_ZERO)
    arena_clean = {"ok": False}

    def fwd(flags=0, src=tracks, flds=fields, raw=None):
        was_clean, arena_clean["ok"] = arena_clean["ok"], False
        st = sim.lut_forward(params, bank, src, flds, npix_capacity=npix, n_events=n_events, flags=flags, out=out, raw=raw,
                             wfs_zero=was_clean)
        fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True, pod=pod, clear_wfs=True)
        arena_clean["ok"] = True
        return st, fs

    # target for the fit-style loss: the ADCs of a slightly different detector (lifetime -10 %)
    st, fs = fwd()
    p_tgt = params.replace(lifetime=params.lifetime * 0.9)
    st_t = sim.lut_forward(p_tgt, bank, tracks, fields, npix_capacity=npix, n_events=n_events)
    fs_t = sim.fee_forward(p_tgt, st_t.wfs_full[:, 1:], st_t.unique_pixels, None, compact=False, pod=pod)
    adc_target = fs_t.adc.clone()
    del st_t, fs_t

    steps_buf = torch.empty(lib.larnd_fee_steps_bytes(npix), dtype=torch.uint8, device=dev)

    def fwd_grad(dense=False):
        st, fs = fwd(flags=0)
        diff = (fs.adc - adc_target) * (st.unique_pixels >= 0).unsqueeze(1)  # only real pixels enter the loss (parse_output)
        loss = (diff * diff).sum()
        if dense:   # the two VJPs through the dense (npix, n_ticks) waveform gradient (what an arbitrary loss on the waveforms needs)
            grad = sim.lut_backward(st, sim.fee_backward(fs, 2.0 * diff))
        else:       # front-end VJP as a step list per pixel row -> accumulate VJP from running-sum differences (no dense gradient)
            grad = sim.hits_backward(st, fs, 2.0 * diff, steps=steps_buf)
        red = torch.cat([loss.reshape(1), grad])
        if world > 1:
            dist.all_reduce(red)
        return red

    hit_host = torch.empty((8, npix * pod.max_adc_values // 4 + 1024), dtype=torch.float32).pin_memory()
    nv_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)

    def read_back_hits(fs):
        """D2H of the compacted hit list on the copy stream (overlaps the next step's kernels)."""
        main = torch.cuda.current_stream()
        done = torch.cuda.Event()
        done.record(main)
        hf, hi = fs.hits
        cap = hit_host.shape[1]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            nv_host.copy_(fs.n_valid, non_blocking=True)
            hit_host[:6].copy_(hf[:, :cap], non_blocking=True)
            hit_host[6:8].copy_(hi[:, :cap].view(torch.float32), non_blocking=True)
            hf.record_stream(copy_stream); hi.record_stream(copy_stream); fs.n_valid.record_stream(copy_stream)

    # HEADLINE end-to-end step, the reference-facing entry for production batches (dataio.simulate_from_raw's path with
    # preallocated buffers): H2D of the RAW, un-chopped rows (what the input file holds), chop on the device, forward, D2H
    # of the hit list.  The reference chops on the host and ships 104 B per chopped segment.
    raw_dev = torch.empty(raw_host.shape, dtype=torch.float32, device=dev)

    def e2e_step():
        raw_dev.copy_(raw_host, non_blocking=True)
        # chop_tracks fused into the prepare kernel (larnd_lut_prepare_raw): the chopped (n, 26) batch is never written
        st, fs = fwd(src=raw_dev, raw=(PRECISION, nseg))
        read_back_hits(fs)
        return fs

    chop_buf = torch.empty_like(tracks)

    def e2e_step_unfused():   # the same entry with a separate chop kernel (round-2 path), reported as e2e_chop_first
        raw_dev.copy_(raw_host, non_blocking=True)
        dataio.chop_tracks(raw_dev, fields, PRECISION, out=chop_buf)
        st, fs = fwd(src=chop_buf)
        read_back_hits(fs)
        return fs

    # Callers that hold CHOPPED batches on the host: every step copies ITS batch host->device and reads its hits back.  The
    # copies run on a second stream and are double-buffered (H2D of step i+1 and D2H of step i-1 overlap the kernels of step
    # i); `packed` uploads only the ten columns the simulation reads (`fields` is an argument of the API: a 10-column batch
    # is a valid input as it is).
    def make_chopped_step(host, flds):
        bufs = [torch.empty(host.shape, dtype=torch.float32, device=dev) for _ in range(2)]
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        buf_free = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0, "primed": False}

        def issue_h2d(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(buf_free[slot])          # the kernels that last read this buffer are done
                bufs[slot].copy_(host, non_blocking=True)
                h2d_done[slot].record(copy_stream)

        def step():
            main = torch.cuda.current_stream()
            slot = state["i"] & 1
            if not state["primed"]:
                buf_free[0].record(main); buf_free[1].record(main)
                issue_h2d(slot)
                state["primed"] = True
            issue_h2d(slot ^ 1)                                   # prefetch the next step's batch while this one computes
            main.wait_event(h2d_done[slot])
            st, fs = fwd(src=bufs[slot], flds=flds)
            buf_free[slot].record(main)
            read_back_hits(fs)
            state["i"] += 1
            return fs
        return step, bufs

    def timed(fn, steps, warmup, sampler=None, join=None):
        if sampler:
            sampler.start()          # nvidia-smi needs ~0.3 s to start streaming: launch it before the warm-up
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark()           # only samples taken from here on (timed region, GPU under load) are summarised
        l0 = lib.larnd_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)   # the timed region ends after the last copy of the last step
        e1.record()
        torch.cuda.synchronize()
        launches = lib.larnd_launch_count() - l0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, int(launches)

    # the clock sampler covers the timed regions value .. fwd_grad: ~100 ms nvidia-smi period
    w3 = max(2, args.warmup // 3)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_fwd, launches = timed(fwd, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(e2e_step, args.steps, w3, join=copy_stream)
    ms_e2e_u, _ = timed(e2e_step_unfused, args.steps, w3, join=copy_stream)
    del chop_buf
    step_c, bufs_c = make_chopped_step(tracks_host, fields)
    ms_e2e_c, _ = timed(step_c, args.steps, w3, join=copy_stream)
    del step_c, bufs_c
    step_p, bufs_p = make_chopped_step(packed_host, pfields)
    ms_e2e_p, _ = timed(step_p, args.steps, w3, join=copy_stream)
    del step_p, bufs_p
    ms_fg, launches_fg = timed(fwd_grad, args.steps, max(1, args.warmup // 3))
    ms_fg_dense, _ = timed(lambda: fwd_grad(dense=True), max(2, args.steps // 2), 1)
    g_check = [fwd_grad(dense=d)[1:].double().cpu().numpy() for d in (False, True)]   # the two backward paths agree (reported below)
    clocks = sampler.stop() if sampler else None
    ms_skip = None
    if args.skip_garbage:
        ms_skip, _ = timed(lambda: fwd(flags=1), args.steps, 1)
        fwd()

    extras = {}
    if not args.no_extras:
        # ---- BASELINE config 3: MC-current mode (simulate_parametrized, mc_diff, diffusion in the current model, n = 0, L = 150
        # as in optimize/fit_test.sh:95-102) on the first MC_SEGMENTS segments of the same batch: prepare + unique + analytic
        # current + scatter + FEE, random numbers from the device Threefry stream (drawn once, outside the timed region)
        n_mc = min(nseg, MC_SEGMENTS)
        p_mc = params.replace(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=True)
        tr_mc = tracks[:n_mc].contiguous()
        rnd_mc = sim.mc_normals(n_mc, 0, dev)
        st_mc = sim.mc_forward(p_mc, tr_mc, fields, rnd_mc, n_events=n_events)
        npix_mc, pod_mc = st_mc.npix, st_mc.pod
        fs_mc = sim.fee_forward(p_mc, st_mc.wfs_full[:, 1:], st_mc.unique_pixels, None, compact=False, pod=pod_mc)
        adc_mc_target = (fs_mc.adc * 0.9).clone()
        del st_mc, fs_mc

        def mc_fwd():
            stm = sim.mc_forward(p_mc, tr_mc, fields, rnd_mc, npix_capacity=npix_mc, n_events=n_events)
            return stm, sim.fee_forward(p_mc, stm.wfs_full[:, 1:], stm.unique_pixels, None, compact=True, pod=pod_mc)

        def mc_fwd_grad():
            stm, fsm = mc_fwd()
            diff = (fsm.adc - adc_mc_target) * (stm.unique_pixels >= 0).unsqueeze(1)
            g_wfs = sim.fee_backward(fsm, 2.0 * diff)
            red = torch.cat([(diff * diff).sum().reshape(1), sim.mc_backward(stm, tr_mc, g_wfs)])
            if world > 1:
                dist.all_reduce(red)
            return red

        ms_mc, _ = timed(mc_fwd, args.steps, w3)
        ms_mc_fg, _ = timed(mc_fwd_grad, args.steps, 1)
        extras["mc_mode"] = {"metric": "segments/s (MC-current mode, n=0, L=150, mc_diff, diffusion in the current model; BASELINE config 3)",
                             "value": n_mc * world / (ms_mc * 1e-3), "unit": "segments/s", "ms_per_step": ms_mc,
                             "fwd_grad": n_mc * world / (ms_mc_fg * 1e-3), "fwd_grad_ms_per_step": ms_mc_fg, "segments_per_gpu": n_mc}
        del tr_mc, rnd_mc, adc_mc_target

        # ---- BASELINE config 4: one Adam step of the --lut fit (optimize/fit_test.sh: 0.01 cm, n = 2, L = 150, ~19.8 k segments,
        # mse_adc, six fitted parameters); every rank holds ITS events, the loss sums and the gradients are all-reduced
        base4 = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=PRECISION, RESET_NOISE_CHARGE=0,
                     UNCORRELATED_NOISE_CHARGE=0)
        fit_np, fit_events = synthetic.synthetic_tracks(FIT_SEGMENTS, seed=500 + rank, precision=PRECISION)
        fit_tracks = torch.as_tensor(fit_np, device=dev)
        p4 = lb.load_geometry_json(lb.build_params_class(list(FIT_NAMES)), GEOM).replace(**base4)
        bank4 = build_response_template(synthetic.synthetic_response(25, 25, 1950), p4, device=dev)
        prob = fit.FitProblem.from_target_params(FIT_NAMES, p4, FIT_TARGET, bank4, fit_tracks, fields, fit_events)
        def time_fit(adam, n_fit=50):
            for _ in range(10):
                adam.step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            l0 = lib.larnd_launch_count()
            for _ in range(n_fit):
                adam.step()
            torch.cuda.synchronize()
            wall = torch.tensor([time.perf_counter() - t0], device=dev)
            nl = (lib.larnd_launch_count() - l0) / n_fit
            if world > 1:
                dist.all_reduce(wall, op=dist.ReduceOp.MAX)
            return float(wall.item()) / n_fit, nl

        adam = fit.FusedAdamFit(prob, FIT_NOMINAL, lr=0.01)
        s_fused, nl_fused = time_fit(adam)
        s_auto, nl_auto = time_fit(fit.AdamFit(prob, FIT_NOMINAL, lr=0.01))
        extras["fit_step"] = {"metric": "fit steps/s (BASELINE config 4: fwd + mse_adc + grads of 6 parameters + Adam, events sharded)",
                              "value": 1.0 / s_fused, "unit": "steps/s", "ms_per_step": 1e3 * s_fused,
                              "segments_per_step": int(fit_tracks.shape[0]) * world, "segments_per_gpu": int(fit_tracks.shape[0]),
                              "our_kernel_launches_per_step": nl_fused,
                              "path": "fit.FusedAdamFit: one asynchronous chain of C-ABI calls (dense mse_adc operator, no hit compaction, "
                                      "no autograd graph), one host synchronisation per step",
                              "autograd_path": {"ms_per_step": 1e3 * s_auto, "steps_per_s": 1.0 / s_auto, "our_kernel_launches_per_step": nl_auto,
                                                "path": "fit.AdamFit: simulate_wfs + simulate_stochastic + losses.mse_adc under torch.autograd"},
                              "timing": "host wall clock around 50 steps (the step is host-bound at this size), max over ranks",
                              "collectives": "2 all-reduces (5 loss sums, 15 gradients)" if world > 1 else "none"}

        # ---- BASELINE config 5b: 16 x 16 likelihood scan over (eField, lifetime): loss + 2 gradients per point; the grid points
        # are dealt to point groups, the events of a point are sharded over the ranks of its group
        event_shards = 2 if world >= 4 else 1
        scan_names = ("eField", "lifetime")
        p5 = lb.load_geometry_json(lb.build_params_class(list(scan_names)), GEOM).replace(**base4)
        scan_np, scan_events = synthetic.synthetic_tracks(FIT_SEGMENTS * event_shards, seed=77, precision=PRECISION)

        def make_problem(es, n_es, group):
            loc, nev, _ = parallel.shard_tracks(scan_np, fields, es, n_es)
            return fit.FitProblem.from_target_params(scan_names, p5, {}, bank4, torch.as_tensor(loc, device=dev), fields, nev, group=group)

        a1, a2 = np.linspace(*SCAN_RANGES["eField"], SCAN_GRID), np.linspace(*SCAN_RANGES["lifetime"], SCAN_GRID)
        fit.scan_2d(make_problem, "eField", a1[:2], "lifetime", a2[:2], event_shards=event_shards)   # warm-up (sizing, caches)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        table = fit.scan_2d(make_problem, "eField", a1, "lifetime", a2, event_shards=event_shards)
        torch.cuda.synchronize()
        wall = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        imin = np.unravel_index(np.nanargmin(table[..., 0]), table.shape[:2])
        extras["scan_2d"] = {"metric": "wall seconds of a %dx%d likelihood scan (BASELINE config 5b)" % (SCAN_GRID, SCAN_GRID),
                             "value": float(wall.item()), "unit": "s", "points": SCAN_GRID * SCAN_GRID, "points_per_s": SCAN_GRID * SCAN_GRID / float(wall.item()),
                             "segments_per_point": int(scan_np.shape[0]), "layout": "%d point groups x %d event shards" % (world // event_shards, event_shards),
                             "argmin": [float(a1[imin[0]]), float(a2[imin[1]])], "nominal": [0.5, 2200.0],
                             "finite_points": int(np.isfinite(table[..., 0]).sum())}
        del prob, adam, fit_tracks, bank4

    if not args.no_extras:
        # ---- BASELINE config 2: forward over ALL 22 prepared inputs (prepared_data/input_*.h5, committed as tests/golden/
        # segments_input_*.npz), settings of optimize/simulate_test.sh (0.005 cm, n = 4, L = 100, --chop, 50 cm batches): the
        # production driver's loop — TracksDataset batches assembled / chopped / padded on the device, simulate_wfs +
        # simulate_stochastic, hits read back per batch.  ~90 batches of ~10 k segments: the small-batch regime.
        import glob
        from larndsim_b200 import dataio as _dio
        files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "segments_input_*.npz")))
        if files and world == 1:
            p2 = params.replace(electron_sampling_resolution=0.005)
            dsets = [_dio.TracksDataset(np.load(f)["segments"], nevents=None, max_nbatch=None, swap_xz=True, max_batch_len=50, chopped=True,
                                        pad=False, electron_sampling_resolution=0.005, device=dev) for f in files]

            cap2 = {"npix": 0}

            def run_all(hits_only):
                nh = 0
                for ds in dsets:
                    f2 = ds.get_track_fields()
                    for ib in range(len(ds)):
                        cap = sim.pad_size(ds.batch_nsteps[ib], "batch_size", 0.5)
                        tr = ds.device_batch(ib, capacity=cap)
                        nev2 = len(ds.get_batch_global_event_ids(ib))
                        if hits_only:   # the driver's path after its first batch: one call over the self-cleaning arena
                            nh += int(sim.simulate_hits(p2, bank, tr, f2, rngseed=ib, npix_capacity=cap2["npix"], n_events=nev2)[0].shape[0])
                        else:           # reference-shaped two-call path (pad_size buckets of the pixel list, waveforms returned)
                            w, u = sim.simulate_wfs(p2, bank, tr, f2, n_events=nev2)
                            cap2["npix"] = max(cap2["npix"], int(u.shape[0]))
                            nh += int(sim.simulate_stochastic(p2, w, u, rngseed=ib)[0].shape[0])
                return nh

            def wall_of(hits_only):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                nh = run_all(hits_only)
                torch.cuda.synchronize()
                return time.perf_counter() - t0, nh
            run_all(False)                       # warm-up + pixel capacity of the arena
            wall2_two, nhits2_two = wall_of(False)
            run_all(True)
            wall2, nhits2 = wall_of(True)
            assert nhits2 == nhits2_two, (nhits2, nhits2_two)
            nseg2 = sum(sum(ds.batch_nsteps) for ds in dsets)
            nb2 = sum(len(ds) for ds in dsets)
            extras["prepared_inputs"] = {"metric": "segments/s over all 22 prepared inputs (BASELINE config 2, production-driver loop)",
                                         "value": nseg2 / wall2, "unit": "segments/s", "wall_s": wall2, "files": len(files), "batches": nb2,
                                         "segments": int(nseg2), "hits": int(nhits2), "ms_per_batch": 1e3 * wall2 / nb2,
                                         "path": "sim.simulate_hits (hits-only, persistent self-cleaning arena, one host sync per batch)",
                                         "two_call_path": {"value": nseg2 / wall2_two, "ms_per_batch": 1e3 * wall2_two / nb2,
                                                           "path": "simulate_wfs + simulate_stochastic with the reference's pad_size buckets"},
                                         "timing": "host wall clock, one pass after a warm-up pass (the loop is launch/sync-bound at ~10 k segments per batch)"}

    # per-kernel device times of the dominant kernels, same launches as above
    lib.larnd_profile_enable(1)
    import ctypes as C
    buf = (C.c_float * 4)()
    kt = np.zeros((args.steps, 4))
    for i in range(args.steps):
        fwd_grad()
        lib.larnd_profile_read(buf)
        kt[i] = list(buf)
    lib.larnd_profile_enable(0)
    k_ms = kt.mean(axis=0)
    n_valid = int(fs.n_valid.item())
    total_seg = nseg * world

    if rank == 0:
        peak, peak_src = measured_peak()
        nticks, L = pod.n_ticks, SIGLEN
        # algorithmic (compulsory) HBM bytes per segment, SURVEY.md §8d: forward = record + one write of every touched waveform
        # row + the LUT window; backward = re-read of the segment records (31 words) + one read of every gradient row
        lut_window = 4 * L * (25 * 3 + (10 * NEIGH + 5) ** 2)
        b_seg = 104 + 4.0 * (n_unique + 1) * nticks / nseg + lut_window / nseg
        # (steps form: 168 bytes of step events per pixel row instead of one read of the n_ticks-float gradient row)
        b_seg_bwd = 4 * 31 + 168.0 * (n_unique + 1) / nseg + lut_window / nseg
        c_seg = (25 + (2 * NEIGH + 1) ** 2) * (2 * L + 2)
        achieved = b_seg * nseg / (k_ms[1] * 1e-3) / 1e9
        achieved_bwd = b_seg_bwd * nseg / (k_ms[2] * 1e-3) / 1e9
        # DRAM bytes and L2 reduction sectors of the dominant kernels per launch come from ONE `ncu --set full` capture of this
        # workload (profiles/r2_traffic_10M.json, written by scripts/ncu_traffic.py), scaled to this run's segment count
        traffic = traffic_bwd = l2_red = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic_10M.json")) as fh:
                tj = json.load(fh)
            scale = nseg / float(tj.get("segments", 10010184))
            traffic = sum(float(e["dram_bytes"]) for e in tj["k_acc_tiles"]) * scale
            traffic_bwd = sum(float(e["dram_bytes"]) for e in tj["k_bwd_tiles"]) * scale
            sectors = sum(float(e.get("l2_red_sectors") or 0.0) for e in tj["k_acc_tiles"])
            if sectors:
                red_bytes = sectors * 32.0 * scale
                l2_red = {"achieved": red_bytes / (k_ms[1] * 1e-3) / 1e9, "peak": 5800.0, "unit": "GB/s",
                          "frac": red_bytes / (k_ms[1] * 1e-3) / 1e9 / 5800.0, "bytes_per_launch": red_bytes,
                          "peak_source": "scripts/ubench_red.cu, lane<->4 ticks red.global.add.v4.f32 into an L2-resident buffer "
                                         "(profiles/r1b_ubench_red.txt)", "sectors_source": "profiles/r2_traffic_10M.json"}
        except Exception:
            pass
        hits_bytes = int(hit_host.numel() * 4 + 4)
        seg_s = lambda ms: total_seg / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": seg_s(ms_fwd), "unit": "segments/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_fwd, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(nseg, n_events),
            "workload_stats": {"n_unique_pixels": n_unique, "npix_padded": npix, "hits": n_valid,
                               "waveform_buffer_mb": npix * sim.wfs_row_stride(nticks) * 4 / 1e6},
            "clocks": clocks,
            "e2e": {"value": seg_s(ms_e2e), "unit": "segments/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(raw_host.numel() * 4), "d2h_bytes_per_step": hits_bytes,
                    "entry": "dataio.simulate_from_raw path: raw (un-chopped) rows uploaded from pinned memory, chop_tracks fused into "
                             "the prepare kernel (larnd_lut_prepare_raw: the chopped batch is never written), simulate_wfs + "
                             "simulate_stochastic, hit list read back"},
            "e2e_chop_first": {"value": seg_s(ms_e2e_u), "unit": "segments/s", "ms_per_step": ms_e2e_u,
                               "note": "same entry with the separate chop kernel (csrc/chop.cu writes the chopped (n, 26) batch, "
                                       "larnd_lut_prepare reads it back): the round-2 path"},
            "e2e_chopped": {"value": seg_s(ms_e2e_c), "unit": "segments/s", "ms_per_step": ms_e2e_c,
                            "h2d_bytes_per_step": int(tracks_host.numel() * 4), "d2h_bytes_per_step": hits_bytes,
                            "note": "host-chopped batch, all 26 columns uploaded (104 B/segment), double-buffered"},
            "e2e_packed": {"value": seg_s(ms_e2e_p), "unit": "segments/s", "ms_per_step": ms_e2e_p,
                           "h2d_bytes_per_step": int(packed_host.numel() * 4), "d2h_bytes_per_step": hits_bytes,
                           "note": "host-chopped batch, the 10 columns the simulation reads (dataio.pack_columns, 40 B/segment)"},
            "fwd_grad": {"metric": "segments/s fwd+grad (LUT mode)", "value": seg_s(ms_fg), "unit": "segments/s",
                         "ms_per_step": ms_fg, "collective": "all_reduce(16 floats)" if world > 1 else "none",
                         "gpu_launches": launches_fg,
                         "backward": "sim.hits_backward: front-end VJP as <= 20 steps per pixel row, accumulate VJP from running-sum "
                                     "differences (no dense waveform gradient)",
                         "dense_gradient_path": {"value": seg_s(ms_fg_dense), "ms_per_step": ms_fg_dense,
                                                 "backward": "sim.fee_backward + sim.lut_backward through the dense (npix, n_ticks) gradient"},
                         "max_rel_diff_between_paths": float(np.max(np.abs(g_check[0] - g_check[1])[np.abs(g_check[1]) > 0] /
                                                                    np.abs(g_check[1])[np.abs(g_check[1]) > 0]))},
            "gpu_launches": launches,  # kernels of liblarnd_b200.so launched inside the timed region of `value` (larnd_launch_count)
            "kernels_ms": {"k_prepare": k_ms[0], "k_lut_accumulate": k_ms[1], "k_lut_backward": k_ms[2], "k_fee_forward": k_ms[3]},
            "roofline": {"kernel": "k_acc_tiles (class-sorted lut_accumulate: run sort + the 4- and 6-position tile kernels + row-0 reduction)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_segment": b_seg,
                         "note": "accumulate is bound by instruction issue and L1/L2 traffic of its FMA/reduction stream, not by HBM "
                                 "(SURVEY §8d): see l2_red, contributions/s and DESIGN.md §4",
                         "l2_red": l2_red,
                         "contributions_per_s": c_seg * nseg / (k_ms[1] * 1e-3),
                         "backward": {"kernel": "k_bwd_tiles<steps> (class-sorted VJP of lut_accumulate from the front end's step events + chain rule)", "bound": "hbm",
                                      "achieved": achieved_bwd, "peak": peak, "unit": "GB/s", "frac": achieved_bwd / peak,
                                      "traffic": traffic_bwd, "algorithmic_bytes_per_segment": b_seg_bwd},
                         # the two streaming kernels of the step, which ARE HBM-bound (algorithmic bytes / measured time):
                         # prepare reads the 104-byte record and writes the 31-word segment record; the front end reads
                         # every waveform row once
                         "hbm_kernels": {
                             "k_prepare": {"achieved": nseg * (104 + 4 * 31) / (k_ms[0] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": nseg * (104 + 4 * 31) / (k_ms[0] * 1e-3) / 1e9 / peak},
                             "k_fee_forward": {"achieved": npix * (nticks - 1) * 4.0 / (k_ms[3] * 1e-3) / 1e9, "peak": peak,
                                               "unit": "GB/s", "frac": npix * (nticks - 1) * 4.0 / (k_ms[3] * 1e-3) / 1e9 / peak}}},
            "setup_s": t_gen,
            "numa_bound_cpus": (len(numa_cpus) if numa_cpus else None),
        }
        line.update(extras)
        if ms_skip is not None:
            line["value_skip_garbage_row"] = seg_s(ms_skip)
        if world == 1 and not args.no_cpu_baseline:
            sample = tracks_np[: args.cpu_sample]
            bank32 = bank[:32].cpu().numpy()
            rate, ns = cpu_oracle_rate(sample, bank32, fields, 1, len(sample))
            line["cpu_baseline"] = {"value": rate, "unit": "segments/s", "cores": 1, "kind": "port",
                                    "sample": "first %d segments of the same workload, numpy oracle port, 1 process" % ns}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
