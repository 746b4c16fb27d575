#!/usr/bin/env python
"""Benchmark of the B200-native larnd-sim hot path (LUT mode), see DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--segments S]

One *step* = one pass of the hot path over one synthetic spill batch of the prepared_data shape
(SURVEY.md §8d config 5): prepare -> unique/renumber -> LUT accumulate -> fused FEE/ADC (+ hit compaction).
`value`     : forward segments/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`       : the same through the public API with the batch in pinned HOST memory: H2D copy of the tracks and
              D2H read-back of the hit list inside the timed region.
`fwd_grad`  : forward + loss + backward (FEE VJP -> accumulate VJP -> 15 parameter gradients, all-reduced over ranks).
`roofline`  : dominant kernel (k_lut_accumulate), algorithmic HBM bytes / CUDA-event kernel time vs the measured peak.
`cpu_baseline`: the numpy oracle (a port of the reference algorithm; JAX is not installable here) on host cores.
--impl reference times that oracle port on all host cores (the reference itself needs jax, absent from the image).
Multi-GPU: events are independent, so every rank simulates its own batch (weak scaling); the only collective is the
16-float all-reduce of (loss, gradients) in the fwd+grad step.
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from larndsim_b200 import synthetic
    from oracle import consts as oc
    cores = os.cpu_count() or 1
    per_proc = 2500
    tracks, _ = synthetic.synthetic_tracks(per_proc * cores, seed=1234, precision=PRECISION)
    p = oracle_params()
    bank32 = oc.build_response_template(synthetic.synthetic_response(), p, n_templates=32)
    fields = synthetic.FIELDS
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
        "config": {"workload": "synthetic_spill (straight tracks chopped at 0.01 cm), LUT mode n=4 L=100, bounded sample of %d segments/step" % nseg,
                   "number_pix_neighbors": NEIGH, "signal_length": SIGLEN},
        "cpu_baseline": {"value": val, "unit": "segments/s", "cores": cores, "kind": "port",
                         "sample": "%d segments/step, %d processes x %d segments, numpy oracle port of the reference algorithm "
                                   "(the reference needs jax, not installable in this image)" % (nseg, cores, per_proc)},
        "e2e": {"value": val, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import larndsim_b200 as lb
    from larndsim_b200 import _lib, sim, synthetic

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
        from larndsim_b200 import parallel
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
    # reference's host-side chop_tracks); the chopped batch is the input of `value` and of `e2e`, the raw rows of `e2e_raw`
    from larndsim_b200 import dataio
    raw_np, n_events = synthetic.synthetic_raw_tracks(args.segments, seed=1234 + rank, precision=PRECISION)
    raw_host = torch.from_numpy(raw_np).pin_memory()
    tracks = dataio.chop_tracks(raw_host.to(dev), fields, PRECISION)
    nseg = tracks.shape[0]
    tracks_host = tracks.cpu().pin_memory()
    tracks_np = tracks_host.numpy()
    resp = synthetic.synthetic_response()
    from larndsim_b200.consts import build_response_template
    bank = build_response_template(resp, params, device=dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    # size the outputs once (the reference pads the pixel list per batch; the capacity mode needs no host sync per step)
    st0 = sim.lut_forward(params, bank, tracks, fields, n_events=n_events)
    n_unique = int(st0.counts[0].item())
    npix = st0.npix
    out = (st0.unique_pixels, st0.wfs_buf)
    pod = st0.pod
    del st0

    def fwd(flags=0, src=tracks):
        st = sim.lut_forward(params, bank, src, fields, npix_capacity=npix, n_events=n_events, flags=flags, out=out)
        fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True, pod=pod)
        return st, fs

    # target for the fit-style loss: the ADCs of a slightly different detector (lifetime -10 %)
    st, fs = fwd()
    p_tgt = params.replace(lifetime=params.lifetime * 0.9)
    st_t = sim.lut_forward(p_tgt, bank, tracks, fields, npix_capacity=npix, n_events=n_events)
    fs_t = sim.fee_forward(p_tgt, st_t.wfs_full[:, 1:], st_t.unique_pixels, None, compact=False, pod=pod)
    adc_target = fs_t.adc.clone()
    del st_t, fs_t

    def fwd_grad():
        st, fs = fwd(flags=0)
        diff = (fs.adc - adc_target) * (st.unique_pixels >= 0).unsqueeze(1)  # only real pixels enter the loss (parse_output)
        loss = (diff * diff).sum()
        g_wfs = sim.fee_backward(fs, 2.0 * diff)
        grad = sim.lut_backward(st, g_wfs)  # rows of ids < 0 carry zero gradient: detected on the device, their work is skipped
        red = torch.cat([loss.reshape(1), grad])
        if world > 1:
            dist.all_reduce(red)
        return red

    hit_host = torch.empty((8, npix * pod.max_adc_values // 4 + 1024), dtype=torch.float32).pin_memory()
    nv_host = torch.zeros(1, dtype=torch.int32).pin_memory()

    # End-to-end step: every step copies ITS batch host->device (pinned, 104 B/segment) and reads its hit list back.
    # The copies run on a second stream and are double-buffered, so the H2D of step i+1 and the D2H of step i-1 overlap
    # the kernels of step i (what a production loop over batches does); all copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty_like(tracks), torch.empty_like(tracks)]
    h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
    buf_free = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def issue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(buf_free[slot])          # the kernels that last read this buffer are done
            dev_bufs[slot].copy_(tracks_host, non_blocking=True)
            h2d_done[slot].record(copy_stream)

    def e2e_step():
        main = torch.cuda.current_stream()
        slot = state["i"] & 1
        if not state["primed"]:
            buf_free[0].record(main); buf_free[1].record(main)
            issue_h2d(slot)
            state["primed"] = True
        issue_h2d(slot ^ 1)                                   # prefetch the next step's batch while this one computes
        main.wait_event(h2d_done[slot])
        st, fs = fwd(src=dev_bufs[slot])
        buf_free[slot].record(main)
        done = torch.cuda.Event()
        done.record(main)
        hf, hi = fs.hits
        cap = hit_host.shape[1]
        with torch.cuda.stream(copy_stream):                  # D2H of the compacted hit list, overlapping the next step
            copy_stream.wait_event(done)
            nv_host.copy_(fs.n_valid, non_blocking=True)
            hit_host[:6].copy_(hf[:, :cap], non_blocking=True)
            hit_host[6:8].copy_(hi[:, :cap].view(torch.float32), non_blocking=True)
            hf.record_stream(copy_stream); hi.record_stream(copy_stream); fs.n_valid.record_stream(copy_stream)
        state["i"] += 1
        return fs

    # End-to-end from the RAW rows: H2D of the un-chopped tracks (a few hundred KB), chop on the device,
    # forward, D2H of the hit list — the reference chops on the host and ships 104 B per chopped segment.
    raw_dev = torch.empty(raw_host.shape, dtype=torch.float32, device=dev)
    chop_buf = torch.empty_like(tracks)

    def e2e_raw_step():
        main = torch.cuda.current_stream()
        raw_dev.copy_(raw_host, non_blocking=True)
        dataio.chop_tracks(raw_dev, fields, PRECISION, out=chop_buf)
        st, fs = fwd(src=chop_buf)
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
        return fs

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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)   # the timed region ends after the last copy of the last step
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, clocks

    # the clock sampler covers all four timed regions (value, e2e, e2e_raw, fwd_grad): ~100 ms nvidia-smi period
    sampler = ClockSampler(local) if rank == 0 else None
    ms_fwd, _ = timed(fwd, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(e2e_step, args.steps, max(2, args.warmup // 3), join=copy_stream)
    ms_e2e_raw, _ = timed(e2e_raw_step, args.steps, max(2, args.warmup // 3), join=copy_stream)
    ms_fg, _ = timed(fwd_grad, args.steps, max(1, args.warmup // 3))
    clocks = sampler.stop() if sampler else None
    ms_skip = None
    if args.skip_garbage:
        ms_skip, _ = timed(lambda: fwd(flags=1), args.steps, 1)
        fwd()

    # MC-current mode (BASELINE config 3: simulate_parametrized, mc_diff, diffusion in the current model, n = 0, L = 150 as in
    # optimize/fit_test.sh:95-102) on the first MC_SEGMENTS segments of the same batch: prepare + unique + analytic current +
    # scatter + FEE, random numbers from the device Threefry stream (drawn once, outside the timed region)
    n_mc = min(nseg, MC_SEGMENTS)
    p_mc = params.replace(number_pix_neighbors=0, signal_length=150, mc_diff=True, diffusion_in_current_sim=True)
    tr_mc = tracks[:n_mc].contiguous()
    rnd_mc = sim.mc_normals(n_mc, 0, dev)
    st_mc = sim.mc_forward(p_mc, tr_mc, fields, rnd_mc, n_events=n_events)
    npix_mc, pod_mc = st_mc.npix, st_mc.pod
    del st_mc

    def mc_fwd():
        stm = sim.mc_forward(p_mc, tr_mc, fields, rnd_mc, npix_capacity=npix_mc, n_events=n_events)
        return sim.fee_forward(p_mc, stm.wfs_full[:, 1:], stm.unique_pixels, None, compact=True, pod=pod_mc)

    ms_mc, _ = timed(mc_fwd, args.steps, max(1, args.warmup // 3))

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
        # algorithmic (compulsory) HBM bytes of the forward path per segment, SURVEY.md §8d:
        lut_window = 4 * L * (25 * 3 + (10 * NEIGH + 5) ** 2)
        b_seg = 104 + 4.0 * (n_unique + 1) * nticks / nseg + lut_window / nseg
        c_seg = (25 + (2 * NEIGH + 1) ** 2) * (2 * L + 2)
        achieved = b_seg * nseg / (k_ms[1] * 1e-3) / 1e9
        # DRAM bytes of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full`
        # capture at the 10 M-segment workload (profiles/r1_traffic_10M.json), scaled to this run's segment count
        # L2 reduction traffic (the kernel's output leaves as red.global.add.f32): lts__t_sectors_srcunit_tex_op_red x 32 B of
        # the same capture, against the reduction throughput a pure flush kernel reaches with the same access pattern and
        # an L2-resident target (scripts/ubench_red.cu pattern A, profiles/r1b_ubench_red.txt)
        traffic, l2_red = None, None
        for tname in ("r1d_traffic_10M.json", "r1b_traffic_10M.json", "r1_traffic_10M.json"):
            try:
                with open(os.path.join(ROOT, "profiles", tname)) as fh:
                    tj = json.load(fh)
                # one forward pass launches k_acc_tiles once per kernel variant (4-position and 6-position tiles): sum them
                ents = tj.get("k_acc_tiles") or tj["k_acc_tiles<4>"]
                traffic = sum(float(e["dram_bytes"]) for e in ents) * nseg / 10010184.0
                sectors = sum(float(e.get("l2_red_sectors") or 0.0) for e in ents)
                if sectors:
                    red_bytes = sectors * 32.0 * nseg / 10010184.0
                    red_peak = 5170.0
                    l2_red = {"achieved": red_bytes / (k_ms[1] * 1e-3) / 1e9, "peak": red_peak, "unit": "GB/s",
                              "frac": red_bytes / (k_ms[1] * 1e-3) / 1e9 / red_peak, "bytes_per_launch": red_bytes,
                              "peak_source": "scripts/ubench_red.cu, lane<->tick 128-byte reductions into an L2-resident buffer "
                                             "(2 GB DRAM-resident target: 2080 GB/s)",
                              "sectors_source": "profiles/" + tname}
                break
            except Exception:
                continue
        line = {
            "metric": METRIC, "value": total_seg / (ms_fwd * 1e-3), "unit": "segments/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_fwd, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic_spill: %d segments/GPU (straight tracks chopped at %.2f cm, %d events), LUT mode, "
                                   "synthetic (45,45,1950) response x 100 templates" % (nseg, PRECISION, n_events),
                       "number_pix_neighbors": NEIGH, "signal_length": SIGLEN, "n_unique_pixels": n_unique, "npix_padded": npix,
                       "hits": n_valid, "l2": "inputs (%.0f MB tracks + %.0f MB waveforms) exceed the 126 MB L2" %
                                              (nseg * 104 / 1e6, npix * nticks * 4 / 1e6),
                       "garbage_row": "computed (reference-identical)"},
            "clocks": clocks,
            "e2e": {"value": total_seg / (ms_e2e * 1e-3), "unit": "segments/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(tracks_host.numel() * 4), "d2h_bytes_per_step": int(hit_host.numel() * 4 + 4)},
            "e2e_raw": {"value": total_seg / (ms_e2e_raw * 1e-3), "unit": "segments/s", "ms_per_step": ms_e2e_raw,
                        "h2d_bytes_per_step": int(raw_host.numel() * 4), "d2h_bytes_per_step": int(hit_host.numel() * 4 + 4),
                        "note": "un-chopped tracks uploaded, chop_tracks on the device (csrc/chop.cu)"},
            "fwd_grad": {"metric": "segments/s fwd+grad (LUT mode)", "value": total_seg / (ms_fg * 1e-3), "unit": "segments/s",
                         "ms_per_step": ms_fg, "collective": "all_reduce(16 floats)" if world > 1 else "none"},
            "mc_mode": {"metric": "segments/s fwd (MC-current mode, n=0, mc_diff, diffusion in the current model)",
                        "value": n_mc * world / (ms_mc * 1e-3), "unit": "segments/s", "ms_per_step": ms_mc, "segments_per_gpu": n_mc},
            "gpu_launches": int(args.steps * 15),  # prepare 1 + unique/scan 4 + sorted accumulate 7 (run sort 3, tiles 2, row 0, boundary) + FEE/compaction 3
            "kernels_ms": {"k_prepare": k_ms[0], "k_lut_accumulate": k_ms[1], "k_lut_backward": k_ms[2], "k_fee_forward": k_ms[3]},
            "roofline": {"kernel": "k_acc_tiles (class-sorted lut_accumulate: run sort + the 4- and 6-position tile kernels + row-0 reduction)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_segment": b_seg,
                         "note": "accumulate is bound by instruction issue and L2 reduction throughput, not HBM (SURVEY §8d): "
                                 "see l2_red and contributions/s",
                         "l2_red": l2_red,
                         "contributions_per_s": c_seg * nseg / (k_ms[1] * 1e-3),
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
        if ms_skip is not None:
            line["value_skip_garbage_row"] = total_seg / (ms_skip * 1e-3)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import consts as oc
            sample = tracks_np[: args.cpu_sample]
            bank32 = bank[:32].cpu().numpy()
            _ = oc
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
