#!/usr/bin/env python3
"""Production driver: the B200 counterpart of the reference's ``python -m optimize.simulate`` (optimize/simulate.py), same
command-line options, same output layout.

    python -m larndsim_b200.simulate --input_file prepared_data/input_0.h5 --output_file out.h5 \
        --electron_sampling_resolution 0.005 --number_pix_neighbors 4 --signal_length 100 --mode lut --lut_file response.npy --chop

Per batch of ``dataio.TracksDataset`` (whole events, assembled / chopped / padded on the device): ``simulate_wfs`` +
``simulate_stochastic`` (``--mode lut``) or ``simulate_parametrized``; hits with ``adc_clean != 0`` of real events are split
per event under their GLOBAL event id and written as ``batch_<i>/event_<id>/{adc_clean, adc, Q, pixels, ticks, eventID,
pix_x, pix_y, pix_z}`` (optimize/simulate.py:127-157) — HDF5 through larndsim_b200.h5io (h5py is not required), or one
``.npz`` with ``--out_np`` (:168-180).  ``--jac`` (forward-mode Jacobians, :59-63,124) is not available: the kernels
implement reverse mode.  ``--lut_cache DIR`` keeps the convolved template bank on disk between runs (consts.load_lut).
"""
import argparse
import logging
import os
import sys

import numpy as np
import torch

from . import _lib, consts, dataio, detsim, fee, h5io, losses, sim

logger = logging.getLogger("larndsim_b200.simulate")
DATASETS = ("adc_clean", "adc", "Q", "pixels", "ticks", "eventID", "pix_x", "pix_y", "pix_z")


def batch_hits(ref_params, adcs, pixel_x, pixel_y, pixel_z, ticks, hit_prob, event, hit_pixels):
    """The per-batch post-processing of optimize/simulate.py:127-133 on the device: baseline-subtracted ADC, the mask of
    stored hits and their charge.  Returns host arrays keyed like the output datasets (eventID still batch-local)."""
    thr = torch.tensor([float(ref_params.DISCRIMINATION_THRESHOLD)], dtype=torch.float32, device=adcs.device)
    adc_lowest = fee.digitize(ref_params, thr)[0]
    adcs_clean = adcs - adc_lowest
    mask = (adcs_clean.flatten() != 0) & (event.flatten() != -1)
    Q = losses.adc2charge(adcs.flatten()[mask], ref_params)
    out = {"adc_clean": adcs_clean.flatten()[mask], "adc": adcs.flatten()[mask], "Q": Q, "pixels": hit_pixels[mask],
           "ticks": ticks.flatten()[mask], "eventID": event.flatten()[mask], "pix_x": pixel_x[mask], "pix_y": pixel_y[mask],
           "pix_z": pixel_z.flatten()[mask], "hit_prob": hit_prob.flatten()[mask]}
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def split_per_event(hits, global_event_ids):
    """{"event_<global id>": {dataset: array}} for the events that have stored hits (optimize/simulate.py:139-157)."""
    ev = hits["eventID"].astype(np.int64)
    groups = {}
    for local in np.unique(ev):
        if local < 0:
            continue
        gid = int(global_event_ids[local]) if local < len(global_event_ids) else int(local)
        sel = ev == local
        g = {k: hits[k][sel] for k in DATASETS if k != "eventID"}
        g["pixels"] = g["pixels"].astype(np.int32)
        g["eventID"] = np.full(int(sel.sum()), gid, dtype=np.int64)
        groups["event_%d" % gid] = {k: g[k] for k in DATASETS}
    return groups


def main(config):
    output_filename = config.output_file
    if not config.out_np and not output_filename.endswith(".h5"):
        output_filename += ".h5"
    if os.path.isfile(output_filename):
        os.remove(output_filename)
    if config.lut_file == "" and config.mode == "lut":
        return 1, 'Error: LUT file is required for mode "lut"'
    if getattr(config, "jac", False):
        return 1, "Error: --jac needs forward-mode differentiation, which the CUDA kernels do not provide"
    if not torch.cuda.is_available():
        raise RuntimeError("larndsim_b200 needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())

    Params = consts.build_params_class([])
    if config.detector_props and config.pixel_layouts:
        ref_params = consts.load_detector_properties(Params, config.detector_props, config.pixel_layouts)
    else:   # the derived module-0 geometry shipped with the package (the YAML files live in the reference checkout)
        ref_params = consts.load_geometry_json(Params, os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "module0_geometry.json"))
    if config.mode == "lut":
        response, ref_params = consts.load_lut(config.lut_file, ref_params, device=dev, cache_dir=getattr(config, "lut_cache", None))
    ref_params = ref_params.replace(diffusion_in_current_sim=config.diffusion_in_current_sim, mc_diff=config.mc_diff,
                                    electron_sampling_resolution=config.electron_sampling_resolution,
                                    number_pix_neighbors=config.number_pix_neighbors, signal_length=config.signal_length,
                                    time_window=config.signal_length)
    if not config.noise:
        ref_params = ref_params.replace(RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)

    dataset = dataio.TracksDataset(filename=config.input_file, nevents=config.n_events, max_nbatch=None, swap_xz=True,
                                   random_nevents=False, data_seed=config.seed if config.seed is not None else 42,
                                   max_batch_len=config.max_batch_len, print_input=False, chopped=config.chop, pad=False,
                                   electron_sampling_resolution=config.electron_sampling_resolution, live_selection=False, device=dev)
    fields = dataset.get_track_fields()
    tree, flat = {}, {k: [] for k in ("adc", "Q", "pix_x", "pix_y", "pix_z", "ticks", "hit_prob", "eventID")}
    n_segments = 0
    npix_cap = None   # pixel capacity of the hits-only arena, sized from the first batch
    for ibatch in range(len(dataset)):
        size = sim.pad_size(dataset.batch_nsteps[ibatch] if config.chop else len(dataset.batch_row_indices[ibatch]), "batch_size", 0.5)
        tracks = dataset.device_batch(ibatch, capacity=size)
        global_event_ids = dataset.get_batch_global_event_ids(ibatch)
        n_ev = len(global_event_ids)
        # the reference validates the local event-id namespace and the id packing on the host per batch
        # (optimize/simulate.py:111-113); the ids are local by construction here, the packing limit is the int32 one
        detsim.validate_event_ids_for_packing(ref_params, np.arange(n_ev, dtype=np.int64), kind="pixel", context="simulate batch %d" % ibatch)
        rngseed = ibatch if config.seed is None else config.seed
        if config.mode == "lut" and not config.save_wfs and npix_cap is not None:
            # hits-only batches after the first: one call over the persistent, self-cleaning waveform arena with a fixed pixel
            # capacity (no per-batch allocation / memset, one host synchronisation); the hits do not depend on the padding.
            # A batch with more pixels than the capacity is flagged on the device: grow and redo it.
            while True:
                try:
                    out = sim.simulate_hits(ref_params, response, tracks, fields, rngseed=rngseed, npix_capacity=npix_cap, n_events=n_ev)
                    break
                except _lib.LarndError as e:
                    if "npix_capacity" not in str(e):
                        raise
                    npix_cap *= 2
            wfs = None
        elif config.mode == "lut":
            wfs, unique_pixels = sim.simulate_wfs(ref_params, response, tracks, fields, n_events=n_ev)
            out = sim.simulate_stochastic(ref_params, wfs, unique_pixels, rngseed=rngseed)
            npix_cap = 2 * int(unique_pixels.shape[0])
        else:
            out = sim.simulate_parametrized(ref_params, tracks, fields, rngseed=rngseed, n_events=n_ev)
            wfs = None
        hits = batch_hits(ref_params, *out)
        n_segments += dataset.batch_nsteps[ibatch] if config.chop else len(dataset.batch_row_indices[ibatch])
        if not config.out_np:
            groups = split_per_event(hits, global_event_ids)
            if config.save_wfs and wfs is not None:
                w = wfs.detach().cpu().numpy()
                for g in groups.values():
                    g["wfs"] = w
            tree["batch_%d" % ibatch] = groups
        else:
            for k in flat:
                flat[k].append(hits[k])
    if not config.out_np:
        h5io.write_h5(output_filename, tree)
    else:
        cat = {k: np.concatenate(v) if v else np.zeros(0, np.float32) for k, v in flat.items()}
        np.savez(config.output_file, adcs=cat["adc"], Q=cat["Q"], x=cat["pix_x"], y=cat["pix_y"], z=cat["pix_z"], ticks=cat["ticks"],
                 hit_prob=cat["hit_prob"], event_id=cat["eventID"])
    logger.info("simulated %d batches, %d segments", len(dataset), n_segments)
    return 0, "Success"


def build_parser():
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("--input_file", required=True, help="Input data file (HDF5 with a 'segments' table)")
    p.add_argument("--output_file", required=True, help="Output data file")
    p.add_argument("--detector_props", default=None, help="Detector properties YAML (default: the module-0 geometry shipped in data/)")
    p.add_argument("--pixel_layouts", default=None, help="Pixel layout YAML")
    p.add_argument("--mode", choices=["lut", "parametrized"], default="lut")
    p.add_argument("--electron_sampling_resolution", type=float, required=True)
    p.add_argument("--number_pix_neighbors", type=int, required=True)
    p.add_argument("--signal_length", type=int, required=True)
    p.add_argument("--lut_file", type=str, default="")
    p.add_argument("--lut_cache", type=str, default=None, help="directory for the on-disk cache of the convolved template bank")
    p.add_argument("--noise", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--diffusion_in_current_sim", action="store_true")
    p.add_argument("--batch_size", type=float, default=500)
    p.add_argument("--gpu", action="store_true", help="accepted for compatibility (the simulation always runs on the GPU)")
    p.add_argument("--jac", action="store_true")
    p.add_argument("--mc_diff", action="store_true")
    p.add_argument("--save_wfs", action="store_true")
    p.add_argument("--n_events", type=int, default=-1)
    p.add_argument("--out_np", action="store_true", default=False)
    p.add_argument("--max_batch_len", type=float, default=50.)
    p.add_argument("--chop", action="store_true", default=False)
    return p


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    args = build_parser().parse_args()
    try:
        if args.save_wfs and args.jac:
            raise ValueError("Cannot save waveforms and compute jacobian at the same time. Please choose one of the two options.")
        retval, status = main(args)
    except Exception:
        import traceback
        print(traceback.format_exc(), file=sys.stderr)
        retval, status = 1, "Error: simulation failed."
    logger.info(status)
    sys.exit(retval)
