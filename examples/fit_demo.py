#!/usr/bin/env python
"""Gradient fit of detector parameters through the CUDA hot path — the counterpart of the reference's
`python -m optimize.example_run --fit_type chain --mode lut` (optimize/fit_test.sh --lut settings: 0.01 cm sampling,
2 neighbours, signal_length 150, loss mse_adc, Adam).  Events are sharded over the ranks of a torchrun launch; the only
collectives are the all-reduce of the loss sums and of the parameter gradients.

    python examples/fit_demo.py [--iterations 30] [--params Ab,kb,eField,lifetime]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/fit_demo.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import larndsim_b200 as lb  # noqa: E402
from larndsim_b200 import parallel, sim, synthetic  # noqa: E402
from larndsim_b200.consts import build_response_template  # noqa: E402
from larndsim_b200.losses import adc2charge, mse_adc  # noqa: E402

GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
NOMINAL = dict(Ab=0.8, kb=0.0486, eField=0.5, lifetime=2.2e3, long_diff=4.0e-6, tran_diff=8.8e-6)
TARGET = dict(Ab=0.83, kb=0.055, eField=0.52, lifetime=1.8e3, long_diff=5.0e-6, tran_diff=10e-6)


def run_fit(names=("Ab", "eField"), iterations=30, n_segments=40000, lr=0.01, seed=5, device=None, verbose=True):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    lb.build_library()
    base = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0,
                UNCORRELATED_NOISE_CHARGE=0)
    fields = synthetic.FIELDS
    tracks_all, _ = synthetic.synthetic_tracks(n_segments, seed=seed, precision=0.01)
    tracks_np, n_events, _ = parallel.shard_tracks(tracks_all, fields, rank, world)
    tracks = torch.as_tensor(tracks_np, device=device)
    p_static = lb.load_geometry_json(lb.build_params_class([]), GEOM).replace(**base)
    bank = build_response_template(synthetic.synthetic_response(25, 25, 1950), p_static, device=device)
    # target hits: same events simulated with the target parameters
    p_tgt = p_static.replace(**{n: TARGET[n] for n in names})
    with torch.no_grad():
        w, u = sim.simulate_wfs(p_tgt, bank, tracks, fields, n_events=n_events)
        tgt = [t.clone() for t in sim.simulate_stochastic(p_tgt, w, u, 0)]
    ref_Q = adc2charge(tgt[0], p_tgt)
    # fitted parameters are normalised by their nominal value, as the reference's ParamFitter does
    theta = torch.ones(len(names), device=device, requires_grad=True)
    opt = torch.optim.Adam([theta], lr=lr)
    Params = lb.build_params_class(list(names))
    p_fit = lb.load_geometry_json(Params, GEOM).replace(**base)  # the geometry file is read once, not per iteration
    reduce = parallel.allreduce_sum_differentiable if world > 1 else None
    history = []
    for it in range(iterations):
        opt.zero_grad()
        vals = {n: theta[i] * NOMINAL[n] for i, n in enumerate(names)}
        params = p_fit.replace(**vals)
        wfs, upix = sim.simulate_wfs(params, bank, tracks, fields, n_events=n_events)
        adcs, x, y, z, ticks, hp, ev, _ = sim.simulate_stochastic(params, wfs, upix, 0)
        loss, aux = mse_adc(params, adc2charge(adcs, params), x, y, z, ticks, hp, ev.float(), ref_Q, tgt[1], tgt[2], tgt[3], tgt[4],
                            tgt[5], tgt[6].float(), reduce=reduce)
        loss.backward()
        parallel.allreduce_sum_(theta.grad)
        opt.step()
        history.append((float(loss.detach()), theta.detach().cpu().numpy().copy()))
        if verbose and rank == 0:
            print("iter %3d  loss %.6e  %s" % (it, history[-1][0], {n: float(theta[i].detach() * NOMINAL[n]) for i, n in enumerate(names)}), flush=True)
    return history, {n: TARGET[n] / NOMINAL[n] for n in names}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--params", default="Ab,eField")
    ap.add_argument("--segments", type=int, default=40000)
    a = ap.parse_args()
    run_fit(tuple(a.params.split(",")), a.iterations, a.segments)
    if dist.is_initialized():
        dist.destroy_process_group()
