#!/usr/bin/env python
"""2-D likelihood scan (loss + gradients on a grid over two detector parameters), grid points distributed over the
ranks of a torchrun launch and gathered with one all-gather — BASELINE.json config 5 ("2-D likelihood LUT scan sharded
across 8xB200").  The reference only has 1-D scans (optimize/fit_params.py:1140-1149).

    python examples/scan_2d.py --grid 16 --out scan.npz
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/scan_2d.py --grid 16
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
RANGES = {"eField": (0.45, 0.55), "lifetime": (500.0, 5000.0), "Ab": (0.78, 0.88), "kb": (0.04, 0.07)}  # optimize/ranges.py down/up


def run_scan(p1="eField", p2="lifetime", grid=16, n_segments=40000, device=None):
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
    tracks_np, n_events = synthetic.synthetic_tracks(n_segments, seed=5, precision=0.01)
    tracks = torch.as_tensor(tracks_np, device=device)
    p0 = lb.load_geometry_json(lb.build_params_class([]), GEOM).replace(**base)
    bank = build_response_template(synthetic.synthetic_response(25, 25, 1950), p0, device=device)
    with torch.no_grad():
        w, u = sim.simulate_wfs(p0, bank, tracks, fields, n_events=n_events)
        tgt = [t.clone() for t in sim.simulate_stochastic(p0, w, u, 0)]
    ref_Q = adc2charge(tgt[0], p0)
    a1, a2 = np.linspace(*RANGES[p1], grid), np.linspace(*RANGES[p2], grid)
    Params = lb.build_params_class([p1, p2])
    mine = parallel.scan_points_for_rank(grid * grid, rank, world)
    per_rank = (grid * grid + world - 1) // world
    local = torch.zeros((per_rank, 4), device=device)          # point index, loss, dloss/dp1, dloss/dp2
    local[:, 0] = -1
    for k, ipt in enumerate(mine):
        i, j = divmod(ipt, grid)
        params = lb.load_geometry_json(Params, GEOM).replace(**base, **{p1: float(a1[i]), p2: float(a2[j])})
        wfs, upix = sim.simulate_wfs(params, bank, tracks, fields, n_events=n_events)
        adcs, x, y, z, ticks, hp, ev, _ = sim.simulate_stochastic(params, wfs, upix, 0)
        loss, _ = mse_adc(params, adc2charge(adcs, params), x, y, z, ticks, hp, ev.float(), ref_Q, tgt[1], tgt[2], tgt[3], tgt[4],
                          tgt[5], tgt[6].float())
        loss.backward()
        local[k] = torch.stack([torch.tensor(float(ipt), device=device), loss.detach(), getattr(params, p1).grad.to(device),
                                getattr(params, p2).grad.to(device)])
    allpts = parallel.allgather(local).reshape(-1, 4).cpu().numpy()
    out = np.full((grid, grid, 3), np.nan)
    for ipt, l, g1, g2 in allpts:
        if ipt >= 0:
            out[int(ipt) // grid, int(ipt) % grid] = (l, g1, g2)
    return a1, a2, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--p1", default="eField")
    ap.add_argument("--p2", default="lifetime")
    ap.add_argument("--segments", type=int, default=40000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    a1, a2, out = run_scan(a.p1, a.p2, a.grid, a.segments)
    if int(os.environ.get("RANK", "0")) == 0:
        i, j = np.unravel_index(np.nanargmin(out[..., 0]), out.shape[:2])
        print("scan %s x %s: min loss %.4e at %s=%.4g %s=%.4g" % (a.p1, a.p2, out[i, j, 0], a.p1, a1[i], a.p2, a2[j]))
        if a.out:
            np.savez(a.out, p1=a1, p2=a2, loss=out[..., 0], grad1=out[..., 1], grad2=out[..., 2])
    if dist.is_initialized():
        dist.destroy_process_group()
