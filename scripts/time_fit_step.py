"""Wall-clock vs device time of one fit step (forward + mse_adc loss + backward + Adam) at the reference's fit batch size
(optimize/fit_test.sh --lut: ~20 k segments, n = 2, L = 150): shows how much of a small-batch step is host/launch overhead."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200")); sys.path.insert(0, os.path.join(ROOT, "examples"))
import numpy as np, torch
import larndsim_b200 as lb
from larndsim_b200 import parallel, sim, synthetic
from larndsim_b200.consts import build_response_template
from larndsim_b200.losses import adc2charge, mse_adc
import fit_demo as fd

nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
names = ("Ab", "kb", "eField", "lifetime", "tran_diff", "long_diff")
dev = torch.device("cuda", 0)
lb.build_library()
base = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
fields = synthetic.FIELDS
tracks_np, _ = synthetic.synthetic_tracks(nseg, seed=5, precision=0.01)
tracks_np, n_events, _ = parallel.shard_tracks(tracks_np, fields, 0, 1)
tracks = torch.as_tensor(tracks_np, device=dev)
p_static = lb.load_geometry_json(lb.build_params_class([]), fd.GEOM).replace(**base)
bank = build_response_template(synthetic.synthetic_response(25, 25, 1950), p_static, device=dev)
p_tgt = p_static.replace(**{n: fd.TARGET[n] for n in names})
with torch.no_grad():
    w, u = sim.simulate_wfs(p_tgt, bank, tracks, fields, n_events=n_events)
    tgt = [t.clone() for t in sim.simulate_stochastic(p_tgt, w, u, 0)]
ref_Q = adc2charge(tgt[0], p_tgt)
theta = torch.ones(len(names), device=dev, requires_grad=True)
opt = torch.optim.Adam([theta], lr=0.01)
Params = lb.build_params_class(list(names))
p_fit = lb.load_geometry_json(Params, fd.GEOM).replace(**base)

def step():
    opt.zero_grad()
    vals = {n: theta[i] * fd.NOMINAL[n] for i, n in enumerate(names)}
    params = p_fit.replace(**vals)
    wfs, upix = sim.simulate_wfs(params, bank, tracks, fields, n_events=n_events)
    adcs, x, y, z, ticks, hp, ev, _ = sim.simulate_stochastic(params, wfs, upix, 0)
    loss, aux = mse_adc(params, adc2charge(adcs, params), x, y, z, ticks, hp, ev.float(), ref_Q, tgt[1], tgt[2], tgt[3], tgt[4], tgt[5], tgt[6].float())
    loss.backward()
    opt.step()
    return loss

for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 20
for _ in range(K): step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / K * 1e3
ms0 = torch.cuda.memory_stats()
per = []
for _ in range(10):
    t1 = time.perf_counter(); step(); torch.cuda.synchronize(); per.append((time.perf_counter() - t1) * 1e3)
ms1 = torch.cuda.memory_stats()
print("per-step wall (ms):", " ".join("%.2f" % v for v in per))
for k in ("segment.all.allocated", "segment.all.freed", "num_alloc_retries", "allocation.all.allocated", "num_device_alloc", "num_device_free"):
    print("  ", k, ms1.get(k, 0) - ms0.get(k, 0), "over 10 steps; reserved MB", ms1.get("reserved_bytes.all.current", 0) / 1e6)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
dev_ms = sum(e.device_time_total for e in ev) / 5 / 1e3
nk = sum(e.count for e in ev if e.device_time_total > 0) / 5
print("segments %d  hits %d  wall %.3f ms/step  device %.3f ms/step  ~%d device ops/step" % (tracks.shape[0], tgt[0].numel(), wall, dev_ms, nk))
top = sorted(ev, key=lambda e: -e.device_time_total)[:12]
for e in top: print("  %-70s %8.3f ms  n=%d" % (e.key[:70], e.device_time_total / 5 / 1e3, e.count // 5))
topc = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:10]
for e in topc: print("  cpu %-66s %8.3f ms  n=%d" % (e.key[:66], e.self_cpu_time_total / 5 / 1e3, e.count // 5))
