"""Host / device split of the production-driver loop over the prepared inputs (BASELINE config 2): wall per batch, cProfile of
the host side, device time per batch (torch profiler).  usage (GPU box): python scripts/prof_prepared.py"""
import cProfile, glob, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import numpy as np, torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic, dataio
from larndsim_b200.consts import build_response_template
import bench
dev = torch.device("cuda", 0)
lb.build_library()
params = lb.load_geometry_json(lb.build_params_class([]), bench.GEOM).replace(number_pix_neighbors=4, signal_length=100, RESET_NOISE_CHARGE=0,
                                                                              UNCORRELATED_NOISE_CHARGE=0, electron_sampling_resolution=0.005)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "segments_input_*.npz")))
dsets = [dataio.TracksDataset(np.load(f)["segments"], nevents=None, max_nbatch=None, swap_xz=True, max_batch_len=50, chopped=True, pad=False,
                              electron_sampling_resolution=0.005, device=dev) for f in files]
cap = {"npix": 0}
def run_all(hits_only):
    nh = 0
    for ds in dsets:
        f2 = ds.get_track_fields()
        for ib in range(len(ds)):
            c = sim.pad_size(ds.batch_nsteps[ib], "batch_size", 0.5)
            tr = ds.device_batch(ib, capacity=c)
            nev = len(ds.get_batch_global_event_ids(ib))
            if hits_only:
                nh += int(sim.simulate_hits(params, bank, tr, f2, rngseed=ib, npix_capacity=cap["npix"], n_events=nev)[0].shape[0])
            else:
                w, u = sim.simulate_wfs(params, bank, tr, f2, n_events=nev)
                cap["npix"] = max(cap["npix"], int(u.shape[0]))
                nh += int(sim.simulate_stochastic(params, w, u, rngseed=ib)[0].shape[0])
    return nh
run_all(False); run_all(True); torch.cuda.synchronize()
nb = sum(len(ds) for ds in dsets)
t0 = time.perf_counter(); run_all(True); torch.cuda.synchronize()
print("hits-only: %.3f ms/batch over %d batches" % ((time.perf_counter() - t0) / nb * 1e3, nb))
pr = cProfile.Profile(); pr.enable(); run_all(True); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run_all(True); torch.cuda.synchronize()
ev = prof.key_averages()
print("device ms/batch %.3f, device ops/batch %.1f" % (sum(e.device_time_total for e in ev) / 1e3 / nb, sum(e.count for e in ev if e.device_time_total > 0) / nb))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]: print("  %-72s %8.4f ms n=%.1f" % (e.key[:72], e.device_time_total / 1e3 / nb, e.count / nb))
