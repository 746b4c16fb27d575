"""Where a fit step (BASELINE config 4 size) spends its time: wall per step, cProfile of the host side, device ops per step.
usage (GPU box): python scripts/fit_latency.py [fast]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
import larndsim_b200 as lb
from larndsim_b200 import fit, synthetic
from larndsim_b200.consts import build_response_template
import bench
dev = torch.device("cuda", 0)
lb.build_library()
base4 = dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
fields = synthetic.FIELDS
fit_np, nev = synthetic.synthetic_tracks(bench.FIT_SEGMENTS, seed=500, precision=0.01)
tr = torch.as_tensor(fit_np, device=dev)
p4 = lb.load_geometry_json(lb.build_params_class(list(bench.FIT_NAMES)), bench.GEOM).replace(**base4)
bank4 = build_response_template(synthetic.synthetic_response(25, 25, 1950), p4, device=dev)
prob = fit.FitProblem.from_target_params(bench.FIT_NAMES, p4, bench.FIT_TARGET, bank4, tr, fields, nev)
mode = sys.argv[1] if len(sys.argv) > 1 else "adam"
stepper = fit.AdamFit(prob, bench.FIT_NOMINAL) if mode == "adam" else fit.FastAdamFit(prob, bench.FIT_NOMINAL)
for _ in range(10): stepper.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): l = stepper.step()
torch.cuda.synchronize()
print("%s: %.3f ms/step, loss %.4e, values %s" % (mode, (time.perf_counter() - t0) / 50 * 1e3, float(l), stepper.values()))
pr = cProfile.Profile(); pr.enable()
for _ in range(30): stepper.step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(38)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): stepper.step()
    torch.cuda.synchronize()
ev = prof.key_averages()
print("device ms/step %.3f, device ops/step %d" % (sum(e.device_time_total for e in ev) / 5e3, sum(e.count for e in ev if e.device_time_total > 0) / 5))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]: print("  %-72s %8.3f ms n=%d" % (e.key[:72], e.device_time_total / 5e3, e.count // 5))
