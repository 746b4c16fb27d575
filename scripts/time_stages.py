"""Per-kernel times of the forward (+ step-event backward) pass at the bench size, from the library's own CUDA-event slots
(larnd_profile_enable): python scripts/time_stages.py [nseg] [reps] [bwd].  Environment switches of the library
(LARND_FEE_WARPS, ...) are read once per process, so experiments run this script once per setting."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic, dataio, _lib
from larndsim_b200.consts import build_response_template
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
dev = torch.device("cuda", 0)
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
P = lb.build_params_class([])
params = lb.load_geometry_json(P, GEOM).replace(number_pix_neighbors=4, signal_length=100, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
raw, nev = synthetic.synthetic_raw_tracks(nseg, seed=1234, precision=0.01)
tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
npix = st.npix
lib = _lib.get_lib()
acc = [0.0] * 8
tot = 0.0
for i in range(reps + 2):
    if i == 2:
        lib.larnd_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, npix_capacity=npix, n_events=nev)
    fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
    if bwd:
        sim.hits_backward(st, fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        buf = (C.c_float * 8)()
        lib.larnd_profile_read(buf)
        for k in range(8):
            acc[k] += buf[k]
        tot += e0.elapsed_time(e1)
lib.larnd_profile_enable(0)
names = ["prepare", "accumulate", "backward", "fee_forward"]
print("segments %d npix %d  total %.3f ms  " % (tracks.shape[0], npix, tot / reps) +
      "  ".join("%s %.3f" % (names[k], acc[k] / reps) for k in range(4)) + "  env " +
      " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("LARND_")))
