"""Small driver for ncu: a few forward (+ optional backward) steps on a synthetic batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic
from larndsim_b200.consts import build_response_template
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
dev = torch.device("cuda", 0)
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
P = lb.build_params_class([])
params = lb.load_geometry_json(P, GEOM).replace(number_pix_neighbors=4, signal_length=100, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
tr, nev = synthetic.synthetic_tracks(nseg, seed=1234, precision=0.01)
tracks = torch.from_numpy(tr).to(dev)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
npix = st.npix
for i in range(steps):
    st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, npix_capacity=npix, n_events=nev)
    fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
    if bwd:
        g = sim.fee_backward(fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
        sim.lut_backward(st, g, skip_garbage=True)
torch.cuda.synchronize()
print("done", nseg, npix)
