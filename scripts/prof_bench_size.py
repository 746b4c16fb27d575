"""ncu driver at the benchmark size: two forward(+backward) steps of the bench workload (10 M segments, device chop)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic, dataio
from larndsim_b200.consts import build_response_template
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
bwd = len(sys.argv) > 2 and sys.argv[2] in ("bwd", "steps")
steps = len(sys.argv) > 2 and sys.argv[2] == "steps"
dev = torch.device("cuda", 0)
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
P = lb.build_params_class([])
params = lb.load_geometry_json(P, GEOM).replace(number_pix_neighbors=4, signal_length=100, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
raw, nev = synthetic.synthetic_raw_tracks(nseg, seed=1234, precision=0.01)
tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
npix = st.npix
for i in range(2):
    st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, npix_capacity=npix, n_events=nev)
    fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
    if steps:
        sim.hits_backward(st, fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
    elif bwd:
        g = sim.fee_backward(fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
        sim.lut_backward(st, g)
torch.cuda.synchronize()
print("done", tracks.shape[0], npix)
