"""Times the class-sorted forward / backward at a given (segments, number_pix_neighbors, signal_length) with CUDA events:
python scripts/time_sorted_config.py 4000000 2 150      (LARND_SORTED_SPLIT=0/1 selects the kernel split)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic, dataio
from larndsim_b200.consts import build_response_template
nseg, nn, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
params = lb.load_geometry_json(lb.build_params_class([]), GEOM).replace(number_pix_neighbors=nn, signal_length=L, RESET_NOISE_CHARGE=0,
                                                                        UNCORRELATED_NOISE_CHARGE=0)
raw, nev = synthetic.synthetic_raw_tracks(nseg, seed=1234, precision=0.01)
tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
bank = build_response_template(synthetic.synthetic_response(), params, device=dev)
st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, n_events=nev)
npix = st.npix
lib = lb.get_lib()
lib.larnd_profile_enable(1)
import ctypes as C
buf = (C.c_float * 4)()
acc = []
for i in range(4):
    st = sim.lut_forward(params, bank, tracks, synthetic.FIELDS, npix_capacity=npix, n_events=nev)
    fs = sim.fee_forward(params, st.wfs_full[:, 1:], st.unique_pixels, None, compact=True)
    g = sim.fee_backward(fs, fs.adc * (st.unique_pixels >= 0).unsqueeze(1))
    sim.lut_backward(st, g)
    lib.larnd_profile_read(buf)
    acc.append(list(buf))
torch.cuda.synchronize()
import numpy as np
m = np.array(acc[1:]).mean(axis=0)
print("split=%s segments=%d n=%d L=%d: accumulate %.3f ms, backward %.3f ms, prepare %.3f, fee %.3f" %
      (os.environ.get("LARND_SORTED_SPLIT", "default"), tracks.shape[0], nn, L, m[1], m[2], m[0], m[3]))
