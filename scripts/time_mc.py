"""Times the MC-current-mode forward (prepare + unique + analytic current + scatter) and backward with CUDA events:
python scripts/time_mc.py 2000000 [path of an alternative liblarnd_b200.so]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "larnd-sim-jax_b200"))
import torch
from larndsim_b200 import _lib
if len(sys.argv) > 2:
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic, dataio
nseg = int(sys.argv[1])
dev = torch.device("cuda", 0)
GEOM = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "data", "module0_geometry.json")
raw, nev = synthetic.synthetic_raw_tracks(nseg, seed=1234, precision=0.01)
tracks = dataio.chop_tracks(torch.from_numpy(raw).to(dev), synthetic.FIELDS, 0.01)
for dic in (True, False):
    params = lb.load_geometry_json(lb.build_params_class([]), GEOM).replace(number_pix_neighbors=0, signal_length=150, mc_diff=True,
                                                                            diffusion_in_current_sim=dic, RESET_NOISE_CHARGE=0,
                                                                            UNCORRELATED_NOISE_CHARGE=0)
    rnd = sim.mc_normals(tracks.shape[0], 0, dev)
    st = sim.mc_forward(params, tracks, synthetic.FIELDS, rnd, n_events=nev)
    npix = st.npix
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(2):
        sim.mc_forward(params, tracks, synthetic.FIELDS, rnd, npix_capacity=npix, n_events=nev)
    e0.record()
    for i in range(5):
        st = sim.mc_forward(params, tracks, synthetic.FIELDS, rnd, npix_capacity=npix, n_events=nev)
    e1.record(); torch.cuda.synchronize()
    print("%s diffusion_in_current=%s segments=%d: %.3f ms/forward, checksum %.6e" %
          (os.path.basename(_lib.LIB_PATH), dic, tracks.shape[0], e0.elapsed_time(e1) / 5, float(st.wfs_full[:, 1:].double().abs().sum())))
    g = torch.ones_like(st.wfs_full[:, 1:]) * (st.unique_pixels >= 0).unsqueeze(1)
    grad = sim.mc_backward(st, tracks, g)
    e0.record()
    for i in range(5):
        grad = sim.mc_backward(st, tracks, g)
    e1.record(); torch.cuda.synchronize()
    print("   backward: %.3f ms, grad[:6] %s" % (e0.elapsed_time(e1) / 5, [float(x) for x in grad[:6]]))
