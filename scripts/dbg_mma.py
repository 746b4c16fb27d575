"""Debug helper: compare the FFMA ("sorted") and tensor-core ("mma") consume loops of the class-sorted accumulate kernel
on the parity-test configurations and print where they differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import common as cm
from larndsim_b200 import sim
dev = torch.device("cuda", 0)
cfgs = [dict(n=4, L=100, prec=0.005, nseg=2500, pad=60, ibatch=1), dict(n=2, L=150, prec=0.01, nseg=2000, pad=0, ibatch=0),
        dict(n=0, L=100, prec=0.01, nseg=800, pad=10, ibatch=2), dict(n=1, L=30, prec=0.01, nseg=800, pad=10, ibatch=3)]
for cfg in cfgs:
    kw = dict(number_pix_neighbors=cfg["n"], signal_length=cfg["L"])
    nx = max(10 * cfg["n"] + 5, 5)
    bank = cm.synthetic_bank(32, nx, nx, 1950)
    tr = cm.small_batch(cfg["nseg"], ibatch=cfg["ibatch"], pad=cfg["pad"], precision=cfg["prec"])
    out = {}
    for impl in ("sorted", "mma"):
        os.environ["LARND_ACC_IMPL"] = impl
        pp = cm.product_params(**kw)
        st = sim.lut_forward(pp, torch.as_tensor(bank, device=dev), torch.as_tensor(tr, device=dev), cm.FIELDS)
        torch.cuda.synchronize()
        out[impl] = st.wfs_full.cpu().numpy()
        T0 = sim.record_fields(st)["T0"].cpu().numpy()
    a, b = out["sorted"], out["mma"]
    d = np.abs(a - b); sc = np.abs(a).max(axis=1, keepdims=True) + 1e-30
    rel = d / sc
    r, c = np.unravel_index(np.argmax(rel), rel.shape)
    print(cfg, "max rel", rel.max(), "at row", r, "tick", c, "vals", a[r, c], b[r, c], "T0 range", T0.min(), T0.max())
    bad = np.argwhere(rel > 2e-5)
    print("  n bad", len(bad), "rows", np.unique(bad[:, 0])[:20], "ticks", np.unique(bad[:, 1])[:40])
    for rr in np.unique(bad[:, 0])[:3]:
        cc = bad[bad[:, 0] == rr][:, 1]
        print("   row", rr, "ticks", cc[:12], "sorted", a[rr, cc[:6]], "mma", b[rr, cc[:6]])
