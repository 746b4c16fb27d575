import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/larnd-sim-jax_b200')
import numpy as np, torch
import common as cm
from oracle import larnd_oracle as lo
from larndsim_b200 import sim
import larndsim_b200 as lb
lb.build_library()
cfg=dict(n=2, L=150, prec=0.01, nseg=2000, pad=0, ibatch=0)
kw = dict(number_pix_neighbors=cfg["n"], signal_length=cfg["L"])
nx = max(10 * cfg["n"] + 5, 5)
bank = cm.synthetic_bank(32, nx, nx, 1950)
tr = cm.small_batch(cfg["nseg"], ibatch=cfg["ibatch"], pad=cfg["pad"], precision=cfg["prec"])
op = cm.oracle_params(**kw); pp=cm.product_params(**kw)
wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, history={}, return_aux=True)
dev='cuda'
st = sim.lut_forward(pp, torch.as_tensor(bank,device=dev), torch.as_tensor(tr,device=dev), cm.FIELDS, npix_capacity=len(uniq_o))
w=st.wfs_full.cpu().numpy()
scale=np.abs(full_o).max(axis=1,keepdims=True)
rel=np.abs(w-full_o)/(scale+1e-30)
idx=np.argsort(rel.ravel())[::-1][:15]
for i in idx:
    r,c=np.unravel_index(i,rel.shape)
    print('row',r,'pix',uniq_o[r],'col',c,'cuda',w[r,c],'oracle',full_o[r,c],'rel',rel[r,c],'rowmax',scale[r,0])
rec={k:v.cpu().numpy() for k,v in sim.record_fields(st).items()}
print('T0 range', rec['T0'].min(), rec['T0'].max(), 'q range', rec['Q'].min(), rec['Q'].max())
# float64 oracle for reference of f32 noise
w64,u64,_,full64=lo.simulate_wfs(op, bank, tr, cm.FIELDS, dt=np.float64, pad_to=len(uniq_o), return_aux=True)
rel64=np.abs(full_o-full64)/(scale+1e-30); print('oracle f32 vs f64 max rel', rel64.max(), np.unravel_index(np.argmax(rel64),rel64.shape))
relc=np.abs(w-full64)/(scale+1e-30); print('cuda vs f64 max rel', relc.max(), np.unravel_index(np.argmax(relc),relc.shape))
