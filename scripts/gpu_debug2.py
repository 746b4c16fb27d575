import sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/larnd-sim-jax_b200')
import numpy as np, torch
import common as cm
from oracle import larnd_oracle as lo
from larndsim_b200 import sim
import larndsim_b200 as lb
lb.build_library()
rng = np.random.default_rng(3)
tr = cm.small_batch(900, ibatch=1, pad=0, precision=0.01)
c = cm.FIELDS.index
near = tr[:300].copy()
shift = 30.45 - np.abs(near[:, c("z")]).max()
for col in ("z", "z_start", "z_end"):
    near[:, c(col)] += np.sign(near[:, c(col)]) * shift
far = tr[300:360].copy()
far[:, c("x")] += 100.0
allr = np.concatenate([tr, near, far])
perm = rng.permutation(len(allr))
kw = dict(number_pix_neighbors=2, signal_length=100)
bank = cm.synthetic_bank(48, 25, 25, 1950)
op = cm.oracle_params(**kw); pp=cm.product_params(**kw)
for name,arr in (('sorted',allr),('shuffled',allr[perm]),('tr only',tr),('near only',near),('far only', far)):
    wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, arr, cm.FIELDS, history={}, return_aux=True)
    st = sim.lut_forward(pp, torch.as_tensor(bank,device='cuda'), torch.as_tensor(arr,device='cuda'), cm.FIELDS, npix_capacity=len(uniq_o))
    w=st.wfs_full.cpu().numpy()
    print(name, 'uniq eq', np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o), 'counts', st.counts.cpu().numpy())
    scale=np.abs(full_o[:,1:]).max(axis=1,keepdims=True)
    rel=np.abs(w[:,1:]-full_o[:,1:])/(scale+1e-30)
    rel[np.broadcast_to(scale<1e-3, rel.shape)]=0
    i=np.argmax(rel); r,cc=np.unravel_index(i,rel.shape)
    print('   worst col>=1: row',r,'pix',uniq_o[r],'col',cc+1,'cuda',w[r,cc+1],'oracle',full_o[r,cc+1],'rel',rel[r,cc],'rowmax',scale[r,0])
    g=np.abs(w[:,0]-full_o[:,0]); j=np.argmax(g/(np.maximum(np.abs(full_o[:,0]),np.abs(full_o).max(axis=1))+1e-30))
    print('   worst col0: row',j,'cuda',w[j,0],'oracle',full_o[j,0],'rowmax',np.abs(full_o[j]).max())
    print('   nan', np.isnan(w).sum(), np.isnan(full_o).sum(), 'idx max', sim.record_fields(st)['IDX'].max().item())
