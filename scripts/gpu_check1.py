import sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/larnd-sim-jax_b200')
import numpy as np, torch
import common as cm
from oracle import larnd_oracle as lo
import larndsim_b200 as lb
from larndsim_b200 import sim
lb.build_library()
op = cm.oracle_params(); pp = cm.product_params()
bank = cm.synthetic_bank(); cum = cm.synthetic_bank_cum()
tr = cm.small_batch(1500, pad=100)
print('tracks', tr.shape)
t=time.time(); wfs_o, uniq_o, d, full_o = lo.simulate_wfs(op, bank, tr, cm.FIELDS, pad_to=None, history={}, response_cum=cum, return_aux=True); print('oracle s', time.time()-t)
dev='cuda'
bank_d = torch.as_tensor(bank, device=dev); trd = torch.as_tensor(tr, device=dev)
st = sim.lut_forward(pp, bank_d, trd, cm.FIELDS, npix_capacity=len(uniq_o))
torch.cuda.synchronize()
print('counts', st.counts.cpu().numpy(), 'oracle nuniq', len(np.unique(d['main_pixels'])))
rec = {k:v.cpu().numpy() for k,v in sim.record_fields(st).items()}
print('mainpix eq', np.array_equal(rec['MAINPIX'], d['main_pixels']))
print('bx eq', np.array_equal(rec['BX'], d['bins_pitches'][:,0]), 'by eq', np.array_equal(rec['BY'], d['bins_pitches'][:,1]))
ts=np.float32(0.1); ft=d['t0_neigh']/ts; ct=np.clip(np.floor(ft).astype(np.int32),0,1949)
print('T0 eq', np.array_equal(rec['T0'], 1950-100-ct), 'frac maxdiff', np.abs(rec['FRAC']-(ft-ct)).max())
tv=np.asarray(op.long_diff_template,dtype=np.float32); idx=np.clip(np.searchsorted(tv,d['long_diff_seg']),1,98)
print('idx eq', np.array_equal(rec['IDX'], idx), 'q relerr', np.nanmax(np.abs(rec['Q']-d['nelectrons_neigh'])/(np.abs(d['nelectrons_neigh'])+1e-9)))
wx=np.stack([rec['WX%d'%i] for i in range(5)],1); print('wx maxabs', np.abs(wx-d['wx']).max(), 'wy', np.abs(np.stack([rec['WY%d'%i] for i in range(5)],1)-d['wy']).max())
print('uniq eq', np.array_equal(st.unique_pixels.cpu().numpy(), uniq_o))
w = st.wfs_full.cpu().numpy()
err = np.abs(w-full_o); scale = np.abs(full_o).max(axis=1, keepdims=True)+1e-30
print('wfs max abs err', err.max(), 'max |wfs|', np.abs(full_o).max(), 'max rel-to-rowmax', (err/scale)[np.abs(full_o).max(axis=1)>1].max())
print('col0 err', err[:,0].max(), np.abs(full_o[:,0]).max(), 'row0 relerr', err[0].max()/ (np.abs(full_o[0]).max()+1e-30))
worst = np.unravel_index(np.argmax(err[:,1:]), err[:,1:].shape); print('worst', worst, w[worst[0],worst[1]+1], full_o[worst[0],worst[1]+1])
# FEE
out_o = lo.simulate_stochastic(op, wfs_o, uniq_o)
out_p = sim.simulate_stochastic(pp, st.wfs_full[:,1:], st.unique_pixels, 0)
names=['adc','x','y','z','ticks','hp','event','pix']
print('nhits', len(out_o[0]), len(out_p[0]))
for n,a,b in zip(names,out_o,out_p):
    b=b.cpu().numpy()
    if len(a)==len(b): print(n, 'maxdiff', np.abs(a.astype(np.float64)-b).max() if len(a) else 0)
# FEE on oracle wfs (isolates FEE kernel)
out_p2 = sim.simulate_stochastic(pp, torch.as_tensor(wfs_o,device=dev), torch.as_tensor(uniq_o,device=dev), 0)
for n,a,b in zip(names,out_o,out_p2):
    b=b.cpu().numpy(); print('iso',n, len(a), len(b), 'exact', np.array_equal(a,b), 'maxdiff', np.abs(a.astype(np.float64)-b).max() if len(a)==len(b) and len(a) else None)
# timing
for _ in range(3):
    torch.cuda.synchronize(); t=time.time(); st2 = sim.lut_forward(pp, bank_d, trd, cm.FIELDS, npix_capacity=len(uniq_o)); torch.cuda.synchronize(); print('fwd ms', (time.time()-t)*1e3)
