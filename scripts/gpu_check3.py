"""Backward-pass checks on the GPU box: kernel VJPs vs central finite differences of the float64 oracle."""
import sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/larnd-sim-jax_b200')
import numpy as np, torch
import common as cm
from oracle import larnd_oracle as lo, consts as oc
import larndsim_b200 as lb
from larndsim_b200 import sim, _lib
lb.build_library()
dev='cuda'
tr=cm.small_batch(400, ifile=0, ibatch=1, pad=0, precision=0.01)
trd=torch.as_tensor(tr,device=dev)
rng=np.random.default_rng(5)
t=np.arange(2000)
names=_lib.PARAM_ORDER
# ---- MC mode
kwm=dict(number_pix_neighbors=0, signal_length=150, mc_diff=True)
opm=cm.oracle_params(**kwm); ppm=cm.product_params(**kwm)
rnd=rng.normal(size=(tr.shape[0],3)).astype(np.float32)
(out_o, wfull_o, uniq_o)=lo.simulate_parametrized(opm, tr, cm.FIELDS, rnd, history={}, return_wfs=True)
stm=sim.mc_forward(ppm, trd, cm.FIELDS, torch.as_tensor(rnd,device=dev), npix_capacity=len(uniq_o))
print('MC uniq eq', np.array_equal(stm.unique_pixels.cpu().numpy(), uniq_o), 'counts', stm.counts.cpu().numpy())
wm=stm.wfs_full.cpu().numpy(); valid=uniq_o>=0
err=np.abs(wm[valid]-wfull_o[valid]); print('MC wfs maxabs', err.max(), 'max', np.abs(wfull_o[valid]).max(), 'rel', err.max()/np.abs(wfull_o[valid]).max())
out_p=sim.simulate_parametrized(ppm, trd, cm.FIELDS, rnd=torch.as_tensor(rnd,device=dev), npix_capacity=len(uniq_o))
print('MC hits', len(out_o[0]), len(out_p[0]), 'ticks eq', np.array_equal(out_o[4], out_p[4].cpu().numpy()) if len(out_o[0])==len(out_p[0]) else None,
      'adc maxdiff', np.abs(out_o[0]-out_p[0].cpu().numpy()).max() if len(out_o[0])==len(out_p[0]) else None)
Gm=(rng.uniform(0.5,1.5,(len(uniq_o),1))*(1+0.5*np.sin(t[None,:]/37.0))).astype(np.float32); Gm[~valid]=0
gmc=sim.mc_backward(stm, trd, torch.as_tensor(Gm,device=dev)).cpu().numpy()
def L_mc(p):
    o,w,u=lo.simulate_parametrized(p, tr, cm.FIELDS, rnd, dt=np.float64, pad_to=len(uniq_o), return_wfs=True)
    return float((w[:,1:]*Gm.astype(np.float64)).sum())
for nme,h in dict(Ab=1e-6,kb=1e-7,eField=1e-7,lifetime=1e-2,long_diff=1e-11,tran_diff=1e-11,shift_x=1e-6,shift_y=1e-6,shift_z=1e-6).items():
    base=getattr(opm,nme)
    fd=(L_mc(opm.replace(**{nme:base+h}))-L_mc(opm.replace(**{nme:base-h})))/(2*h)
    g=gmc[names.index(nme)]
    print('MC grad %-10s kernel % .6e fd % .6e rel %.2e'%(nme,g,fd,abs(g-fd)/(abs(fd)+1e-30)))
