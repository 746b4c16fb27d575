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
kw=dict(number_pix_neighbors=2, signal_length=150)
op=cm.oracle_params(**kw); pp=cm.product_params(**kw)
bank=cm.synthetic_bank(32,25,25,1950)
tr=cm.small_batch(400, ifile=0, ibatch=1, pad=0, precision=0.01)
print('tracks',tr.shape)
# reference forward (f64) to fix the pixel list
wfs64, uniq, d, full64 = lo.simulate_wfs(op, bank, tr, cm.FIELDS, dt=np.float64, history={}, return_aux=True)
Npix=len(uniq)
rng=np.random.default_rng(5)
t=np.arange(2000)
G=(rng.uniform(0.5,1.5,(Npix,1))*(1+0.5*np.sin(t[None,:]/37.0+rng.uniform(0,6,(Npix,1))))).astype(np.float32)
def L_oracle(p):
    w,u=lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, pad_to=Npix)
    assert np.array_equal(u,uniq)
    return float((w*G.astype(np.float64)).sum())
st=sim.lut_forward(pp, torch.as_tensor(bank,device=dev), torch.as_tensor(tr,device=dev), cm.FIELDS, npix_capacity=Npix)
w32=st.wfs_full[:,1:].cpu().numpy()
print('fwd relerr vs f64', np.abs(w32-wfs64).max()/np.abs(wfs64).max())
grad=sim.lut_backward(st, torch.as_tensor(G,device=dev)).cpu().numpy()
names=_lib.PARAM_ORDER
steps=dict(Ab=1e-6,kb=1e-7,eField=1e-7,lifetime=1e-2,long_diff=1e-11,tran_diff=1e-11,shift_x=1e-6,shift_y=1e-6,shift_z=1e-6,lArDensity=1e-6,MeVToElectrons=1e-1)
for nme,h in steps.items():
    base=getattr(op,nme)
    fd=(L_oracle(op.replace(**{nme:base+h}))-L_oracle(op.replace(**{nme:base-h})))/(2*h)
    g=grad[names.index(nme)]
    print('LUT grad %-16s kernel % .6e  fd % .6e  rel %.2e'%(nme,g,fd,abs(g-fd)/(abs(fd)+1e-30)))
# box model
for mode,extra in ((oc.BOX,['alpha','beta']),(oc.ELLIPSOID,['alpha','beta','R_param'])):
    opm=op.replace(recombination_mode=mode); ppm=pp.replace(recombination_mode=lb.RecombinationMode(mode))
    stm=sim.lut_forward(ppm, torch.as_tensor(bank,device=dev), torch.as_tensor(tr,device=dev), cm.FIELDS, npix_capacity=Npix)
    gm=sim.lut_backward(stm, torch.as_tensor(G,device=dev)).cpu().numpy()
    def Lm(p):
        w,u=lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, pad_to=Npix); return float((w*G.astype(np.float64)).sum())
    for nme in extra+['eField']:
        base=getattr(opm,nme); h=1e-6
        fd=(Lm(opm.replace(**{nme:base+h}))-Lm(opm.replace(**{nme:base-h})))/(2*h)
        g=gm[names.index(nme)]
        print('mode %d grad %-10s kernel % .6e fd % .6e rel %.2e'%(mode,nme,g,fd,abs(g-fd)/(abs(fd)+1e-30)))
# ---- FEE backward: directional derivative
wfs_t=torch.as_tensor(w32,device=dev).requires_grad_(True)
up=torch.as_tensor(uniq,device=dev)
adcs,ticks,_,_,_=sim._FeeAdc.apply(wfs_t, pp, up, None)
ga=torch.as_tensor(rng.uniform(0.5,1.5,adcs.shape).astype(np.float32),device=dev)
(adcs*ga).sum().backward()
gw=wfs_t.grad.cpu().numpy()
delta=rng.normal(size=w32.shape)*np.abs(w32).max()*1e-3
def L_fee(w):
    a,tk=lo.get_adc_values(op, w, dt=np.float64); return float((lo.digitize(op,a,np.float64)*ga.cpu().numpy()).sum()), tk
eps=1e-4
lp,tkp=L_fee(w32.astype(np.float64)+eps*delta); lm,tkm=L_fee(w32.astype(np.float64)-eps*delta)
print('FEE ticks stable', np.array_equal(tkp,tkm), 'nhits', int((tkp<1997).sum()))
print('FEE directional: kernel % .6e fd % .6e'%(float((gw*delta).sum()), (lp-lm)/(2*eps)))
# ---- end-to-end autograd through simulate_wfs + simulate_stochastic + a loss
P=cm.product_params(grad=('Ab','kb','eField','lifetime','tran_diff','long_diff'), **kw)
bank_d=torch.as_tensor(bank,device=dev); trd=torch.as_tensor(tr,device=dev)
wfs,upix=sim.simulate_wfs(P, bank_d, trd, cm.FIELDS)
out=sim.simulate_stochastic(P, wfs, upix, 0)
loss=(out[0]**2).sum()*1e-4 + (out[3]**2).sum()*1e-3
loss.backward()
print('e2e loss', float(loss), {n: float(getattr(P,n).grad) for n in ('Ab','kb','eField','lifetime','tran_diff','long_diff')})
def L_e2e(p):
    w,u=lo.simulate_wfs(p, bank, tr, cm.FIELDS, dt=np.float64, history={})
    o=lo.simulate_stochastic(p, w, u, dt=np.float64)
    return float((o[0]**2).sum()*1e-4+(o[3]**2).sum()*1e-3), len(o[0])
print('oracle e2e', L_e2e(op), 'nhits kernel', len(out[0]))
for nme,h in dict(Ab=1e-6,kb=1e-7,eField=1e-7,lifetime=1e-2,tran_diff=1e-11,long_diff=1e-11).items():
    base=getattr(op,nme)
    (a,na),(b,nb)=L_e2e(op.replace(**{nme:base+h})),L_e2e(op.replace(**{nme:base-h}))
    print('e2e grad %-10s autograd % .6e fd % .6e (hits %d/%d)'%(nme,float(getattr(P,nme).grad),(a-b)/(2*h),na,nb))
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
