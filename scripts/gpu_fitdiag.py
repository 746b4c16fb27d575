import os, sys
ROOT='/root/repo'; sys.path.insert(0, os.path.join(ROOT,'larnd-sim-jax_b200')); sys.path.insert(0, os.path.join(ROOT,'examples'))
import numpy as np, torch
import larndsim_b200 as lb
from larndsim_b200 import sim, synthetic
from larndsim_b200.consts import build_response_template
from larndsim_b200.losses import adc2charge, mse_adc
import fit_demo as fd
dev=torch.device('cuda',0)
base=dict(number_pix_neighbors=2, signal_length=150, electron_sampling_resolution=0.01, RESET_NOISE_CHARGE=0, UNCORRELATED_NOISE_CHARGE=0)
fields=synthetic.FIELDS
tr,nev=synthetic.synthetic_tracks(12000, seed=5, precision=0.01)
tracks=torch.as_tensor(tr,device=dev)
p0=lb.load_geometry_json(lb.build_params_class([]), fd.GEOM).replace(**base)
bank=build_response_template(synthetic.synthetic_response(25,25,1950), p0, device=dev)
ptgt=p0.replace(Ab=0.83)
with torch.no_grad():
    w,u=sim.simulate_wfs(ptgt,bank,tracks,fields,n_events=nev); tgt=[t.clone() for t in sim.simulate_stochastic(ptgt,w,u,0)]
refQ=adc2charge(tgt[0],ptgt)
print('target hits', len(tgt[0]), 'sumQ', float(refQ.sum()))
P=lb.build_params_class(['Ab'])
def loss_at(ab, grad=True):
    params=lb.load_geometry_json(P, fd.GEOM).replace(**base, Ab=ab)
    wfs,upix=sim.simulate_wfs(params,bank,tracks,fields,n_events=nev)
    adcs,x,y,z,ticks,hp,ev,_=sim.simulate_stochastic(params,wfs,upix,0)
    loss,aux=mse_adc(params, adc2charge(adcs,params), x,y,z,ticks,hp,ev.float(), refQ,tgt[1],tgt[2],tgt[3],tgt[4],tgt[5],tgt[6].float())
    loss.backward()
    return float(loss.detach()), float(params.Ab.grad), len(adcs), float(aux['charge_loss']), float(aux['mmd_loss_term'])
for ab in [0.78,0.79,0.80,0.81,0.82,0.825,0.83,0.835,0.84,0.86]:
    print('Ab %.3f loss %.4e grad % .4e nhits %d charge %.3e mmd %.3e'%((ab,)+loss_at(ab)))
h=1e-4
for ab in (0.80,0.82):
    lp=loss_at(ab+h)[0]; lm=loss_at(ab-h)[0]; print('FD at',ab,(lp-lm)/(2*h))
