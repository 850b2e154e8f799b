import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import lompc_oracle as orc
from oracle.price_oracle import PriceOracle, nnqp_exact
import proto_nnqp as pn
rng=np.random.default_rng(0)
consts=orc.large_ev_consts(); N=12; pt='linear-convex'; lmbd_r=0.0
po=PriceOracle(N,consts,pt); pn.orc_wmax=consts.w_max
w=consts.w_max*rng.random(N)*(rng.random(N)<0.8); w_ref=consts.w_max*rng.random(N)
lam_k=consts.theta*0.05*rng.random(po.r)*(rng.random(po.r)<0.6)
A_bar,A_bar_inv=po._metric(lmbd_r)
P,q=po.price_step_matrices(A_bar_inv,w_ref,w,lam_k)
qs=3*consts.theta/(4*consts.w_max); m=po.m; eps=po.eps_reg; theta=consts.theta; kappa=0.0
nb=3; coef=np.zeros((nb,N)); coef[0]=theta; coef[1]=-theta; coef[2]=2*qs*w
dk=np.full(N,kappa)
l=rng.random(po.r)
u=(coef*l.reshape(nb,N)).sum(0); v=pn.ric_solve(dk,1.0,u)
print('A_bar_inv u err',np.abs(v-A_bar_inv@u).max())
Pl=eps*l+(coef*v[None,:]).reshape(-1)/(2*m)
print('P l err',np.abs(Pl-P@l).max(), np.abs(P@l).max())
x=nnqp_exact(P,q)
lam,it,st=pn.nnqp_pn(N,po.r,theta,qs,m,eps,kappa,w,w_ref,lam_k,trace=True,max_it=8)
