import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from proto_hybrid import hybrid
consts=large_ev_consts(); N=24; B=20000
rng=np.random.default_rng(0)
theta=consts.theta
lm=0.05*theta*rng.random((B,3*N))*(rng.random((B,3*N))<0.5)
lr=np.zeros(B); gam=consts.y_max-(0.3+0.2*rng.random(B))
w,f,iters,done,nfb,kkt=hybrid(N,consts,lm,lr,gam)
print('stuck',(iters<0).sum(),'fail',(~done).sum(),'iters mean',iters[iters>0].mean(),'max',iters.max(), 'fallbacks',nfb.mean())
bad=np.flatnonzero(iters<0)[:3]
for b in bad:
    wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
    print(b,'err',np.abs(w[b]-wo).max()/consts.w_max, 'it',iters[b])
    print(np.round(w[b]/consts.w_max,4)); print(np.round(wo/consts.w_max,4))
