import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from proto_pdas import make_batch
from proto_ddp import ddp
import proto_ddp
rng=np.random.default_rng(1)
consts=large_ev_consts(); N=24
for mode in (2,3):
    lm,lr,gam=make_batch(rng,N,consts,1000,mode)
    w,iters,done=ddp(N,consts,lm,lr,gam)
    bad=np.flatnonzero(~done)[:3]
    for b in bad:
        ws=[]
        for mi in (96,97,98,99,100):
            w1,_,_=ddp(N,consts,lm[b:b+1],lr[b:b+1],gam[b:b+1],max_it=mi)
            ws.append(w1[0])
        wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
        print('mode',mode,'b',b,'gam',gam[b])
        for x in ws: print('   ',np.round(x/consts.w_max,3), 'cost',lompc_cost(N,consts,x,lm[b],lr[b],gam[b])-co)
        print(' opt',np.round(wo/consts.w_max,3))
