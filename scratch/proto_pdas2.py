import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data,riccati,smooth_grad

def pdas2(N,consts,lm,lr,gam,max_it=60,variant=1,init='free'):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    if init=='free':
        code=np.ones((B,N),dtype=int)  # free in seg 0
    else:
        code=np.zeros((B,N),dtype=int)
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int)
    for it in range(max_it):
        fixed=(code%2==0); fval=brk[np.minimum(code//2,nseg)]
        seg=np.minimum(code//2,nseg-1)
        h=gh+np.where(fixed,0.0,slope[seg])
        w,S=riccati(N,d,c,h,gam,fixed,fval)
        q=smooth_grad(N,d,c,gh,gam,w,S)
        new=code.copy()
        # free coords out of their segment
        lo=brk[seg]; hi=brk[seg+1]
        fr=~fixed
        if variant==1:
            new=np.where(fr&(w<lo),2*seg,new)
            new=np.where(fr&(w>hi),2*(seg+1),new)
        else:
            # move to the containing segment as free; clamp at box ends
            segw=np.clip(np.searchsorted(brk,w,side='right')-1,0,nseg-1) if False else np.clip((w[...,None]>=brk[None,None,1:-1]).sum(-1),0,nseg-1)
            out=fr&((w<lo)|(w>hi))
            cand=2*segw+1
            cand=np.where(w<=brk[0],0,cand); cand=np.where(w>=brk[-1],2*nseg,cand)
            new=np.where(out,cand,new)
        # fixed coords with multiplier outside the subdifferential
        i=code//2
        rel_r=fixed&(-q>shi[np.minimum(i,nseg)])
        rel_l=fixed&(-q<slo[np.minimum(i,nseg)])
        new=np.where(rel_r,2*i+1,new)
        new=np.where(rel_l,2*i-1,new)
        conv=(new==code).all(axis=1)
        newly=conv&~done; iters[newly]=it+1; done|=conv
        if done.all(): break
        code=np.where(done[:,None],code,new)
    iters[~done]=max_it
    return w,iters,done

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=1000
    for variant in (1,2):
     for init in ('free','zero'):
      print('variant',variant,'init',init)
      for consts in (small_ev_consts(),large_ev_consts()):
       for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            w,iters,done=pdas2(N,consts,lm,lr,gam,variant=variant,init=init)
            err=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                if done[b]: err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d  fail %d  err %.2e'%(iters[done].mean(),np.percentile(iters,99),iters.max(),(~done).sum(),err))
