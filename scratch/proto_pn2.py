import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data,riccati,smooth_grad
from proto_pn import cost_fn

def projnewton2(N,consts,lm,lr,gam,max_it=100,tol=1e-11,verbose=False,refix=0):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    gscale=np.maximum(1.0,np.abs(gh).max(axis=1)+c*N*consts.y_max)[:,None]
    tq=tol*gscale
    w=np.zeros((B,N))
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int); nls=np.zeros(B,dtype=int); nric=np.zeros(B,dtype=int)
    f=cost_fn(N,d,c,gh,gam,w,brk,slope)
    Hkk=d+c*np.arange(N,0,-1)[None,:]
    for it in range(max_it):
        S=np.cumsum(w,axis=1)
        q=smooth_grad(N,d,c,gh,gam,w,S)
        atb=np.full((B,N),-1)
        for i in range(nseg+1):
            atb=np.where(w==brk[i],i,atb)
        segin=np.clip((w[...,None]>=brk[None,None,1:-1]).sum(-1),0,nseg-1)
        i=np.maximum(atb,0)
        go_r=(atb>=0)&(-q>shi[i]+tq)
        go_l=(atb>=0)&(-q<slo[i]-tq)
        binding=(atb>=0)&~go_r&~go_l
        seg=np.where(atb<0,segin,np.where(go_r,np.minimum(i,nseg-1),np.maximum(i-1,0)))
        conv=binding|((atb<0)&(np.abs(q+slope[seg])<=tq))
        conv=conv.all(axis=1)
        newly=conv&~done; iters[newly]=it; done|=conv
        if done.all(): break
        lo=brk[seg]; hi=brk[seg+1]
        fixed=binding.copy(); fval=w.copy()
        h=gh+slope[seg]
        wt,_=riccati(N,d,c,h,gam,fixed,fval); nric[~done]+=1
        for rf in range(refix):
            out_lo=~fixed&(wt<lo); out_hi=~fixed&(wt>hi)
            if not (out_lo|out_hi).any(): break
            fval=np.where(out_lo,lo,np.where(out_hi,hi,fval)); fixed=fixed|out_lo|out_hi
            wt,_=riccati(N,d,c,h,gam,fixed,fval); nric[~done]+=1
        p=wt-w
        alpha=np.ones(B)
        act=~done
        wn=w.copy(); fn=f.copy()
        for ls in range(40):
            cand=np.clip(w+alpha[:,None]*p,lo,hi)
            cand=np.where(binding,w,cand)
            fc=cost_fn(N,d,c,gh,gam,cand,brk,slope)
            ok=act&(fc<f)
            wn=np.where(ok[:,None],cand,wn); fn=np.where(ok,fc,fn)
            nls[act]+=1
            act=act&~ok
            if not act.any(): break
            alpha=np.where(act,alpha*0.5,alpha)
        if act.any() and verbose:
            b=np.flatnonzero(act)[0]
            print('stuck',it,b,'f',f[b]); 
            print(' w/wm',np.round(w[b]/consts.w_max,4)); print(' p',p[b]); print(' q+s',(q+slope[seg])[b]); print(' bind',binding[b].astype(int)); print('seg',seg[b])
            return
        w=wn; f=fn
    iters[~done]=max_it
    return w,iters,done,nls,nric

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=1000
    for refix in (0,3):
     print('refix',refix)
     for consts in (small_ev_consts(),large_ev_consts()):
       for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            r=projnewton2(N,consts,lm,lr,gam,verbose=False,refix=refix)
            if r is None: continue
            w,iters,done,nls,nric=r
            err=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d  fail %d  ls/it %.2f ric mean %.1f err %.2e'%(iters[done].mean(),np.percentile(iters,99),iters.max(),(~done).sum(),nls.sum()/max(1,iters.sum()),nric.mean(),err))
