import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data,riccati,smooth_grad

def cost_fn(N,d,c,gh,gam,w,brk,slope):
    S=np.cumsum(w,axis=1)
    f=0.5*np.sum(d*w*w,axis=1)+np.sum(gh*w,axis=1)+0.5*c*np.sum((S-gam[:,None])**2,axis=1)
    # psi(w)= sum_j (slope_j - slope_{j-1}) (w-b_j)+
    for j in range(1,len(slope)):
        f=f+(slope[j]-slope[j-1])*np.sum(np.maximum(w-brk[j],0),axis=1)
    return f

def projnewton(N,consts,lm,lr,gam,max_it=100):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    w=np.zeros((B,N))
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int); nls=np.zeros(B,dtype=int)
    f=cost_fn(N,d,c,gh,gam,w,brk,slope)
    for it in range(max_it):
        S=np.cumsum(w,axis=1)
        q=smooth_grad(N,d,c,gh,gam,w,S)
        # breakpoint index if exactly at a breakpoint else -1
        atb=np.full((B,N),-1)
        for i in range(nseg+1):
            atb=np.where(w==brk[i],i,atb)
        segin=np.clip((w[...,None]>=brk[None,None,1:-1]).sum(-1),0,nseg-1)
        i=np.maximum(atb,0)
        go_r=(atb>=0)&(-q>shi[i])
        go_l=(atb>=0)&(-q<slo[i])
        binding=(atb>=0)&~go_r&~go_l
        seg=np.where(atb<0,segin,np.where(go_r,np.minimum(i,nseg-1),np.maximum(i-1,0)))
        conv=binding|((atb<0)&(np.abs(q+slope[seg])<=1e-12*np.maximum(1,np.abs(gh).max(axis=1))[:,None]))
        conv=conv.all(axis=1)
        newly=conv&~done; iters[newly]=it; done|=conv
        if done.all(): break
        h=gh+np.where(binding,0.0,slope[seg])
        wt,_=riccati(N,d,c,h,gam,binding,w)
        p=wt-w
        lo=brk[seg]; hi=brk[seg+1]
        alpha=np.ones(B)
        act=~done
        wn=w.copy(); fn=f.copy()
        for ls in range(30):
            cand=np.clip(w+alpha[:,None]*p,lo,hi)
            cand=np.where(binding,w,cand)
            fc=cost_fn(N,d,c,gh,gam,cand,brk,slope)
            ok=act&(fc<=f-1e-16*np.abs(f))
            take=ok
            wn=np.where(take[:,None],cand,wn); fn=np.where(take,fc,fn)
            nls[act]+=1
            act=act&~ok
            if not act.any(): break
            alpha=np.where(act,alpha*0.5,alpha)
        # those that failed line search: mark done (stuck)
        w=wn; f=fn
    iters[~done]=max_it
    return w,iters,done,nls

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=1000
    for consts in (small_ev_consts(),large_ev_consts()):
       for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            w,iters,done,nls=projnewton(N,consts,lm,lr,gam)
            err=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d  fail %d  ls/it %.2f err %.2e'%(iters[done].mean(),np.percentile(iters,99),iters.max(),(~done).sum(),nls.sum()/max(1,iters.sum()),err))
