import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data,riccati,smooth_grad

def pdas3(N,consts,lm,lr,gam,max_it=60,tolw=1e-11,tolq=1e-11,init='free',code0=None,trace=False):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    tw=tolw*consts.w_max
    tq=tolq*np.maximum(1.0,np.abs(gh).max(axis=1)+c*N*consts.y_max)[:,None]
    code=np.ones((B,N),dtype=int) if init=='free' else np.zeros((B,N),dtype=int)
    if code0 is not None: code=code0.copy()
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int)
    tr=[]
    for it in range(max_it):
        fixed=(code%2==0); fval=brk[np.minimum(code//2,nseg)]
        seg=np.minimum(code//2,nseg-1)
        h=gh+np.where(fixed,0.0,slope[seg])
        w,S=riccati(N,d,c,h,gam,fixed,fval)
        q=smooth_grad(N,d,c,gh,gam,w,S)
        new=code.copy()
        lo=brk[seg]; hi=brk[seg+1]
        fr=~fixed
        new=np.where(fr&(w<lo-tw),2*seg,new)
        new=np.where(fr&(w>hi+tw),2*(seg+1),new)
        i=code//2
        rel_r=fixed&(-q>shi[np.minimum(i,nseg)]+tq)
        rel_l=fixed&(-q<slo[np.minimum(i,nseg)]-tq)
        new=np.where(rel_r,2*i+1,new)
        new=np.where(rel_l,2*i-1,new)
        if trace: tr.append(code.copy())
        conv=(new==code).all(axis=1)
        newly=conv&~done; iters[newly]=it+1; done|=conv
        if done.all(): break
        code=np.where(done[:,None],code,new)
    iters[~done]=max_it
    return w,iters,done,code,tr

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=1000
    for consts in (large_ev_consts(),):
       for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            w,iters,done,code,_=pdas3(N,consts,lm,lr,gam)
            err=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                if done[b]: err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d  fail %d  err %.2e'%(iters[done].mean(),np.percentile(iters,99),iters.max(),(~done).sum(),err))
            if (~done).any() and N==24:
                b=np.flatnonzero(~done)[0]
                w1,it1,dn1,cd1,tr=pdas3(N,consts,lm[b:b+1],lr[b:b+1],gam[b:b+1],trace=True)
                for t in tr[-6:]: print('    ',''.join(str(x) for x in t[0]))
