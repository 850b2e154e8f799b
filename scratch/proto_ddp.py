import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data,smooth_grad
from proto_pn import cost_fn

def ddp(N,consts,lm,lr,gam,max_it=100,tol=1e-11,init_code=None):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    # config: fixed flag, seg (if free), value (if fixed)
    fixed=np.zeros((B,N),dtype=bool); seg=np.zeros((B,N),dtype=int); fval=np.zeros((B,N))
    code=np.full((B,N),-7)
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int)
    fhist=[]
    w=np.zeros((B,N))
    for it in range(max_it):
        # backward
        P=np.zeros(B); r=np.zeros(B)
        Qs=np.zeros((B,N)); rps=np.zeros((B,N))
        for k in range(N-1,-1,-1):
            Q=c+P; rp=r-c*gam
            Qs[:,k]=Q; rps[:,k]=rp
            h=gh[:,k]+slope[seg[:,k]]
            inv=1.0/(d[:,k]+Q)
            Pf=Q*d[:,k]*inv; rf=(d[:,k]*rp-Q*h)*inv
            Px=Q; rx=Q*fval[:,k]+rp
            P=np.where(fixed[:,k],Px,Pf); r=np.where(fixed[:,k],rx,rf)
        # forward with exact scalar minimisation
        s=np.zeros(B)
        newcode=np.zeros((B,N),dtype=int)
        wn=np.zeros((B,N))
        for k in range(N):
            Q=Qs[:,k]; rp=rps[:,k]
            inv=1.0/(d[:,k]+Q)
            num=Q*s+rp+gh[:,k]
            wk=np.full(B,np.nan); ck=np.full(B,-1)
            for j in range(nseg):
                cand=-(num+slope[j])*inv
                ok=(cand>brk[j])&(cand<brk[j+1])&(ck<0)
                wk=np.where(ok,cand,wk); ck=np.where(ok,2*j+1,ck)
            for i in range(nseg+1):
                gi=num+(d[:,k]+Q)*brk[i]   # smooth-part derivative at brk[i]
                slo_=slope[i-1] if i>0 else -np.inf
                shi_=slope[i] if i<nseg else np.inf
                ok=(-gi>=slo_)&(-gi<=shi_)&(ck<0)
                wk=np.where(ok,brk[i],wk); ck=np.where(ok,2*i,ck)
            assert (ck>=0).all()
            wn[:,k]=wk; newcode[:,k]=ck; s=s+wk
        conv=(newcode==code).all(axis=1)
        newly=conv&~done; iters[newly]=it; done|=conv
        w=np.where(done[:,None]&~newly[:,None],w,wn)
        if done.all(): break
        code=np.where(done[:,None],code,newcode)
        fixed=(code%2==0); fval=brk[np.minimum(code//2,nseg)]; seg=np.minimum(code//2,nseg-1)
    iters[~done]=max_it
    return w,iters,done

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=1000
    for consts in (small_ev_consts(),large_ev_consts()):
       for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            w,iters,done=ddp(N,consts,lm,lr,gam)
            err=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                if done[b]: err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d  fail %d err %.2e'%(iters[done].mean() if done.any() else -1,np.percentile(iters,99),iters.max(),(~done).sum(),err))
