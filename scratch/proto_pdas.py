"""Prototype (numpy, batched) of the candidate GPU algorithm: block primal-dual
active set with O(N) scalar-Riccati equality solves.  Scratch, not shipped."""
import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments

def make_batch(rng, N, consts, B, mode):
    th=consts.theta
    if mode==0:
        lm=th*rng.random((B,3*N)); lr=3*N*consts.delta*rng.random(B); gam=consts.y_max*rng.random(B)
    elif mode==1:
        lm=0.05*th*rng.random((B,3*N)); lr=np.zeros(B); gam=consts.y_max-(0.3+0.2*rng.random(B))
    elif mode==2:
        lm=np.zeros((B,3*N)); lm[:,:2*N]=0.05*th*rng.random((B,2*N)); lr=np.zeros(B); gam=consts.y_max*rng.random(B)
    else:
        lm=np.zeros((B,3*N)); lr=np.zeros(B); gam=consts.y_max*rng.random(B)
    return lm,lr,gam

def data(N,consts,lm,lr,gam):
    th,wm,dl=consts.theta,consts.w_max,consts.delta
    q=3*th/(4*wm); c=2*dl*th**2
    d=2*(lr[:,None]*th**2+q*lm[:,2*N:])
    if consts.ev_type=='small': d=d+2*th**2/0.81
    gh=th*(lm[:,:N]-lm[:,N:2*N])   # g without the gamma term
    return d,c,gh

def riccati(N,d,c,h,gam,fixed,fval):
    """min sum 1/2 d w^2 + h w + c/2 sum (s-gam)^2, w_k=fval where fixed."""
    B=d.shape[0]
    P=np.zeros(B); r=np.zeros(B)
    K=np.zeros((B,N)); kap=np.zeros((B,N))
    for k in range(N-1,-1,-1):
        Q=c+P; rp=r-c*gam
        inv=1.0/(d[:,k]+Q)
        K[:,k]=Q*inv; kap[:,k]=(h[:,k]+rp)*inv
        Pf=Q*d[:,k]*inv; rf=(d[:,k]*rp-Q*h[:,k])*inv
        Px=Q; rx=Q*fval[:,k]+rp
        fx=fixed[:,k]
        P=np.where(fx,Px,Pf); r=np.where(fx,rx,rf)
    w=np.zeros((B,N)); s=np.zeros(B); S=np.zeros((B,N))
    for k in range(N):
        wk=np.where(fixed[:,k],fval[:,k],-K[:,k]*s-kap[:,k])
        w[:,k]=wk; s=s+wk; S[:,k]=s
    # smooth gradient q = d w + gh + c * revcumsum(s-gam)
    return w,S

def smooth_grad(N,d,c,gh,gam,w,S):
    t=np.cumsum((S-gam[:,None])[:,::-1],axis=1)[:,::-1]
    return d*w+gh+c*t

def pdas(N,consts,lm,lr,gam,max_it=50,init='zero'):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts)
    nseg=len(slope)
    Hkk=d+c*np.arange(N,0,-1)[None,:]
    # config: code 2*i = fixed at brk[i]; 2*j+1 = free in seg j
    w=np.zeros((B,N)); S=np.zeros((B,N))
    q=smooth_grad(N,d,c,gh,gam,w,S)
    code=np.zeros((B,N),dtype=int)
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int)
    hist=[]
    for it in range(max_it):
        # scalar prox: minimise 1/2 Hkk (x-w)^2 + q (x-w) + psi(x) on [0,wm]
        # candidates: for each seg j: x_j = w - (q+slope_j)/Hkk, valid if in seg; else breakpoints
        newcode=np.full((B,N),-1)
        for j in range(nseg):
            xj=w-(q+slope[j])/Hkk
            ins=(xj>brk[j])&(xj<brk[j+1])&(newcode<0)
            newcode=np.where(ins,2*j+1,newcode)
        for i in range(nseg+1):
            slo=slope[i-1] if i>0 else -np.inf
            shi=slope[i] if i<nseg else np.inf
            gi=q+Hkk*(brk[i]-w)   # gradient of the scalar model at brk[i]
            ok=(-gi>=slo)&(-gi<=shi)&(newcode<0)
            newcode=np.where(ok,2*i,newcode)
        assert (newcode>=0).all()
        if it>0:
            conv=(newcode==code).all(axis=1)
            newly=conv&~done
            iters[newly]=it
            done|=conv
            if done.all(): break
        code=np.where(done[:,None],code,newcode)
        fixed=(code%2==0); fval=brk[np.minimum(code//2,nseg)]
        h=gh+np.where(fixed,0.0,slope[np.minimum(code//2,nseg-1)])
        w,S=riccati(N,d,c,h,gam,fixed,fval)
        q=smooth_grad(N,d,c,gh,gam,w,S)
    iters[~done]=max_it
    return w,iters,done

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=2000
    for consts in (small_ev_consts(),large_ev_consts()):
      for N in (12,24,48,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            t0=time.time()
            w,iters,done=pdas(N,consts,lm,lr,gam)
            # check vs oracle on a subset
            err=0;kv=0
            for b in range(0,B,100):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                if done[b]: err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
            print(consts.ev_type,N,'mode',mode,'iters mean %.2f max %d  fail %d  err %.2e'%(iters[done].mean(),iters.max(),(~done).sum(),err))
