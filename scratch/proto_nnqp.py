"""Prototype of K3: exact NNQP price step with O(N) structured Newton steps (scalar python)."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import lompc_oracle as orc
from oracle.price_oracle import PriceOracle, nnqp_exact, nnqp_kkt

def ric_solve(d,c,b):
    """x = (diag(d)+c A'A)^-1 b"""
    N=len(d); K=np.zeros(N); kap=np.zeros(N); P=0.0; r=0.0
    for k in range(N-1,-1,-1):
        Q=c+P; inv=1.0/(d[k]+Q); g=-b[k]
        K[k]=Q*inv; kap[k]=(r+g)*inv
        P,r=Q*d[k]*inv,(d[k]*r-Q*g)*inv
    x=np.zeros(N); s=0.0
    for k in range(N):
        x[k]=-(K[k]*s+kap[k]); s+=x[k]
    return x

def nnqp_pn(N,r,theta,qs,m,eps,kappa,w,w_ref,lam_k,tol=1e-12,max_it=100,trace=False):
    nb=r//N
    coef=np.zeros((nb,N)); coef[0]=theta; coef[1]=-theta
    if nb==3: coef[2]=2*qs*w
    dk=np.full(N,kappa)
    def Bt(l): return (coef*l.reshape(nb,N)).sum(0)           # u = B' lambda
    def Pmul(l): 
        v=ric_solve(dk,1.0,Bt(l)); return eps*l+(coef*v[None,:]).reshape(-1)/(2*m), v
    phi=lambda x: np.concatenate([theta*x, theta*(orc_wmax-x), qs*x*x])[:r]
    Plk,_=Pmul(lam_k)
    rho=Plk+0.5*(phi(w)-phi(w_ref))
    def cost(l):
        u=Bt(l); v=ric_solve(dk,1.0,u); return eps*l@l+(u@v)/(2*m)-2*rho@l
    lam=lam_k.copy(); F=cost(lam)
    gs=max(1.0,np.abs(rho).max())
    for it in range(max_it):
        Pl,_=Pmul(lam); g=Pl-rho           # half gradient
        binding=(lam<=0)&(g>=-tol*gs)
        viol=np.max(np.where(lam<=0,np.maximum(-g,0),np.abs(g)))
        if trace: print(it,'viol',viol/gs,'F',F)
        if viol<=tol*gs: return lam,it,0
        free=~binding
        c2=(coef.reshape(-1))
        t=((c2**2)*free).reshape(nb,N).sum(0)
        rhs=((c2*rho)*free).reshape(nb,N).sum(0)
        z=ric_solve(t+2*m*eps*kappa,2*m*eps,rhs)
        lt=np.where(free,(rho-c2*np.tile(z,nb))/eps,0.0)
        alpha=1.0; ok=False
        for ls in range(60):
            cand=np.where(free,np.maximum(lam+alpha*(lt-lam),0.0),lam)
            Fc=cost(cand)
            if Fc<=F+1e-15*abs(F) if ls==0 else Fc<F: ok=True;break
            alpha*=0.5
        if not ok: return lam,it,1
        lam=cand;F=min(F,Fc)
    return lam,max_it,1

if __name__=='__main__':
    rng=np.random.default_rng(0)
    worst=0; its=[]
    for trial in range(200):
        consts=orc.small_ev_consts() if trial%2 else orc.large_ev_consts()
        pt='linear' if (trial//2)%2 else 'linear-convex'
        N=[12,24][(trial//4)%2]
        lmbd_r=[0.0,0.0,N*1.0,3.0*N][trial%4] if trial%8<4 else 0.0
        po=PriceOracle(N,consts,pt)
        orc_wmax=consts.w_max
        w=consts.w_max*rng.random(N)*(rng.random(N)<0.8); w_ref=consts.w_max*rng.random(N)
        lam_k=consts.theta*0.05*rng.random(po.r)*(rng.random(po.r)<0.6)
        A_bar,A_bar_inv=po._metric(lmbd_r)
        P,q=po.price_step_matrices(A_bar_inv,w_ref,w,lam_k)
        x=nnqp_exact(P,q)
        qs=3*consts.theta/(4*consts.w_max)
        lam,it,st=nnqp_pn(N,po.r,consts.theta,qs,po.m,po.eps_reg,lmbd_r/consts.delta,w,w_ref,lam_k)
        err=np.abs(lam-x).max()/max(1,np.abs(x).max())
        worst=max(worst,err); its.append(it)
        if st or err>1e-8: print('trial',trial,'st',st,'it',it,'err',err,'kkt nnls',nnqp_kkt(P,q,x),'kkt pn',nnqp_kkt(P,q,lam))
    print('worst err',worst,'iters mean',np.mean(its),'max',np.max(its))

def nnqp_pdas(N,r,theta,qs,m,eps,kappa,w,w_ref,lam_k,wmax,tol=1e-12,max_it=100):
    nb=r//N
    coef=np.zeros((nb,N)); coef[0]=theta; coef[1]=-theta
    if nb==3: coef[2]=2*qs*w
    c2=coef.reshape(-1)
    dk=np.full(N,kappa)
    def Bt(l): return (coef*l.reshape(nb,N)).sum(0)
    def Pmul(l):
        v=ric_solve(dk,1.0,Bt(l)); return eps*l+(coef*v[None,:]).reshape(-1)/(2*m)
    phi=lambda x: np.concatenate([theta*x, theta*(wmax-x), qs*x*x])[:r]
    rho=Pmul(lam_k)+0.5*(phi(w)-phi(w_ref))
    gs=max(1.0,np.abs(rho).max())
    free=(lam_k>0)|(Pmul(lam_k)-rho<0)
    for it in range(max_it):
        t=((c2**2)*free).reshape(nb,N).sum(0)
        rhs=((c2*rho)*free).reshape(nb,N).sum(0)
        z=ric_solve(t+2*m*eps*kappa,2*m*eps,rhs)
        lam=np.where(free,(rho-c2*np.tile(z,nb))/eps,0.0)
        g=Pmul(lam)-rho
        newfree=np.where(free,lam>0,g<-tol*gs)
        if (newfree==free).all(): return lam,it+1,0
        free=newfree
    return lam,max_it,1

if __name__=='__main__':
    rng=np.random.default_rng(0)
    worst=0; its=[]; fails=0
    for trial in range(400):
        consts=orc.small_ev_consts() if trial%2 else orc.large_ev_consts()
        pt='linear' if (trial//2)%2 else 'linear-convex'
        N=[12,24][(trial//4)%2]
        lmbd_r=[0.0,0.0,N*1.0,3.0*N][trial%4] if trial%8<4 else 0.0
        po=PriceOracle(N,consts,pt)
        w=consts.w_max*rng.random(N)*(rng.random(N)<0.8); w_ref=consts.w_max*rng.random(N)
        lam_k=consts.theta*0.05*rng.random(po.r)*(rng.random(po.r)<0.6)
        A_bar,A_bar_inv=po._metric(lmbd_r)
        P,q=po.price_step_matrices(A_bar_inv,w_ref,w,lam_k)
        x=nnqp_exact(P,q)
        qs=3*consts.theta/(4*consts.w_max)
        lam,it,st=nnqp_pdas(N,po.r,consts.theta,qs,po.m,po.eps_reg,lmbd_r/consts.delta,w,w_ref,lam_k,consts.w_max)
        err=np.abs(lam-x).max()/max(1,np.abs(x).max())
        if st: fails+=1
        else: worst=max(worst,err); its.append(it)
    print('PDAS: fails',fails,'worst err',worst,'iters mean',np.mean(its),'max',np.max(its))
