"""Prototype of the kernel algorithm: per iteration ONE backward sweep (gradient + config + Riccati)
and ONE forward sweep (stage-optimal rollout + cost); safeguarded by a projected-Newton fallback."""
import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data

def hybrid(N,consts,lm,lr,gam,max_it=200,tol=1e-11,use_pn_config=True,stats=None):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    tq=tol*np.maximum(1.0,np.abs(gh).max(axis=1)+c*N*consts.y_max)
    w=np.zeros((B,N)); sN=np.zeros(B)          # s_{N-1}
    f=np.full(B,np.inf)
    f=0.5*c*N*gam**2+0*gam    # cost at w=0 (without kappa0): (c/2) sum (0-gam)^2
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int); nfb=np.zeros(B,dtype=int)
    kkt=np.zeros(B)
    Qa=np.zeros((B,N)); ra=np.zeros((B,N)); inva=np.zeros((B,N)); bind=np.zeros((B,N),dtype=bool); sga=np.zeros((B,N),dtype=int)
    for it in range(max_it):
        # ---- backward sweep: costate, gradient, config, Riccati
        P=np.zeros(B); r=np.zeros(B); p=np.zeros(B); s=sN.copy(); viol=np.zeros(B)
        for k in range(N-1,-1,-1):
            wk=w[:,k]
            p=p+c*(s-gam)
            q=d[:,k]*wk+gh[:,k]+p
            # breakpoint index / containing segment
            atb=np.full(B,-1)
            for i in range(nseg+1): atb=np.where(wk==brk[i],i,atb)
            segin=np.clip((wk[:,None]>=brk[None,1:-1]).sum(-1),0,nseg-1)
            i=np.maximum(atb,0)
            go_r=(atb>=0)&(-q>shi[i]+tq)
            go_l=(atb>=0)&(-q<slo[i]-tq)
            binding=(atb>=0)&~go_r&~go_l
            seg=np.where(atb<0,segin,np.where(go_r,np.minimum(i,nseg-1),np.maximum(i-1,0)))
            if not use_pn_config:
                binding=(atb>=0)
            v=np.where(atb<0,np.abs(q+slope[seg]),np.where(go_r,-q-shi[i],np.where(go_l,slo[i]+q,0.0)))
            v=np.where(np.isfinite(v),v,0.0)
            viol=np.maximum(viol,v/ (tq/tol))
            Q=c+P; rp=r-c*gam
            inv=1.0/(d[:,k]+Q)
            h=gh[:,k]+slope[seg]
            Pf=Q*d[:,k]*inv; rf=(d[:,k]*rp-Q*h)*inv
            Px=Q; rx=Q*wk+rp
            P=np.where(binding,Px,Pf); r=np.where(binding,rx,rf)
            Qa[:,k]=Q; ra[:,k]=rp; inva[:,k]=inv; bind[:,k]=binding; sga[:,k]=seg
            s=s-wk
        conv=viol<=tol
        newly=conv&~done; iters[newly]=it; kkt[newly]=viol[newly]; done|=conv
        if done.all(): break
        # ---- forward sweep: stage-optimal rollout + cost
        s=np.zeros(B); fn=np.zeros(B); wn=np.zeros((B,N))
        for k in range(N):
            num=Qa[:,k]*s+ra[:,k]+gh[:,k]
            x=-(num+slope[nseg-1])*inva[:,k]
            for j in range(nseg-2,-1,-1):
                x=np.minimum(-(num+slope[j])*inva[:,k],np.maximum(brk[j+1],x))
            x=np.minimum(np.maximum(x,0.0),consts.w_max)
            wn[:,k]=x; s=s+x
            psi=0
            for j in range(1,nseg): psi=psi+(slope[j]-slope[j-1])*np.maximum(x-brk[j],0)
            fn=fn+0.5*d[:,k]*x*x+gh[:,k]*x+psi+0.5*c*(s-gam)**2
        acc=(~done)&(fn<f)
        rej=(~done)&~acc
        # ---- fallback: projected Newton line search for rejected ones
        if rej.any():
            nfb[rej]+=1
            s=np.zeros(B); wt=np.zeros((B,N))
            for k in range(N):
                num=Qa[:,k]*s+ra[:,k]+gh[:,k]
                x=np.where(bind[:,k],w[:,k],-(num+slope[sga[:,k]])*inva[:,k])
                wt[:,k]=x; s=s+x
            lo=brk[sga]; hi=brk[sga+1]
            alpha=np.ones(B); act=rej.copy()
            for ls in range(50):
                cand=np.where(bind,w,np.clip(w+alpha[:,None]*(wt-w),lo,hi))
                S=np.cumsum(cand,axis=1)
                fc=0.5*np.sum(d*cand*cand,axis=1)+np.sum(gh*cand,axis=1)+0.5*c*np.sum((S-gam[:,None])**2,axis=1)
                for j in range(1,nseg): fc=fc+(slope[j]-slope[j-1])*np.sum(np.maximum(cand-brk[j],0),axis=1)
                ok=act&(fc<f)
                wn=np.where(ok[:,None],cand,wn); fn=np.where(ok,fc,fn); acc|=ok
                act&=~ok
                if not act.any(): break
                alpha=np.where(act,alpha*0.5,alpha)
            # stuck ones: declare done with current w
            if act.any():
                iters[act]=-it; done|=act
        w=np.where(acc[:,None],wn,w); f=np.where(acc,fn,f); sN=w.sum(axis=1)
    iters[~done]=max_it
    return w,f,iters,done,nfb,kkt

if __name__=='__main__':
    rng=np.random.default_rng(1)
    B=2000
    for pnc in (True,False):
     print('use_pn_config',pnc)
     for consts in (small_ev_consts(),large_ev_consts()):
       for N in (12,24,48,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            w,f,iters,done,nfb,kkt=hybrid(N,consts,lm,lr,gam,use_pn_config=pnc)
            err=0;cerr=0
            for b in range(0,B,200):
                wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
                err=max(err,np.max(np.abs(w[b]-wo))/consts.w_max)
                k0=consts.theta*consts.w_max*lm[b,N:2*N].sum()
                cerr=max(cerr,abs(f[b]-0.5*2*consts.delta*consts.theta**2*N*gam[b]**2+k0-co)/max(1,abs(co)))
            print(' ',consts.ev_type,N,'mode',mode,'iters mean %.2f p99 %d max %d stuck %d fail %d fallbacks %.3f err %.2e cerr %.1e'%(iters[done].mean(),np.percentile(iters,99),iters.max(),(iters<0).sum(),(~done).sum(),nfb.mean(),err,cerr))
