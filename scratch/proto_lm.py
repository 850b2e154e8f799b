"""Variant: rejected rollouts are retried with a proximal (Levenberg-Marquardt) term mu instead of a
projected-Newton line search -> one uniform code path (backward + forward)."""
import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments
from proto_pdas import make_batch,data

def lm_solve(N,consts,lm,lr,gam,max_it=400,tol=1e-11,mu0_scale=1.0,grow=4.0,shrink=0.25):
    d,c,gh=data(N,consts,lm,lr,gam)
    B=d.shape[0]
    brk,slope=_segments(consts); nseg=len(slope)
    slo=np.concatenate([[-np.inf],slope]); shi=np.concatenate([slope,[np.inf]])
    gscale=np.maximum(1.0,np.abs(gh).max(axis=1)+c*N*consts.y_max); tq=tol*gscale
    band=1e-9*consts.w_max
    ftol=1e-15*(c*N*consts.y_max**2+N*consts.w_max*(np.abs(gh).max(axis=1)+0.5*d.max(axis=1)*consts.w_max+slope[-1]))
    w=np.zeros((B,N)); sN=np.zeros(B); f=0.5*c*N*gam**2
    mu=np.zeros(B); mu0=mu0_scale*c*np.ones(B)
    done=np.zeros(B,dtype=bool); iters=np.zeros(B,dtype=int); nrej=np.zeros(B,dtype=int)
    Qa=np.zeros((B,N)); ra=np.zeros((B,N)); inva=np.zeros((B,N))
    for it in range(max_it):
        P=np.zeros(B); r=np.zeros(B); p=np.zeros(B); s=sN.copy(); viol=np.zeros(B)
        for k in range(N-1,-1,-1):
            wk=w[:,k]
            p=p+c*(s-gam)
            q=d[:,k]*wk+gh[:,k]+p
            atb=np.full(B,-1)
            for i in range(nseg+1): atb=np.where(np.abs(wk-brk[i])<=band,i,atb)
            segin=np.clip((wk[:,None]>brk[None,1:-1]+band).sum(-1),0,nseg-1)
            i=np.maximum(atb,0)
            go_r=(atb>=0)&(-q>shi[i]+tq); go_l=(atb>=0)&(-q<slo[i]-tq)
            binding=(atb>=0)&~go_r&~go_l
            seg=np.where(atb<0,segin,np.where(go_r,np.minimum(i,nseg-1),np.maximum(i-1,0)))
            v=np.where(atb<0,np.abs(q+slope[seg]),np.where(go_r,-q-shi[i],np.where(go_l,slo[i]+q,0.0)))
            v=np.where(np.isfinite(v),v,0.0); viol=np.maximum(viol,v)
            # proximal term: d -> d+mu, g -> g - mu*w_old
            dk=d[:,k]+mu; gk=gh[:,k]-mu*wk
            Q=c+P; rp=r-c*gam; inv=1.0/(dk+Q); h=gk+slope[seg]
            Pf=Q*dk*inv; rf=(dk*rp-Q*h)*inv; Px=Q; rx=Q*wk+rp
            P=np.where(binding,Px,Pf); r=np.where(binding,rx,rf)
            Qa[:,k]=Q; ra[:,k]=rp; inva[:,k]=inv
            s=s-wk
        conv=viol<=tq
        newly=conv&~done; iters[newly]=it; done|=conv
        if done.all(): break
        s=np.zeros(B); fn=np.zeros(B); wn=np.zeros((B,N))
        for k in range(N):
            gk=gh[:,k]-mu*w[:,k]
            num=Qa[:,k]*s+ra[:,k]+gk
            x=-(num+slope[nseg-1])*inva[:,k]
            for j in range(nseg-2,-1,-1): x=np.minimum(-(num+slope[j])*inva[:,k],np.maximum(brk[j+1],x))
            x=np.minimum(np.maximum(x,0.0),consts.w_max)
            wn[:,k]=x; s=s+x
            psi=0
            for j in range(1,nseg): psi=psi+(slope[j]-slope[j-1])*np.maximum(x-brk[j],0)
            fn=fn+0.5*d[:,k]*x*x+gh[:,k]*x+psi+0.5*c*(s-gam)**2
        acc=(~done)&(fn<=f+ftol)
        rej=(~done)&~acc
        nrej[rej]+=1
        w=np.where(acc[:,None],wn,w); f=np.where(acc,np.minimum(f,fn),f); sN=w.sum(axis=1)
        mu=np.where(acc,np.where(mu*shrink<mu0*0.5,0.0,mu*shrink),np.maximum(mu0,mu*grow))
    iters[~done]=max_it
    return w,iters,done,nrej

if __name__=='__main__':
    from proto_hybrid import hybrid
    rng=np.random.default_rng(1)
    B=2000
    for consts in (small_ev_consts(),large_ev_consts()):
      for N in (24,96):
        for mode in (0,1,2,3):
            lm,lr,gam=make_batch(rng,N,consts,B,mode)
            if mode==1: lm=lm*(rng.random(lm.shape)<0.5)
            w0,f0,it0,dn0,nfb0,_=hybrid(N,consts,lm,lr,gam)
            res=[]
            for mu0s,grow,shrink in ((1.0,4.0,0.25),(4.0,4.0,0.0),(16.0,4.0,0.0),(0.25,8.0,0.0)):
                w,iters,done,nrej=lm_solve(N,consts,lm,lr,gam,mu0_scale=mu0s,grow=grow,shrink=shrink)
                err=np.abs(w-w0).max()/consts.w_max
                res.append('mu0=%g g=%g s=%g: it %.2f p99 %d max %d fail %d rej %.2f err %.0e'%(mu0s,grow,shrink,iters[done].mean(),np.percentile(iters,99),iters.max(),(~done).sum(),nrej.mean(),err))
            print(consts.ev_type,N,'mode',mode,'| PN: it %.2f p99 %d max %d fb %.2f'%(it0.mean(),np.percentile(it0,99),it0.max(),nfb0.mean()))
            for r_ in res: print('      ',r_)
