"""Scalar Python replica of lompc_solve_kernel (for debugging)."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle.lompc_oracle import *
from oracle.lompc_oracle import _segments

def solve(N,consts,lm,lr,gam,max_iter=200,tol=1e-11,trace=False,ftol_mode=0,band_rel=0.0):
    th,wm,dl=consts.theta,consts.w_max,consts.delta
    qs=3*th/(4*wm); c=2*dl*th**2
    brk,slope=_segments(consts); nseg=len(slope)
    D=2*(lr*th**2+qs*lm[2*N:])+(2*th**2/0.81 if consts.ev_type=='small' else 0)
    G=th*(lm[:N]-lm[N:2*N])
    gscale=max(1.0,np.abs(G).max()+c*N*consts.y_max); tq=tol*gscale; cg=c*gam
    band=band_rel*wm
    W=np.zeros(N); WN=np.zeros(N); sN=0.0; f=0.5*c*N*gam*gam
    KK=np.zeros(N);KAP=np.zeros(N);INV=np.zeros(N);BIND=np.zeros(N,bool);SEG=np.zeros(N,int)
    def psi(x): return sum((slope[j]-slope[j-1])*max(x-brk[j],0.0) for j in range(1,nseg))
    fs=c*N*consts.y_max**2+N*wm*(np.abs(G).max()+0.5*D.max()*wm+slope[-1])
    for it in range(max_iter):
        P=0.0;r=0.0;p=0.0;s=sN;viol=0.0
        for k in range(N-1,-1,-1):
            wk=W[k];dk=D[k];gk=G[k]
            p=c*s+p-cg
            q=dk*wk+gk+p
            at=-1
            for i in range(nseg+1):
                if abs(wk-brk[i])<=band: at=i
            binding=False;seg=0
            if at>=0:
                mq=-q
                if at<nseg and mq>slope[at]+tq: seg=at;v=mq-slope[seg]
                elif at>0 and mq<slope[at-1]-tq: seg=at-1;v=slope[seg]-mq
                else: binding=True;seg=min(at,nseg-1);v=0.0
            else:
                seg=sum(1 for j in range(1,nseg) if wk>brk[j]); v=abs(q+slope[seg])
            viol=max(viol,v)
            Q=c+P;rp=r-cg;inv=1.0/(dk+Q);h=gk+slope[seg]
            if binding: P=Q;r=Q*wk+rp
            else: P=Q*dk*inv;r=(dk*rp-Q*h)*inv
            KK[k]=Q*inv;KAP[k]=(rp+gk)*inv;INV[k]=inv;BIND[k]=binding;SEG[k]=seg
            s-=wk
        if trace: print('it',it,'viol/gs %.3e'%(viol/gscale),'f',repr(f))
        if viol<=tq: return W,it,0,viol/gscale
        fn=0.0;s=0.0
        for k in range(N):
            x0=-(KK[k]*s+KAP[k])
            x=x0-slope[nseg-1]*INV[k]
            for j in range(nseg-2,-1,-1): x=min(x0-slope[j]*INV[k],max(brk[j+1],x))
            x=min(max(x,0.0),wm)
            WN[k]=x;s+=x;e=s-gam
            fn+=x*(0.5*D[k]*x+G[k])+0.5*c*e*e+psi(x)
        if trace: print('   rollout fn',repr(fn),'df',fn-f)
        acc = fn<f if ftol_mode==0 else fn<=f+1e-15*fs
        if acc:
            W,WN=WN,W;f=min(fn,f) if ftol_mode else fn;sN=s;continue
        s=0.0
        for k in range(N):
            if BIND[k]: x=W[k]
            else: x=-(KK[k]*s+KAP[k])-slope[SEG[k]]*INV[k]
            WN[k]=x;s+=x
        alpha=1.0;ok=False
        for ls in range(60):
            fn=0.0;s=0.0
            for k in range(N):
                x=W[k]
                if not BIND[k]:
                    x=W[k]+alpha*(WN[k]-W[k]); x=min(max(x,brk[SEG[k]]),brk[SEG[k]+1])
                s+=x;e=s-gam
                fn+=x*(0.5*D[k]*x+G[k])+0.5*c*e*e+psi(x)
            ok=fn<f
            if ok: break
            alpha*=0.5
        if trace: print('   fallback alpha',alpha,'ok',ok,'fn-f',fn-f)
        if not ok: return W,it,1,viol/gscale
        s=0.0
        for k in range(N):
            if not BIND[k]:
                x=W[k]+alpha*(WN[k]-W[k]); x=min(max(x,brk[SEG[k]]),brk[SEG[k]+1]); W[k]=x
            s+=W[k]
        f=fn;sN=s
    return W,max_iter,1,viol/gscale

if __name__=='__main__':
    z=np.load('/root/repo/gpurun_out/failures.npz')
    keys=sorted(set('_'.join(k.split('_')[:3]) for k in z.files))
    for key in keys:
        ev,Ns,m=key.split('_'); N=int(Ns[1:]); consts=small_ev_consts() if ev=='small' else large_ev_consts()
        lm=z[key+'_lmbd'];lr=z[key+'_lmbd_r'];gam=z[key+'_gamma']
        for b in range(len(gam)):
            W,it,st,kk=solve(N,consts,lm[b],lr[b],gam[b])
            W2,it2,st2,kk2=solve(N,consts,lm[b],lr[b],gam[b],ftol_mode=1,band_rel=1e-9)
            wo,co,_=solve_active_set(N,consts,lm[b],lr[b],gam[b])
            print(key,b,'gpu it',z[key+'_iters'][b],'kkt %.1e'%z[key+'_kkt'][b],'| replica it',it,'st',st,'kkt %.1e'%kk,'err %.1e'%(np.abs(W-wo).max()/consts.w_max),'| ftol: it',it2,'st',st2,'kkt %.1e'%kk2,'err %.1e'%(np.abs(W2-wo).max()/consts.w_max))
