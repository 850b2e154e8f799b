import numpy as np, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from oracle import lompc_oracle as orc
from oracle import price_oracle as po_mod
from oracle.price_oracle import PriceOracle, nnqp_exact
from proto_nnqp import nnqp_pdas
rng=np.random.default_rng(1)
stats={'n':0,'fail':0,'its':[],'err':0}
class PO(PriceOracle):
    def price_gradient_descent_step(self, A_bar_inv, w_ref, w, lmbd):
        P,q=self.price_step_matrices(A_bar_inv,w_ref,w,lmbd)
        x=nnqp_exact(P,q)
        qs=3*self.consts.theta/(4*self.consts.w_max)
        lam,it,st=nnqp_pdas(self.N,self.r,self.consts.theta,qs,self.m,self.eps_reg,self._lr/self.consts.delta,w,w_ref,lmbd,self.consts.w_max)
        stats['n']+=1; stats['fail']+=st; stats['its'].append(it)
        if not st: stats['err']=max(stats['err'],np.abs(lam-x).max()/max(1,np.abs(x).max()))
        dual_cost=lmbd@P@lmbd+q@lmbd
        return x, dual_cost-(x@P@x+q@x)
for trial in range(12):
    consts=orc.small_ev_consts() if trial%2 else orc.large_ev_consts()
    pt='linear' if (trial//2)%2 else 'linear-convex'
    N=12; lr=[0.0,0.0,0.0,float(N)][(trial//4)%4]
    p=PO(N,consts,pt); p._lr=lr
    y0=(1/36.)*consts.y_max*rng.random(5) if trial%3 else 0.3+0.05*rng.random(8)
    p.set_charge_levels(y0)
    w_ref=consts.w_max*rng.random(N)
    lm,st=p.compute_optimal_prices(w_ref,lr,max_iter=150)
    print(trial,consts.ev_type,pt,'lr',lr,'iters',st['iter'],'nnqp calls',stats['n'],'fails',stats['fail'],'pdas its max',max(stats['its']) if stats['its'] else 0,'err %.1e'%stats['err'],flush=True)
