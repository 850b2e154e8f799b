import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/incentive-design-mpc_b200')
from oracle import lompc_oracle as orc, price_oracle as po
from chargingstation.lompc import LoMPCConstants
from chargingstation.price_solver import PriceSolver
import torch
o=orc.small_ev_consts(); c=LoMPCConstants(o.delta,o.theta,o.y_max,o.w_max,o.ev_type)
N=12
for pt in ("linear","linear-convex"):
    ps=PriceSolver(N,c,pt); ora=po.PriceOracle(N,o,pt)
    rng=np.random.default_rng(21)
    w=o.w_max*rng.random(N); w_ref=o.w_max*rng.random(N); lam=0.05*o.theta*rng.random(ps.r)
    A_bar,A_bar_inv=ora._metric(0.0)
    P,q=ora.price_step_matrices(A_bar_inv,w_ref,w,lam); x=po.nnqp_exact(P,q)
    try:
        ln,dec=ps._price_gradient_descent_step(A_bar_inv,w_ref,w,lam,lmbd_r=0.0)
        torch.cuda.synchronize()
        print(pt,'err',np.abs(ln-x).max(),'dec',dec,(lam@P@lam+q@lam)-(x@P@x+q@x))
        print(' in ',lam[:6]); print(' out',ln[:6]); print(' orc',x[:6])
    except Exception as e:
        print('EXC',e)
