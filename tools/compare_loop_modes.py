"""Teacher-forced comparison of the two device-resident price loops inside the fleet closed loop: the state of the
thread-per-EV fleet (loop mode 3) is copied into the parametric one (loop mode 2) before every step, and the
iteration counts / prices of every group are compared.   python tools/compare_loop_modes.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'incentive-design-mpc_b200'), os.path.join(ROOT, 'tools')):
    sys.path.insert(0, p)
import numpy as np, torch
from run_fleet import fleet_consts, fleet_demand
from chargingstation.fleet import ChargingStationFleet
S, T = int(os.environ.get('CMP_S', '256')), int(os.environ.get('CMP_T', '6'))
dump = None
consts = fleet_consts(T, 24, 24, 500, 12)
demand = fleet_demand(consts, S, T, 24)
fl = {}
for m in (3, 2):
    f = ChargingStationFleet(consts, S, demand=demand, seed=4, rng="device", chain="reference")
    for k in ("s","l"): f.solver[k].set_loop_mode(m)
    fl[m] = f
for t in range(T):
    # teacher-force: copy state of mode-3 fleet into mode-2 fleet before each step
    a, b = fl[3], fl[2]
    for k in ("s","l"):
        b.y[k].copy_(a.y[k]); b.prev[k].copy_(a.prev[k]); b.ncharged[k].copy_(a.ncharged[k])
    b.x.copy_(a.x)
    a.step(); b.step(); torch.cuda.synchronize()
    for k in ("s","l"):
        ia, ib = a.w[k]["iters"].cpu().numpy(), b.w[k]["iters"].cpu().numpy()
        pa, pb = a.prices[k].cpu().numpy(), b.prices[k].cpu().numpy()
        nd = int((ia != ib).sum())
        small = np.abs(pa) <= 1e3
        rel = np.max(np.abs(pa-pb)[small]) if small.any() else 0
        if nd and dump is None:
            g = int(np.nonzero(ia != ib)[0][0])
            off = a.w[k]["off"].cpu().numpy()
            dump = dict(ev=k, g=g, step=t, y0=a.w[k]["ysort"].cpu().numpy()[off[g]:off[g + 1]], w_ref=a.w[k]["w_ref"].cpu().numpy()[g],
                        iters_thread=ia[g], iters_param=ib[g], prices_thread=pa[g], prices_param=pb[g],
                        all_off=off, S=S, P=12)
        print(f"step {t} {k}: groups {ia.size} iter mismatches {nd} (capped a {int((ia>=999).sum())} b {int((ib>=999).sum())}) max |dprice| (entries<=1e3) {rel:.2e}", [ (int(x),int(y)) for x,y in zip(ia[ia!=ib][:6], ib[ia!=ib][:6])])

if dump is not None:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "loop_mode_mismatch.npz"), **dump)
    print("first mismatch dumped:", dump["ev"], dump["g"], dump["step"], dump["iters_thread"], dump["iters_param"], len(dump["y0"]))
