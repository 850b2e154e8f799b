import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/incentive-design-mpc_b200'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, torch
from run_fleet import fleet_consts, fleet_demand
from chargingstation.fleet import ChargingStationFleet
S=1024
consts = fleet_consts(4, 24, 24, 500, 12)
demand = fleet_demand(consts, S, 4, 24)
fleet = ChargingStationFleet(consts, S, demand=demand, seed=4, rng="device", chain="partition")
for t in range(4):
    fleet.step()
    st = fleet.bi["status"].cpu().numpy(); it = fleet.bi["iters"].cpu().numpy()
    bad = np.nonzero(st)[0]
    print(t, "failed", bad[:8], st[bad][:8], it[bad][:8])
    if len(bad):
        s = bad[0]
        np.savez('/root/repo/gpurun_out/bimpc_fail.npz', **{k: v[s].cpu().numpy() for k, v in fleet.bi.items()})
        break
