#!/usr/bin/env python
"""Robustness sweep of the closed loop: several scenario shapes / seeds, asserting that no LoMPC solve fails
inside the fused price loop (price_solve_*_dev would return LOMPC_ERR_NOT_CONVERGED), that every BiMPC
converges and that no NaN appears.  Prints one line per scenario."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incentive-design-mpc_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from run_fleet import fleet_consts, fleet_demand  # noqa: E402
from chargingstation.fleet import ChargingStationFleet  # noqa: E402

cases = [  # (stations, steps, evs, P, N_lo, N_bi, price_type, chain)
    (256, 30, 500, 12, 24, 24, "linear-convex", "reference"),
    (256, 30, 500, 12, 12, 16, "linear-convex", "reference"),
    (256, 30, 500, 12, 24, 24, "linear", "reference"),
    (128, 30, 200, 8, 12, 12, "linear", "partition"),
    (512, 20, 100, 12, 24, 24, "linear-convex", "partition"),
    (64, 40, 1000, 12, 12, 24, "linear-convex", "reference"),
]
for i, (S, T, M, P, N_lo, N_bi, pt, chain) in enumerate(cases):
    consts = fleet_consts(T, N_lo, N_bi, M, P)
    consts.price_type = pt
    demand = fleet_demand(consts, S, T, N_bi, seed=100 + i)
    fleet = ChargingStationFleet(consts, S, demand=demand, seed=100 + i, rng="device", chain=chain)
    log = fleet.simulate()
    bad_bimpc = int((log["bimpc_status"] != 0).sum())
    finite = all(bool(np.isfinite(v.cpu().numpy()[~np.isnan(v.cpu().numpy())]).all()) for k, v in log.items()
                 if v.dtype.is_floating_point)
    ni = np.concatenate([log["niter_s"].cpu().numpy().ravel(), log["niter_l"].cpu().numpy().ravel()])
    print(f"case {i}: S={S} T={T} M={M} P={P} N_lo={N_lo} N_bi={N_bi} {pt} {chain}: bimpc failures {bad_bimpc}, "
          f"finite {finite}, price iters mean {ni[ni >= 0].mean():.1f} capped {(ni >= 999).sum()}, "
          f"x in [{float(log['x'].min()):.4f}, {float(log['x'].max()):.4f}], qp solves {fleet.qp_solves}", flush=True)
    assert finite
print("stress ok")
