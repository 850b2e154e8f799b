#!/usr/bin/env python
"""Price loops at the long horizons of BASELINE configs[4] (N = 48, 96): the device-resident parametric loop (mode 2,
automatic) against the phase-split loop (mode 1), 1,024 groups x 32 EVs per EV type, capped at 60 iterations.
    python tools/time_long_horizon_loops.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
from chargingstation import settings  # noqa: E402
from chargingstation.lompc import LoMPCConstants  # noqa: E402
from chargingstation.price_solver import PriceSolver  # noqa: E402

settings.PRINT_LEVEL = 0
EV = {"small": (0.05, 10.0, 0.9, 0.25), "large": (0.025, 50.0, 0.9, 0.15)}
G, n = 1024, 32
for N in (48, 96):
    for ev, (delta, theta, y_max, w_max) in EV.items():
        rng = np.random.default_rng(N)
        off = (np.arange(G + 1) * n).astype(np.int32)
        y0 = np.sort(0.3 + 0.2 * rng.random(G * n)).reshape(G, n)[rng.permutation(G)].ravel()
        w_ref = w_max * rng.random((G, N)) * 0.6
        res = {}
        for mode in (2, 1):
            ps = PriceSolver(N, LoMPCConstants(delta, theta, y_max, w_max, ev), "linear-convex")
            ps.set_loop_mode(mode)
            args = (off, y0, w_ref, np.zeros(G), np.zeros((G, 3 * N)))
            ps.compute_optimal_prices_batch(*args, max_iter=60)
            t0 = time.perf_counter()
            prices, st = ps.compute_optimal_prices_batch(*args, max_iter=60)
            res[mode] = (time.perf_counter() - t0, prices, st["iter"].copy(), int(ps._lib.price_last_qp_solves(ps._h)))
        same = bool(np.array_equal(res[1][2], res[2][2]))
        print(json.dumps({"N": N, "ev": ev, "groups": G, "evs": G * n, "parametric_ms": res[2][0] * 1e3,
                          "phase_split_ms": res[1][0] * 1e3, "iters_mean": float(res[2][2].mean()),
                          "same_iteration_counts": same, "parametric_qp_solves": res[2][3],
                          "max_price_diff": float(np.max(np.abs(res[1][1] - res[2][1])[np.abs(res[1][1]) <= 1e3]))}), flush=True)
