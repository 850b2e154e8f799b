#!/usr/bin/env python
"""Latency of ONE group's price loop (no contention): a group whose reference trajectory cannot be
tracked runs the full iteration cap; prints microseconds and SM cycles per MM iteration."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incentive-design-mpc_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from bench import EV_CONSTS  # noqa: E402
from chargingstation.lompc import LoMPCConstants  # noqa: E402
from chargingstation.price_solver import PriceSolver  # noqa: E402

N, nev, cap = 24, int(os.environ.get("NEV", "40")), 400
for ev in ("small", "large"):
    delta, theta, y_max, w_max = EV_CONSTS[ev]
    ps = PriceSolver(N, LoMPCConstants(delta, theta, y_max, w_max, ev), "linear-convex")
    ps.set_loop_mode(int(os.environ.get("LOOP_MODE", "0")))
    rng = np.random.default_rng(0)
    y0 = 0.3 + 0.05 * rng.random(nev)
    w_ref = np.zeros(N)
    w_ref[::2] = w_max  # alternating full / zero charging: not reachable within the tolerance
    off = np.array([0, nev], dtype=np.int32)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prices, st = ps.compute_optimal_prices_batch(off, y0, w_ref[None], np.zeros(1), np.zeros((1, 3 * N)), max_iter=cap)
        dt = time.perf_counter() - t0
    it = int(st["iter"][0]) + 1
    cq = ps._lib.price_last_cycles(ps._h, 0)
    cs = ps._lib.price_last_cycles(ps._h, 1)
    k1 = ps._lib.price_last_cycles(ps._h, 2)
    rounds = ps._lib.price_last_cycles(ps._h, 3)
    print(f"{ev}: {nev} EVs, {it} iterations, {dt * 1e6 / it:.1f} us/iteration (wall, incl. launch), "
          f"cycles/iteration: LoMPC pass {cq / it:.0f}, price step {cs / it:.0f}; QP solves {ps._lib.price_last_qp_solves(ps._h) / it:.1f} "
          f"per iteration, K1 iterations {k1 / it:.1f}, solver rounds {rounds / it:.2f} (parametric loop)")
