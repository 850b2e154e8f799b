#!/usr/bin/env python
"""BASELINE.json configs[4]: horizon sweep N = 12/24/48/96 at 16,384 independent QPs (8,192 small + 8,192
large, inputs as test/test_lompc.py:34-36, seed 5).  Prints one JSON line per horizon: QP/s, iterations,
kernel variant, registers / shared memory limits are in profiles/ (ncu)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "incentive-design-mpc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from bench import EV_CONSTS  # noqa: E402
from chargingstation.lompc import LoMPC, LoMPCConstants  # noqa: E402

dev = torch.device("cuda:0")
B = 8192
for N in (12, 24, 48, 96):
    rng = np.random.default_rng(5)
    out = {"N": N, "batch": 2 * B, "kernel": "register-resident, one QP per thread (lompc_solve_reg_kernel)" if N in (12, 24)
           else "warp-cooperative, one QP per N/3 lanes (lompc_solve_warp_kernel)"}
    tot_ms = 0.0
    for ev in ("small", "large"):
        delta, theta, y_max, w_max = EV_CONSTS[ev]
        solver = LoMPC(N, LoMPCConstants(delta, theta, y_max, w_max, ev))
        lm = torch.from_numpy(theta * rng.random((B, 3 * N))).to(dev)
        lr = torch.from_numpy(3 * N * delta * rng.random(B)).to(dev)
        gam = torch.from_numpy(y_max * rng.random(B)).to(dev)
        w, cost, info = solver.solve_lompc_batch(lm, lr, gam, return_info=True)
        assert int(info["status"].max()) == 0 and float(info["kkt_res"].max()) <= 1e-10
        outb = (torch.empty((B, N), dtype=torch.float64, device=dev), torch.empty((B,), dtype=torch.float64, device=dev))
        for _ in range(3):
            solver.solve_lompc_batch(lm, lr, gam, out=outb)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            solver.solve_lompc_batch(lm, lr, gam, out=outb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot_ms += ms
        out[ev] = {"ms": ms, "qp_per_s": B / (ms * 1e-3), "iters_mean": float(info["iters"].double().mean()),
                   "iters_max": int(info["iters"].max()), "kkt_max": float(info["kkt_res"].max())}
    out["qp_per_s"] = 2 * B / (tot_ms * 1e-3)
    print(json.dumps(out), flush=True)
